// Host side of the device-initiated halo exchange (halo.cu): device arrays of a vector space's send side, the
// IPC-shared arena that holds every rank's ghost slots and flags.  The geometry itself is pure host code: halo_geom.h.
#pragma once
#include <memory>
#include <vector>

#include "common.cuh"
#include "halo.cuh"
#include "halo_geom.h"

struct HaloSpace {
    std::shared_ptr<HaloGeom> g;
    // device arrays of THIS rank's send side
    int *d_rows = nullptr;
    PushChunk *d_chunks = nullptr;
    int *d_pos = nullptr;
    int n_pairs = 0;                              // (chunk, destination) pairs of this rank
    ~HaloSpace();
};

// upload the send side of a geometry
int halo_space_upload(ctl_handle_s *h, const std::shared_ptr<HaloGeom> &g, std::shared_ptr<HaloSpace> &out);

// One cudaMalloc per rank, opened by every peer through CUDA IPC: the ghost slots of every plan instance.
struct HaloArena {
    struct Inst {
        std::shared_ptr<HaloSpace> sp;
        std::unique_ptr<HaloPlan> plan;
        std::vector<size_t> slot_off;             // [rank] byte offset inside that rank's arena
        PushDst *d_dsts = nullptr;
    };
    std::vector<size_t> cursor;                   // [rank] bytes claimed so far
    std::vector<Inst> inst;
    char *base = nullptr;
    std::vector<char *> peer_base;
    std::vector<void *> opened;
    bool finalized = false;
};

void halo_arena_free(ctl_handle_s *h, HaloArena &a);
// claim room for one more exchange stream of a space on every rank (before halo_arena_finalize)
HaloPlan *halo_arena_add(ctl_handle_s *h, HaloArena &a, const std::shared_ptr<HaloSpace> &sp);
// allocate, exchange the IPC handles, resolve every destination pointer (collective over the communicator)
int halo_arena_finalize(ctl_handle_s *h, HaloArena &a);
// start of a sequence of exchanges (captured at the head of the sweep graph): barrier across ranks, new epoch,
// static exchange counters back to zero
int halo_epoch_begin(ctl_handle_s *h);
