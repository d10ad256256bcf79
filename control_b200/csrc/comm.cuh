// Communicator state shared by comm.cu (NCCL: time-batched halo panels, Krylov all-reduces) and halo.cu (the
// device-initiated exchange of the single-column sweep kernels).
#pragma once
#include <nccl.h>

#include <vector>

#include "common.cuh"
#include "halo.cuh"

struct CommState {
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1;
    // ---- NCCL exchange of time-fastest panels (KKT apply, batched Chebyshev, Schur right-hand side)
    std::vector<int> peers;               // ranks we exchange with
    std::vector<int> recv_off, recv_cnt;  // ghost segment of each peer (rows)
    std::vector<int> send_off, send_cnt;  // segment of each peer in the packed send list
    int n_send = 0;
    int *d_send_rows = nullptr;           // owned local rows to pack, peer after peer
    double *d_sendbuf = nullptr;          // [2][n_send x ld]
    // ---- device-initiated exchange (halo.cu)
    unsigned long long *d_epoch = nullptr;    // bumped once per sweep replay, after a barrier across ranks
    int *d_err = nullptr;                     // a bounded wait gave up
    int *d_barrier = nullptr;                 // operand of the barrier all-reduce
    long long max_spins = 0;
    std::vector<HaloPlan *> plans;            // live plan instances: their static counters restart with every epoch
};

#define CTL_NCCL(call)                                                                     \
    do {                                                                                   \
        ncclResult_t r__ = (call);                                                         \
        if (r__ != ncclSuccess) {                                                          \
            ctl_set_error(h, std::string(#call) + ": " + ncclGetErrorString(r__));         \
            return CTL_ERR_NCCL;                                                           \
        }                                                                                  \
    } while (0)
