// Device-initiated halo exchange of the single-column sweep kernels (multi-GPU; SURVEY.md section 8e, hard part H4).
//
// The time sweeps are ~10^4 dependent kernels of a few microseconds each, and the boundary rows of step k on one
// rank need the boundary rows of step k-1 of its neighbours: a latency-bound ping-pong.  A separate exchange kernel,
// an NCCL send/recv or a fence + flag handshake per product would cost more than the products.  Instead:
//   * the PRODUCER of a vector pushes the rows its neighbours will gather straight into their memory over NVLink:
//     the first warps of the producing grid ("push warps") compute the listed boundary rows before anything else and
//     store them locally AND into the ghost slot of every rank that gathers them; the regular CTAs follow (they
//     compute those rows once more, with identical results);
//   * every pushed value travels as two 8-byte words (low half | sequence number, high half | sequence number) in one
//     16-byte store -- the protocol NCCL calls LL: an 8-byte store is atomic, so a reader that sees the expected
//     sequence number in both words has the value; no fence, no separate flag, no round trip;
//   * the CONSUMER gathers a ghost entry by spinning (bounded) on exactly that entry until both words carry the
//     sequence number of the exchange it expects; rows that gather no ghost column never wait.
// Ghost values live in THREE rotating slots per exchange stream: a rank can run at most one dependent kernel ahead
// of a neighbour, and the neighbour may still be reading the slot of the exchange before (argument in halo.cu).
// Sequence numbers are (epoch << 20 | index + 1), 32 bits: the index is static per launch (the sweeps are replayed
// from a CUDA graph), the epoch is a device word bumped once per replay after a barrier across ranks.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

struct HaloCtx {                                 // per-handle constants of the exchange, passed to every kernel
    const unsigned long long *epoch = nullptr;   // null: one GPU
    int *err = nullptr;                          // set when a bounded spin gives up
    long long max_spins = 0;
    int early_wait = 0;                          // CTL_MP_EARLY_WAIT=1: kernels wait for their predecessor before
                                                 // anything else (no loads of constant data ahead of the wait)
};

struct PushDst {
    ulonglong2 *base;             // slot 0 of the peer's ghost array
    long long stride;             // entries between two slots on the peer
    const int *pos;               // [<= 32] position of each row of the chunk inside the peer's ghost array
};

struct PushChunk {
    int start, count;             // segment of the plan's row list (count <= 32)
    int dst_begin, n_dst;         // destinations (n_dst > 1: the replicating exchange)
};

struct HaloPush {
    const PushChunk *chunks = nullptr;
    const PushDst *dsts = nullptr;
    const int *rows = nullptr;                   // local rows to push, chunk after chunk
    int n_chunks = 0;                            // 0: nothing to push
    int slot = 0;
    unsigned idx1 = 0;                           // index + 1 of the exchange inside the epoch
    int all_rows = 0;                            // the chunks cover every row of the grid, in order (replicating
                                                 // exchange): the regular CTAs have nothing left to do
};

#ifdef __CUDACC__
// CTAs of `threads` threads: how many push CTAs precede the regular ones (one warp per chunk)
static inline int halo_push_ctas(const HaloPush &p, int threads) { return (p.n_chunks + threads / 32 - 1) / (threads / 32); }

// high bits of every sequence number of the running epoch
__device__ __forceinline__ unsigned halo_epoch_bits(const HaloCtx &c)
{
    if (!c.epoch) return 0u;
    return (unsigned)(*reinterpret_cast<const volatile unsigned long long *>(c.epoch)) << 20;
}

__device__ __forceinline__ void halo_ll_store(ulonglong2 *p, double v, unsigned seq)
{
    const unsigned long long s = (unsigned long long)seq << 32;
    const unsigned long long lo = (unsigned long long)(unsigned)__double2loint(v) | s;
    const unsigned long long hi = (unsigned long long)(unsigned)__double2hiint(v) | s;
    asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(lo), "l"(hi) : "memory");
}

__device__ __forceinline__ double halo_ll_read(const ulonglong2 *p, unsigned seq, const HaloCtx &c)
{
    unsigned long long lo, hi;
    long long spins = 0;
    while (true) {
        asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(lo), "=l"(hi) : "l"(p) : "memory");
        if ((unsigned)(lo >> 32) == seq && (unsigned)(hi >> 32) == seq) break;
        if (++spins > c.max_spins) {
            *c.err = 1;
            break;
        }
        if (spins > 64) __nanosleep(40);
    }
    return __hiloint2double((int)(unsigned)hi, (int)(unsigned)lo);
}
#endif

// ---------------------------------------------------------------------------------------------------------
// host side (halo.cu)
// ---------------------------------------------------------------------------------------------------------
struct ctl_handle_s;
struct GVec;

// One exchange stream of a vector space (a hierarchy level on this rank): who gathers which of my rows, where my
// ghosts arrive.  Two instances share the geometry of a space: one for the iterates, one for the right-hand side
// (whose ghosts must survive the smoothing steps in between).
struct HaloPlan {
    int n_own = 0, n_ghost = 0;
    bool replicate = false;                  // every rank receives every row of every other rank
    // producer side (device arrays; chunks and rows are shared between the instances of a space)
    const PushChunk *d_chunks = nullptr;
    const PushDst *d_dsts = nullptr;         // per instance: the destinations point into the peers' instance
    const int *d_rows = nullptr;
    int n_chunks = 0;
    // consumer side (this rank's arena): 3 slots of `stride` 16-byte entries
    ulonglong2 *slots = nullptr;
    long long stride = 0;
    // static exchange counter (host): the index of the next exchange inside the current epoch
    unsigned idx = 0;
};

// per-handle constants for the kernels (all zero on one GPU)
HaloCtx halo_ctx(const ctl_handle_s *h);
// the next exchange of the plan, to be handed to the kernel that produces the vector (null plan: nothing to push)
HaloPush halo_push(HaloPlan *p);
// the ghosts of the plan's LAST exchange, as a consuming kernel reads them: slot pointer and index + 1
const ulonglong2 *halo_ll(const HaloPlan *p);
unsigned halo_idx1(const HaloPlan *p);
// push the boundary rows of a vector that some other kernel produced (one small kernel)
int halo_exchange_now(ctl_handle_s *h, HaloPlan *p, const double *x);
// wait for the ghosts of the last exchange and store them plainly at dst[0 .. n_ghost): behind the owned entries of a
// sweep vector (they outlive the slot rotation there), or behind the own rows of a replicated right-hand side
int halo_unpack(ctl_handle_s *h, HaloPlan *p, double *dst);
