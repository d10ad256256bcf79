// Device-initiated halo exchange of the single-column sweep kernels (multi-GPU; SURVEY.md section 8e, hard part H4).
//
// The time sweeps are ~10^4 dependent kernels of a few microseconds each; a separate exchange kernel (or an NCCL
// send/recv) per product would double that chain.  Instead the PRODUCER of a vector pushes the rows its neighbours
// will gather straight into their memory over NVLink, and the CONSUMER waits for the arrival inside the kernel that
// gathers:
//   * the first warps of a producing grid ("push warps") compute the listed boundary rows before anything else,
//     store them locally AND into the ghost slot of every rank that gathers them, fence, and publish a sequence
//     number into a per-chunk flag on that rank; the regular CTAs follow (they compute those rows once more, with
//     identical results);
//   * a consuming CTA whose rows gather ghost columns spins (bounded) until every chunk flag of the exchange carries
//     the expected sequence number, then reads the slot with L1-bypassing loads.
// Ghost values live in THREE rotating slots per vector space: a rank can run at most one dependent kernel ahead of a
// neighbour, and the neighbour may still be reading the slot of the exchange before (see halo.cu for the argument).
// Sequence numbers are (epoch << 24 | index): the index is static per launch (the sweeps are replayed from a CUDA
// graph), the epoch is a device word bumped once per replay after a barrier across ranks.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

struct HaloWait {
    const unsigned long long *flags = nullptr;   // local memory, written by the neighbours
    const unsigned long long *epoch = nullptr;
    int *err = nullptr;                          // set when the bounded wait gives up
    int n_flags = 0;                             // 0: nothing to wait for
    unsigned seq = 0;
    int skip_lo = 0, skip_hi = 0;                // rows in [skip_lo, skip_hi) gather no ghost column: their CTAs do not wait
    long long max_spins = 0;
};

struct PushDst {
    double *base;                 // slot 0 of the peer's ghost array
    unsigned long long *flag;     // the chunk's flag on the peer
    long long stride;             // doubles between two slots on the peer
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
    const unsigned long long *epoch = nullptr;
    int n_chunks = 0;                            // 0: nothing to push
    int slot = 0;
    unsigned seq = 0;
};

#ifdef __CUDACC__
// CTAs of `threads` threads: how many push CTAs precede the regular ones (one warp per chunk)
static inline int halo_push_ctas(const HaloPush &p, int threads) { return (p.n_chunks + threads / 32 - 1) / (threads / 32); }

__device__ __forceinline__ unsigned long long halo_seq(const unsigned long long *epoch, unsigned seq)
{
    return (*reinterpret_cast<const volatile unsigned long long *>(epoch) << 24) | (unsigned long long)seq;
}

// Called by every thread of the CTA before the first gather.  row_lo / row_hi: the rows this CTA computes.
__device__ __forceinline__ void halo_wait(const HaloWait &w, int row_lo, int row_hi)
{
    if (w.n_flags == 0) return;
    if (row_lo >= w.skip_lo && row_hi <= w.skip_hi) return;      // uniform per CTA
    const unsigned long long want = halo_seq(w.epoch, w.seq);
    for (int i = threadIdx.x; i < w.n_flags; i += blockDim.x) {
        const volatile unsigned long long *f = w.flags + i;
        long long spins = 0;
        while (*f < want) {
            __nanosleep(20);
            if (++spins > w.max_spins) {
                *w.err = 1;
                break;
            }
        }
    }
    __syncthreads();
    __threadfence();      // order the gathers below behind the flag reads (acquire side of the fence pairing)
}

// one gathered entry: owned columns through the read-only path, ghost columns from the slot, around L1
__device__ __forceinline__ double halo_gather(const double *__restrict__ x, const double *ghost, int n_own, int c)
{
    if (c < n_own) return __ldg(x + c);
    return __ldcg(ghost + (c - n_own));
}

// Once every lane of a push warp has stored its rows (locally and on the peers): the publication of the chunk
__device__ __forceinline__ void halo_push_publish(const HaloPush &p, const PushChunk &ch, int lane)
{
    __threadfence_system();
    __syncwarp();
    if (lane < ch.n_dst) {
        const unsigned long long seq = halo_seq(p.epoch, p.seq);
        *reinterpret_cast<volatile unsigned long long *>(p.dsts[ch.dst_begin + lane].flag) = seq;
    }
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
    bool replicate = false;                  // every rank receives every row: slots hold the whole vector, own rows first
    // producer side (device arrays, shared between the instances of a space)
    const PushChunk *d_chunks = nullptr;
    const PushDst *d_dsts = nullptr;         // per instance: the destinations point into the peers' instance
    const int *d_rows = nullptr;
    int n_chunks = 0;
    // consumer side (this rank's arena)
    double *slots = nullptr;                 // 3 x stride doubles
    long long stride = 0;
    unsigned long long *flags = nullptr;
    int n_flags = 0;
    const unsigned long long *d_epoch = nullptr;
    int *d_err = nullptr;
    long long max_spins = 0;
    // static exchange counter (host): the index of the next exchange inside the current epoch
    unsigned idx = 0;
};

// the next exchange of the plan, to be handed to the kernel that produces the vector (null plan: nothing to push)
HaloPush halo_push(HaloPlan *p);
// ghosts of the plan's LAST exchange (slot pointer and the wait descriptor of a consuming kernel)
const double *halo_ghost(const HaloPlan *p);
HaloWait halo_wait_for(const HaloPlan *p, int skip_lo, int skip_hi);
// replicating plan: where the producer of the NEXT exchange writes its own rows / where the consumer of the
// LAST exchange reads the whole vector
double *halo_full_next(HaloPlan *p);
const double *halo_full_last(const HaloPlan *p);
// push the boundary rows of a vector that some other kernel produced (one small kernel)
int halo_exchange_now(ctl_handle_s *h, HaloPlan *p, const double *x);
// copy the ghosts of the last exchange behind the owned entries of x (they outlive the slot rotation there)
int halo_persist(ctl_handle_s *h, HaloPlan *p, double *x_tail);
