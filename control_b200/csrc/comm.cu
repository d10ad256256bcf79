// Multi-GPU plumbing: one process per GPU, rows of the spatial mesh partitioned in
// contiguous blocks (PETSc's ownership split, as the reference's MPIAIJ matrices are
// distributed: preconditioner/preconditioner.py:706-722 sizes the shell Mat ((n_local,
// N_global), ...)).  Every rank keeps ALL N time columns of its rows, so the time axis
// never communicates.  Two collectives over NCCL / NVLink, both enqueued on the handle's
// stream (and therefore capturable into the sweep CUDA graph):
//   * halo exchange of boundary rows before a sparse product (what PETSc's VecScatter
//     does inside MatMult_MPIAIJ): rows x ld doubles per neighbour for the time-batched
//     kernels, one double per row for the single-column sweep kernels;
//   * all-reduce of a handful of doubles per Krylov iteration (VecMDot / VecNorm) and of
//     the restricted coarse right-hand side inside an AMG cycle (coarse levels are
//     replicated on every rank).
// The exchange plan needs no communication: every rank holds the global sparsity pattern
// and derives, deterministically, both what it needs and what each peer needs from it.
#include <nccl.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "common.cuh"

struct CommState {
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1;
    std::vector<int> peers;               // ranks we exchange with
    std::vector<int> recv_off, recv_cnt;  // ghost segment of each peer (rows)
    std::vector<int> send_off, send_cnt;  // segment of each peer in the packed send list
    int n_send = 0;
    int *d_send_rows = nullptr;           // owned local rows to pack, peer after peer
    double *d_sendbuf = nullptr;          // [2][n_send x ld]
    // peer-to-peer path of the single-vector exchange (the latency-critical one inside the
    // time sweeps): boundary values are stored straight into the neighbour's mailbox over
    // NVLink and signalled with a sequence number; no NCCL kernel, no host involvement
    bool p2p = false;
    std::vector<int> peer_halo_n, peer_dst_off;   // host: per peer, its ghost count and where my rows land in it
    double *mailbox = nullptr;                    // local, [2 slots][n_halo]
    unsigned long long *flags = nullptr;          // local, [world]: flags[p] = last sequence peer p delivered
    unsigned long long *d_seq = nullptr;          // local sequence counter
    int *d_p2p_err = nullptr;
    struct PeerDesc *d_peers = nullptr;
    std::vector<void *> opened;                   // cudaIpcOpenMemHandle results
};

struct PeerDesc {
    double *mailbox;               // peer's mailbox (mapped)
    unsigned long long *flags;     // peer's flag array (mapped)
    int send_off, send_cnt;        // my packed send list segment
    int dst_off, halo_n;           // offset inside the peer's ghost ordering, peer's ghost count (slot stride)
    int peer_rank;
};

#define CTL_NCCL(call)                                                                     \
    do {                                                                                   \
        ncclResult_t r__ = (call);                                                         \
        if (r__ != ncclSuccess) {                                                          \
            ctl_set_error(h, std::string(#call) + ": " + ncclGetErrorString(r__));         \
            return CTL_ERR_NCCL;                                                           \
        }                                                                                  \
    } while (0)

void ctl_comm_free(ctl_handle_s *h)
{
    if (!h->comm) return;
    CommState &c = *h->comm;
    cudaFree(c.d_send_rows);
    cudaFree(c.d_sendbuf);
    for (void *p : c.opened) cudaIpcCloseMemHandle(p);
    cudaFree(c.mailbox);
    cudaFree(c.flags);
    cudaFree(c.d_seq);
    cudaFree(c.d_p2p_err);
    cudaFree(c.d_peers);
    if (c.comm) ncclCommDestroy(c.comm);
    h->comm.reset();
}

namespace {

void owner_range(int n, int world, int rank, int *begin, int *count)
{
    const int base = n / world, rem = n % world;
    *count = base + (rank < rem ? 1 : 0);
    *begin = rank * base + std::min(rank, rem);
}

// gather rows of a [rows x width] row-major array: out[i, :] = in[rows[i], :]
__global__ void pack_rows_kernel(const double *__restrict__ in, const int *__restrict__ rows, int n_rows, int width,
                                 double *__restrict__ out)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)n_rows * width) return;
    const int i = (int)(t / width), j = (int)(t % width);
    out[t] = in[(size_t)rows[i] * width + j];
}

}  // namespace

// plan: ghosts of every rank owned by me (send) and my ghosts owned by every rank (recv)
static int build_plan(ctl_handle_s *h, CommState &c)
{
    const int n = h->n, world = c.world, me = c.rank;
    const std::vector<int> &ip = h->h_indptr, &ix = h->h_indices;
    int my_b, my_c;
    owner_range(n, world, me, &my_b, &my_c);
    std::vector<int> send_rows;
    for (int p = 0; p < world; ++p) {
        if (p == me) continue;
        int pb, pc;
        owner_range(n, world, p, &pb, &pc);
        // what p needs from me: columns in my range referenced by p's rows (sorted, unique):
        // exactly the segment of p's ghost list that I own, in p's ghost order
        std::vector<int> need, pghost;
        for (int r = pb; r < pb + pc; ++r)
            for (int k = ip[r]; k < ip[r + 1]; ++k) {
                if (ix[k] >= my_b && ix[k] < my_b + my_c) need.push_back(ix[k]);
                if (ix[k] < pb || ix[k] >= pb + pc) pghost.push_back(ix[k]);
            }
        std::sort(need.begin(), need.end());
        need.erase(std::unique(need.begin(), need.end()), need.end());
        std::sort(pghost.begin(), pghost.end());
        pghost.erase(std::unique(pghost.begin(), pghost.end()), pghost.end());
        // what I need from p: my ghosts in p's range (halo_global is sorted by global id)
        const auto lo = std::lower_bound(h->halo_global.begin(), h->halo_global.end(), pb);
        const auto hi = std::lower_bound(h->halo_global.begin(), h->halo_global.end(), pb + pc);
        const int rc = (int)(hi - lo);
        if (need.empty() && rc == 0) continue;
        c.peers.push_back(p);
        c.recv_off.push_back((int)(lo - h->halo_global.begin()));
        c.recv_cnt.push_back(rc);
        c.send_off.push_back((int)send_rows.size());
        c.send_cnt.push_back((int)need.size());
        c.peer_halo_n.push_back((int)pghost.size());
        c.peer_dst_off.push_back((int)(std::lower_bound(pghost.begin(), pghost.end(), my_b) - pghost.begin()));
        for (int g : need) send_rows.push_back(g - my_b);
    }
    c.n_send = (int)send_rows.size();
    CTL_TRY(ctl_upload(h, &c.d_send_rows, send_rows.data(), send_rows.size()));
    if (c.n_send > 0)
        CTL_CUDA(cudaMalloc((void **)&c.d_sendbuf, (size_t)2 * c.n_send * h->ld * sizeof(double)));
    return CTL_OK;
}

// exchange `panels` time-fastest panels: ghost rows land in h->d_halo[p]
static int exchange_panels(ctl_handle_s *h, const double *x_tf, int panels)
{
    CommState &c = *h->comm;
    const int ld = h->ld;
    const size_t panel = (size_t)h->n_loc * ld;
    for (int p = 0; p < panels; ++p) {
        if (c.n_send == 0) break;
        const int64_t total = (int64_t)c.n_send * ld;
        pack_rows_kernel<<<ceil_div(total, 256), 256, 0, h->stream>>>(x_tf + p * panel, c.d_send_rows, c.n_send, ld,
                                                                     c.d_sendbuf + (size_t)p * c.n_send * ld);
        h->launches++;
    }
    CTL_CUDA(cudaGetLastError());
    CTL_NCCL(ncclGroupStart());
    for (size_t i = 0; i < c.peers.size(); ++i) {
        for (int p = 0; p < panels; ++p) {
            if (c.send_cnt[i])
                CTL_NCCL(ncclSend(c.d_sendbuf + ((size_t)p * c.n_send + c.send_off[i]) * ld, (size_t)c.send_cnt[i] * ld,
                                  ncclDouble, c.peers[i], c.comm, h->stream));
            if (c.recv_cnt[i])
                CTL_NCCL(ncclRecv(h->d_halo + ((size_t)p * h->n_halo + c.recv_off[i]) * ld, (size_t)c.recv_cnt[i] * ld,
                                  ncclDouble, c.peers[i], c.comm, h->stream));
        }
    }
    CTL_NCCL(ncclGroupEnd());
    return CTL_OK;
}

int ctl_halo_exchange(ctl_handle_s *h, const double *x_tf)
{
    if (!h->comm || h->n_halo == 0) return CTL_OK;
    return exchange_panels(h, x_tf, 2);
}

int ctl_halo_exchange_panel(ctl_handle_s *h, const double *panel_tf)
{
    if (!h->comm || h->n_halo == 0) return CTL_OK;
    return exchange_panels(h, panel_tf, 1);
}

namespace {

// One CTA per rank.  Phase 1: store my boundary values into every neighbour's mailbox slot and
// publish the sequence number.  Phase 2: wait for every neighbour's number, then move the slot
// into the ghost region of x.  Two slots (sequence parity) are enough: a neighbour can only be
// one exchange ahead, because finishing exchange s+1 needs my flag s+1, which I publish after
// I have emptied slot s.
__global__ void __launch_bounds__(1024) halo_p2p_kernel(double *x, int n_loc, int n_halo, const int *__restrict__ send_rows,
                                                       const PeerDesc *__restrict__ peers, int n_peers, const double *mailbox,
                                                       volatile unsigned long long *my_flags, unsigned long long *seq_counter,
                                                       int my_rank, int *err)
{
    __shared__ unsigned long long s_seq;
    if (threadIdx.x == 0) s_seq = *seq_counter + 1ull;
    __syncthreads();
    const unsigned long long seq = s_seq;
    const int slot = (int)(seq & 1ull);
    for (int i = 0; i < n_peers; ++i) {
        const PeerDesc p = peers[i];
        double *dst = p.mailbox + (size_t)slot * p.halo_n + p.dst_off;
        for (int j = threadIdx.x; j < p.send_cnt; j += blockDim.x) dst[j] = x[send_rows[p.send_off + j]];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x < n_peers) {
        *reinterpret_cast<volatile unsigned long long *>(peers[threadIdx.x].flags + my_rank) = seq;
        // wait for the neighbour's delivery of the same exchange (bounded: never hang the GPU)
        const int pr = peers[threadIdx.x].peer_rank;
        long spins = 0;
        while (my_flags[pr] < seq) {
            __nanosleep(40);
            if (++spins > 50000000L) {
                *err = 1;
                break;
            }
        }
    }
    __syncthreads();
    __threadfence_system();
    const double *src = mailbox + (size_t)slot * n_halo;
    for (int j = threadIdx.x; j < n_halo; j += blockDim.x) x[n_loc + j] = __ldcv(src + j);
    if (threadIdx.x == 0) *seq_counter = seq;
}

}  // namespace

static int setup_p2p(ctl_handle_s *h, CommState &c)
{
    if (const char *e = getenv("CTL_NO_P2P"))
        if (e[0] == '1') return CTL_OK;
    const int world = c.world;
    const size_t mb = (size_t)2 * std::max(h->n_halo, 1) * sizeof(double);
    CTL_CUDA(cudaMalloc((void **)&c.mailbox, mb));
    CTL_CUDA(cudaMemset(c.mailbox, 0, mb));
    CTL_CUDA(cudaMalloc((void **)&c.flags, world * sizeof(unsigned long long)));
    CTL_CUDA(cudaMemset(c.flags, 0, world * sizeof(unsigned long long)));
    CTL_CUDA(cudaMalloc((void **)&c.d_seq, sizeof(unsigned long long)));
    CTL_CUDA(cudaMemset(c.d_seq, 0, sizeof(unsigned long long)));
    CTL_CUDA(cudaMalloc((void **)&c.d_p2p_err, sizeof(int)));
    CTL_CUDA(cudaMemset(c.d_p2p_err, 0, sizeof(int)));
    // exchange the IPC handles of (mailbox, flags) through the NCCL communicator itself
    struct Handles {
        cudaIpcMemHandle_t mailbox, flags;
    } mine;
    if (cudaIpcGetMemHandle(&mine.mailbox, c.mailbox) != cudaSuccess || cudaIpcGetMemHandle(&mine.flags, c.flags) != cudaSuccess) {
        cudaGetLastError();
        return CTL_OK;          // no IPC (e.g. restricted container): stay on the NCCL path
    }
    Handles *d_all = nullptr;
    CTL_CUDA(cudaMalloc((void **)&d_all, world * sizeof(Handles)));
    CTL_CUDA(cudaMemcpy(d_all + c.rank, &mine, sizeof(Handles), cudaMemcpyHostToDevice));
    CTL_NCCL(ncclAllGather(d_all + c.rank, d_all, sizeof(Handles), ncclChar, c.comm, h->stream));
    CTL_CUDA(cudaStreamSynchronize(h->stream));
    std::vector<Handles> all(world);
    CTL_CUDA(cudaMemcpy(all.data(), d_all, world * sizeof(Handles), cudaMemcpyDeviceToHost));
    cudaFree(d_all);
    std::vector<PeerDesc> descs;
    bool ok = true;
    for (size_t i = 0; i < c.peers.size() && ok; ++i) {
        const int p = c.peers[i];
        void *pm = nullptr, *pf = nullptr;
        if (cudaIpcOpenMemHandle(&pm, all[p].mailbox, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess ||
            cudaIpcOpenMemHandle(&pf, all[p].flags, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
            cudaGetLastError();
            ok = false;
            break;
        }
        c.opened.push_back(pm);
        c.opened.push_back(pf);
        PeerDesc d;
        d.mailbox = (double *)pm;
        d.flags = (unsigned long long *)pf;
        d.send_off = c.send_off[i];
        d.send_cnt = c.send_cnt[i];
        d.dst_off = c.peer_dst_off[i];
        d.halo_n = c.peer_halo_n[i];
        d.peer_rank = p;
        descs.push_back(d);
    }
    // every rank must take the same path: agree through an all-reduce of the success flags
    int *d_ok = nullptr;
    CTL_CUDA(cudaMalloc((void **)&d_ok, sizeof(int)));
    const int mine_ok = ok ? 1 : 0;
    CTL_CUDA(cudaMemcpy(d_ok, &mine_ok, sizeof(int), cudaMemcpyHostToDevice));
    CTL_NCCL(ncclAllReduce(d_ok, d_ok, 1, ncclInt, ncclMin, c.comm, h->stream));
    CTL_CUDA(cudaStreamSynchronize(h->stream));
    int all_ok = 0;
    CTL_CUDA(cudaMemcpy(&all_ok, d_ok, sizeof(int), cudaMemcpyDeviceToHost));
    cudaFree(d_ok);
    if (!all_ok || c.peers.size() > 1024) return CTL_OK;
    CTL_CUDA(cudaMalloc((void **)&c.d_peers, std::max<size_t>(descs.size(), 1) * sizeof(PeerDesc)));
    CTL_CUDA(cudaMemcpy(c.d_peers, descs.data(), descs.size() * sizeof(PeerDesc), cudaMemcpyHostToDevice));
    c.p2p = true;
    return CTL_OK;
}

// single spatial vector with ghost entries appended behind the n_loc owned ones
int ctl_halo_exchange_vec(ctl_handle_s *h, double *x)
{
    if (!h->comm || h->n_halo == 0) return CTL_OK;
    CommState &c = *h->comm;
    if (c.p2p) {
        halo_p2p_kernel<<<1, 1024, 0, h->stream>>>(x, h->n_loc, h->n_halo, c.d_send_rows, c.d_peers, (int)c.peers.size(), c.mailbox,
                                                  c.flags, c.d_seq, c.rank, c.d_p2p_err);
        h->launches++;
        CTL_CUDA(cudaGetLastError());
        return CTL_OK;
    }
    if (c.n_send > 0) {
        pack_rows_kernel<<<ceil_div(c.n_send, 256), 256, 0, h->stream>>>(x, c.d_send_rows, c.n_send, 1, c.d_sendbuf);
        h->launches++;
        CTL_CUDA(cudaGetLastError());
    }
    CTL_NCCL(ncclGroupStart());
    for (size_t i = 0; i < c.peers.size(); ++i) {
        if (c.send_cnt[i])
            CTL_NCCL(ncclSend(c.d_sendbuf + c.send_off[i], (size_t)c.send_cnt[i], ncclDouble, c.peers[i], c.comm, h->stream));
        if (c.recv_cnt[i])
            CTL_NCCL(ncclRecv(x + h->n_loc + c.recv_off[i], (size_t)c.recv_cnt[i], ncclDouble, c.peers[i], c.comm, h->stream));
    }
    CTL_NCCL(ncclGroupEnd());
    return CTL_OK;
}

// has a peer-to-peer exchange given up waiting (bounded spin)?  Called at the end of a solve.
int ctl_comm_check(ctl_handle_s *h)
{
    if (!h->comm || !h->comm->p2p) return CTL_OK;
    int err = 0;
    CTL_CUDA(cudaMemcpyAsync(&err, h->comm->d_p2p_err, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CTL_CUDA(cudaStreamSynchronize(h->stream));
    CTL_CHECK(err == 0, CTL_ERR_NCCL, "peer-to-peer halo exchange timed out waiting for a neighbour");
    return CTL_OK;
}

int ctl_allreduce_sum(ctl_handle_s *h, double *dev, int count)
{
    if (!h->comm) return CTL_OK;
    CTL_NCCL(ncclAllReduce(dev, dev, (size_t)count, ncclDouble, ncclSum, h->comm->comm, h->stream));
    return CTL_OK;
}

extern "C" {

int ctl_comm_unique_id(void *id128)
{
    if (!id128) return CTL_ERR_ARG;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is expected to be 128 bytes");
    ncclUniqueId id;
    if (ncclGetUniqueId(&id) != ncclSuccess) return CTL_ERR_NCCL;
    memcpy(id128, &id, sizeof(id));
    return CTL_OK;
}

int ctl_comm_init(ctl_handle h, const void *id128)
{
    CTL_CHECK(h && id128, CTL_ERR_ARG, "ctl_comm_init: null argument");
    CTL_CHECK(h->cfg.world > 1, CTL_ERR_STATE, "ctl_comm_init: the handle was created with world = 1");
    CTL_CHECK(h->assembled, CTL_ERR_STATE, "ctl_comm_init: call ctl_assemble first");
    CTL_CUDA(cudaSetDevice(h->cfg.device));
    ctl_comm_free(h);
    h->comm = std::make_shared<CommState>();
    CommState &c = *h->comm;
    c.rank = h->cfg.rank;
    c.world = h->cfg.world;
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    CTL_NCCL(ncclCommInitRank(&c.comm, c.world, id, c.rank));
    CTL_TRY(build_plan(h, c));
    return setup_p2p(h, c);
}

}  // extern "C"
