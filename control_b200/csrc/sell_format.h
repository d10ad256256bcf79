// Storage formats of the single-column sweep matrices (time sweeps, AMG cycles) and their decoding.
//
// A sweep is ~10^4 dependent products with the SAME few matrices (control/control.py:2053-2189: every time step
// re-applies the same diagonal / off-diagonal blocks), so the bytes of a matrix are paid 20 times per time step.
// All formats below are EXACT re-encodings of (column, fp64 value): the products are bit-identical to the plain
// int32 + fp64 format, only the bytes per stored entry change.
//
//   FMT_F64     int32 column + fp64 value                                     12 B / entry
//   FMT_D16     uint16 (column - base) + fp64 value                           10 B   base: per SELL slice / per CSR row
//   FMT_PK      uint16 (column - base) + uint16 value code -> fp64 table       4 B   few distinct values (P, R on uniform meshes)
//   FMT_DICT16  uint16 code -> (column - row, fp64 value) table                2 B   translation-invariant stencils
//   FMT_DICT8   uint8  code -> (column - row, fp64 value) table                1 B   ... with at most 256 distinct pairs
//   FMT_STENCIL DICT16 whose SELL slices mostly consist of 32 rows with the SAME sequence of (column - row, value)
//               pairs (the interior of a uniform mesh in its natural numbering): such a slice stores one stencil
//               id, its rows gather x[row + delta_k] with warp-uniform delta_k and value_k -- about 5 instructions
//               per entry instead of 20, which is what bounds these kernels once the matrix stream is gone.
//               The other slices (mesh boundary, partition boundary) keep their per-entry DICT16 codes.
//
// 16-bit column offsets that do not fit (0xFFFF) escape to the int32 column array, which is always kept.
// Which format a matrix gets is decided from its entries alone (sell_choose_*): a mesh matrix assembled on a uniform
// mesh (every BASELINE config) has a handful of distinct (offset, value) pairs, a Galerkin operator of such a matrix a
// few thousand; a matrix with arbitrary values (unstructured mesh, per-level Jacobians) falls back to D16 or F64.
//
// The decode functions are __host__ __device__ so that tests/native/sell_format_check.cu runs the kernels' own indexing
// on the CPU.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <cstring>
#include <vector>

#include "pdl.cuh"

#ifdef __CUDACC__
#define SF_HD __host__ __device__ __forceinline__
#else
#define SF_HD inline
#endif

enum SellFmt { FMT_F64 = 0, FMT_D16 = 1, FMT_PK = 2, FMT_DICT16 = 3, FMT_DICT8 = 4, FMT_STENCIL = 5 };
constexpr unsigned SF_ESCAPE = 0xFFFFu;
constexpr int SF_PS = 16;         // longest stencil that can travel with the kernel parameters
// Dictionaries must stay L1-resident: a lookup with 32 different codes per warp is a gather, and a gather from L2
// costs as much as the one from x it feeds (measured: a 9175-pair dictionary made the 175 k-row level of C2 slower than
// the plain format).  4096 entries = 32 KB (values) / 64 KB (pairs).
constexpr int SF_DICT_MAX = 4096;

struct alignas(16) DictEnt {
    int delta;       // column - row
    int pad;
    double v;
};

// Pointers of one matrix in any format (device pointers inside kernels, host pointers in the CPU check).
struct MatView {
    int fmt = FMT_F64;
    int n_rows = 0;
    int n_own = 0;                    // columns < n_own are gathered from x, the others from the ghost array
    const int2 *sp = nullptr;         // SELL-32: per slice (first stored position, column base), n_slices + 1
    const int *ptr = nullptr;         // CSR-vector: row pointers ...
    const int *rbase = nullptr;       // ... and per-row column base (16-bit formats)
    const int *cols = nullptr;        // int32 columns (F64; escape entries of D16 / PK)
    const uint16_t *dcol = nullptr;   // D16 / PK
    const double *vals = nullptr;     // F64 / D16
    const uint16_t *vcode = nullptr;  // PK
    const double *vdict = nullptr;    // PK
    const void *code = nullptr;       // DICT8 / DICT16 / STENCIL (non-uniform slices)
    const DictEnt *dict = nullptr;    // DICT8 / DICT16 / STENCIL
    const int4 *sp4 = nullptr;        // STENCIL: per slice (first stored position, width, stencil offset or -1, 0)
    const DictEnt *stab = nullptr;    // STENCIL: the stencils, `width` entries each
    // STENCIL: the most frequent stencil travels with the kernel parameters (constant bank: its offsets and values
    // cost no load instructions); ps_off = its offset in stab (-2: none), at most SF_PS entries (15 = P1 tetrahedra)
    int ps_off = -2, ps_w = 0;
    int ps_dmax = 0;                  // largest offset of that stencil: rows below n_own - ps_dmax gather no ghost
    int ps_dmin = 0;                  // smallest offset (<= 0 on a square matrix)
    int ps_delta[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    double ps_v[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
};

// PRE: a load of the first chunk of a row, issued before the wait for the predecessor grid (pdl.cuh); it must keep
// its place in the instruction stream
template <bool STREAM, bool PRE = false, typename T>
SF_HD T sf_ld(const T *p)
{
#ifdef __CUDA_ARCH__
    if (PRE) return (T)pre_ld<STREAM>(p);
    return STREAM ? __ldcs(p) : __ldg(p);      // STREAM: read once per pass, evict first
#else
    return *p;
#endif
}

SF_HD DictEnt sf_dict(const DictEnt *d, unsigned k)
{
#ifdef __CUDA_ARCH__
    const int4 q = __ldg(reinterpret_cast<const int4 *>(d) + k);
    DictEnt e;
    e.delta = q.x;
    e.pad = 0;
    e.v = __hiloint2double(q.w, q.z);
    return e;
#else
    return d[k];
#endif
}

// Raw words of stored entry p (everything that comes from the matrix stream), and their decoding into
// (column, value).  Split in two so that a kernel can issue the stream loads of a whole chunk before the
// first dependent access.
struct SfRaw {
    int c;
    unsigned k;
    double v;
};

template <int FMT, bool STREAM, bool PRE = false>
SF_HD SfRaw sf_load(const MatView &A, int p)
{
    SfRaw r;
    r.c = 0;
    r.k = 0;
    r.v = 0.0;
    if (FMT == FMT_F64) {
        r.c = sf_ld<STREAM, PRE>(A.cols + p);
        r.v = sf_ld<STREAM, PRE>(A.vals + p);
    } else if (FMT == FMT_D16) {
        r.c = (int)sf_ld<STREAM, PRE>(A.dcol + p);
        r.v = sf_ld<STREAM, PRE>(A.vals + p);
    } else if (FMT == FMT_PK) {
        r.c = (int)sf_ld<STREAM, PRE>(A.dcol + p);
        r.k = sf_ld<STREAM, PRE>(A.vcode + p);
    } else if (FMT == FMT_DICT16) {
        r.k = sf_ld<STREAM, PRE>(reinterpret_cast<const uint16_t *>(A.code) + p);
    } else {
        r.k = sf_ld<STREAM, PRE>(reinterpret_cast<const uint8_t *>(A.code) + p);
    }
    return r;
}

// base: column base of the slice (SELL) or row (CSR-vector); row: the row of the entry
template <int FMT>
SF_HD void sf_decode(const MatView &A, const SfRaw &r, int p, int row, int base, int &c, double &v)
{
    if (FMT == FMT_F64) {
        c = r.c;
        v = r.v;
    } else if (FMT == FMT_D16 || FMT == FMT_PK) {
        c = ((unsigned)r.c == SF_ESCAPE) ? sf_ld<false>(A.cols + p) : base + r.c;
        v = (FMT == FMT_D16) ? r.v : sf_ld<false>(A.vdict + r.k);
    } else {
        const DictEnt e = sf_dict(A.dict, r.k);
        c = row + e.delta;
        v = e.v;
    }
}

// ---------------------------------------------------------------------------------------------------------
// host side: layouts and format choice
// ---------------------------------------------------------------------------------------------------------
struct SfCsr {                       // borrowed CSR arrays
    int n_rows = 0, n_cols = 0;
    const int *indptr = nullptr, *indices = nullptr;
};

// SELL-32 layout of a pattern: entry k of row r is stored at sp[r / 32].x + 32 k + r % 32.  Padding entries have
// value 0 and the first column of their row.
struct SfSellLayout {
    int n_rows = 0, n_cols = 0, n_slices = 0;
    int64_t n_stored = 0, nnz = 0;
    std::vector<int2> sp;                 // n_slices + 1
    std::vector<int> cols;                // n_stored
    std::vector<uint16_t> dcol;           // n_stored, empty when the 16-bit offsets are not worth it
    std::vector<int64_t> csr_to_sell;     // CSR entry -> stored position
};

inline void sf_sell_layout(const SfCsr &A, SfSellLayout &L)
{
    L.n_rows = A.n_rows;
    L.n_cols = A.n_cols;
    L.n_slices = (A.n_rows + 31) / 32;
    L.nnz = A.indptr[A.n_rows];
    L.sp.assign(L.n_slices + 1, make_int2(0, 0));
    for (int s = 0; s < L.n_slices; ++s) {
        int w = 0;
        for (int r = 32 * s; r < std::min(A.n_rows, 32 * s + 32); ++r) w = std::max(w, A.indptr[r + 1] - A.indptr[r]);
        L.sp[s + 1].x = L.sp[s].x + 32 * w;
    }
    L.n_stored = L.sp[L.n_slices].x;
    L.cols.assign((size_t)L.n_stored, 0);
    L.csr_to_sell.resize((size_t)L.nnz);
    int64_t escapes = 0;
    for (int s = 0; s < L.n_slices; ++s) {
        const int w = (L.sp[s + 1].x - L.sp[s].x) / 32;
        int base = INT32_MAX;
        // padding entries (value 0) point at a column the row gathers anyway: the row's first column, or the
        // slice's first column for an empty row -- close to the real columns, so that they never need an escape
        int slice_first = 0;
        for (int r = 32 * s; r < std::min(A.n_rows, 32 * s + 32); ++r)
            if (A.indptr[r + 1] > A.indptr[r]) {
                slice_first = A.indices[A.indptr[r]];
                break;
            }
        for (int lane = 0; lane < 32; ++lane) {
            const int r = 32 * s + lane;
            const int len = r < A.n_rows ? A.indptr[r + 1] - A.indptr[r] : 0;
            const int pad = len > 0 ? A.indices[A.indptr[r]] : slice_first;
            for (int k = 0; k < w; ++k) {
                const int64_t pos = (int64_t)L.sp[s].x + 32 * k + lane;
                if (k < len) {
                    L.cols[pos] = A.indices[A.indptr[r] + k];
                    L.csr_to_sell[A.indptr[r] + k] = pos;
                } else {
                    L.cols[pos] = pad;
                }
                base = std::min(base, L.cols[pos]);
            }
        }
        L.sp[s].y = (w > 0) ? base : 0;
    }
    // 16-bit offsets from the slice base; entries further away escape to the int32 array
    L.dcol.assign((size_t)L.n_stored, 0);
    for (int s = 0; s < L.n_slices; ++s)
        for (int64_t p = L.sp[s].x; p < L.sp[s + 1].x; ++p) {
            const int64_t d = (int64_t)L.cols[p] - L.sp[s].y;
            if (d >= (int64_t)SF_ESCAPE) {
                L.dcol[p] = (uint16_t)SF_ESCAPE;
                ++escapes;
            } else {
                L.dcol[p] = (uint16_t)d;
            }
        }
    if (escapes * 50 > L.n_stored) L.dcol.clear();      // more than 2 % escapes: every escape costs a second load
}

// open-addressing table: (delta, value bits) -> code, codes in order of first appearance
class SfPairTable {
public:
    explicit SfPairTable(int max_codes) : max_codes_(max_codes), slot_(1u << 18, -1) {}
    // returns the code, or -1 when the table would exceed max_codes
    int code(int delta, double v)
    {
        uint64_t bits;
        memcpy(&bits, &v, 8);
        uint64_t hsh = (bits ^ ((uint64_t)(uint32_t)delta * 0x9E3779B97F4A7C15ull)) * 0xD6E8FEB86659FD93ull;
        unsigned i = (unsigned)(hsh >> 46);      // 18 bits
        while (true) {
            const int s = slot_[i];
            if (s < 0) {
                if ((int)ents_.size() >= max_codes_) return -1;
                slot_[i] = (int)ents_.size();
                DictEnt e;
                e.delta = delta;
                e.pad = 0;
                e.v = v;
                ents_.push_back(e);
                return slot_[i];
            }
            uint64_t b2;
            memcpy(&b2, &ents_[s].v, 8);
            if (b2 == bits && ents_[s].delta == delta) return s;
            i = (i + 1) & ((1u << 18) - 1);
        }
    }
    const std::vector<DictEnt> &entries() const { return ents_; }

private:
    int max_codes_;
    std::vector<int> slot_;
    std::vector<DictEnt> ents_;
};

// Value arrays of one matrix on a SELL layout, in the most compact format its entries allow (at most max_fmt in
// the order F64 < D16 < PK < DICT16 < DICT8 of bytes saved; used by experiments and tests to force a format).
struct SfSellValues {
    int fmt = FMT_F64;
    std::vector<double> vals;        // F64 / D16
    std::vector<uint16_t> vcode;     // PK
    std::vector<double> vdict;       // PK
    std::vector<uint16_t> code16;    // DICT16
    std::vector<uint8_t> code8;      // DICT8
    std::vector<DictEnt> dict;       // DICT8 / DICT16 / STENCIL
    std::vector<int4> sp4;           // STENCIL
    std::vector<DictEnt> stab;       // STENCIL
    int ps_off = -2, ps_w = 0;       // STENCIL: most frequent stencil (offset in stab, width), -2: none short enough
    double uniform_fraction = 0.0;   // share of the slices whose 32 rows carry one stencil
    int64_t bytes_per_pass = 0;      // matrix stream bytes of one product
};

inline int sf_fmt_rank(int fmt) { return fmt; }   // enum order = order of compactness

inline void sf_sell_values(const SfSellLayout &L, const double *csr_values, int max_fmt, SfSellValues &V)
{
    const int64_t ns = L.n_stored;
    V = SfSellValues();
    // 1. (column - row, value) dictionary: square-ish matrices only (the padding entries point at their own row)
    if (max_fmt >= FMT_DICT16 && L.n_cols >= L.n_rows) {
        SfPairTable tab(SF_DICT_MAX);
        std::vector<uint16_t> code((size_t)ns);
        std::vector<double> sv((size_t)ns, 0.0);
        for (int64_t k = 0; k < L.nnz; ++k) sv[L.csr_to_sell[k]] = csr_values[k];
        bool ok = true;
        for (int s = 0; s < L.n_slices && ok; ++s) {
            for (int64_t p = L.sp[s].x; p < L.sp[s + 1].x; ++p) {
                const int row = 32 * s + (int)((p - L.sp[s].x) & 31);
                const int c = tab.code(L.cols[p] - row, sv[p]);
                if (c < 0) {
                    ok = false;
                    break;
                }
                code[p] = (uint16_t)c;
            }
        }
        if (ok) {
            V.dict = tab.entries();
            // slices whose 32 rows carry the same code sequence
            if (max_fmt >= FMT_STENCIL && L.n_slices > 0) {
                std::vector<int4> sp4((size_t)L.n_slices);
                std::vector<std::pair<std::vector<uint16_t>, int>> known;      // few distinct stencils: linear search
                std::vector<int64_t> known_count;
                std::vector<DictEnt> stab;
                int64_t n_uniform = 0, general_stored = 0;
                std::vector<uint16_t> seq;
                for (int s = 0; s < L.n_slices; ++s) {
                    const int start = L.sp[s].x, w = (L.sp[s + 1].x - start) / 32;
                    bool uni = 32 * s + 32 <= L.n_rows && w > 0;
                    for (int k = 0; k < w && uni; ++k)
                        for (int lane = 1; lane < 32; ++lane)
                            if (code[start + 32 * k + lane] != code[start + 32 * k]) {
                                uni = false;
                                break;
                            }
                    int off = -1;
                    if (uni) {
                        seq.resize(w);
                        for (int k = 0; k < w; ++k) seq[k] = code[start + 32 * k];
                        for (size_t q = 0; q < known.size(); ++q)
                            if (known[q].first == seq) {
                                off = known[q].second;
                                known_count[q]++;
                                break;
                            }
                        if (off < 0 && known.size() < 4096) {
                            off = (int)stab.size();
                            for (int k = 0; k < w; ++k) stab.push_back(V.dict[seq[k]]);
                            known.emplace_back(seq, off);
                            known_count.push_back(1);
                        }
                    }
                    if (off >= 0) ++n_uniform;
                    else general_stored += 32 * w;
                    sp4[s] = make_int4(start, w, off, 0);
                }
                V.uniform_fraction = (double)n_uniform / L.n_slices;
                if (V.uniform_fraction >= 0.5) {
                    V.fmt = FMT_STENCIL;
                    int64_t best = 0;
                    for (size_t q = 0; q < known.size(); ++q)
                        if ((int)known[q].first.size() <= SF_PS && known_count[q] > best) {
                            best = known_count[q];
                            V.ps_off = known[q].second;
                            V.ps_w = (int)known[q].first.size();
                        }
                    V.sp4.swap(sp4);
                    V.stab.swap(stab);
                    V.code16.swap(code);
                    V.bytes_per_pass = 2 * general_stored + 16ll * L.n_slices;
                    return;
                }
            }
            if (V.dict.size() <= 256 && max_fmt >= FMT_DICT8) {
                V.fmt = FMT_DICT8;
                V.code8.resize((size_t)ns);
                for (int64_t p = 0; p < ns; ++p) V.code8[p] = (uint8_t)code[p];
                V.bytes_per_pass = ns;
            } else {
                V.fmt = FMT_DICT16;
                V.code16.swap(code);
                V.bytes_per_pass = 2 * ns;
            }
            return;
        }
    }
    // 2. value dictionary next to 16-bit column offsets
    if (max_fmt >= FMT_PK && !L.dcol.empty()) {
        SfPairTable tab(SF_DICT_MAX);
        std::vector<uint16_t> code((size_t)ns, 0);
        bool ok = tab.code(0, 0.0) == 0;          // padding entries: code 0 = 0.0
        for (int64_t k = 0; k < L.nnz && ok; ++k) {
            const int c = tab.code(0, csr_values[k]);
            if (c < 0) ok = false;
            else code[L.csr_to_sell[k]] = (uint16_t)c;
        }
        if (ok) {
            V.fmt = FMT_PK;
            V.vcode.swap(code);
            V.vdict.resize(tab.entries().size());
            for (size_t i = 0; i < V.vdict.size(); ++i) V.vdict[i] = tab.entries()[i].v;
            V.bytes_per_pass = 4 * ns;
            return;
        }
    }
    V.vals.assign((size_t)ns, 0.0);
    for (int64_t k = 0; k < L.nnz; ++k) V.vals[L.csr_to_sell[k]] = csr_values[k];
    if (max_fmt >= FMT_D16 && !L.dcol.empty()) {
        V.fmt = FMT_D16;
        V.bytes_per_pass = 10 * ns;
    } else {
        V.fmt = FMT_F64;
        V.bytes_per_pass = 12 * ns;
    }
}

// CSR-vector layout (several lanes per row): per-row column base, 16-bit offsets, optional value dictionary
struct SfCsrvData {
    int fmt = FMT_F64;
    std::vector<int> rbase;
    std::vector<uint16_t> dcol, vcode;
    std::vector<double> vdict;
    int64_t bytes_per_pass = 0;
};

inline void sf_csrv_data(const SfCsr &A, const double *values, int max_fmt, SfCsrvData &D)
{
    D = SfCsrvData();
    const int64_t nnz = A.indptr[A.n_rows];
    D.bytes_per_pass = 12 * nnz;
    if (max_fmt < FMT_D16 || nnz == 0) return;
    D.rbase.assign(A.n_rows, 0);
    D.dcol.assign((size_t)nnz, 0);
    int64_t escapes = 0;
    for (int r = 0; r < A.n_rows; ++r) {
        int base = INT32_MAX;
        for (int k = A.indptr[r]; k < A.indptr[r + 1]; ++k) base = std::min(base, A.indices[k]);
        D.rbase[r] = base == INT32_MAX ? 0 : base;
        for (int k = A.indptr[r]; k < A.indptr[r + 1]; ++k) {
            const int64_t d = (int64_t)A.indices[k] - base;
            if (d >= (int64_t)SF_ESCAPE) {
                D.dcol[k] = (uint16_t)SF_ESCAPE;
                ++escapes;
            } else {
                D.dcol[k] = (uint16_t)d;
            }
        }
    }
    if (escapes * 50 > nnz) {
        D.rbase.clear();
        D.dcol.clear();
        return;
    }
    D.fmt = FMT_D16;
    D.bytes_per_pass = 10 * nnz;
    if (max_fmt < FMT_PK) return;
    SfPairTable tab(SF_DICT_MAX);
    std::vector<uint16_t> code((size_t)nnz);
    for (int64_t k = 0; k < nnz; ++k) {
        const int c = tab.code(0, values[k]);
        if (c < 0) return;
        code[k] = (uint16_t)c;
    }
    D.fmt = FMT_PK;
    D.vcode.swap(code);
    D.vdict.resize(tab.entries().size());
    for (size_t i = 0; i < D.vdict.size(); ++i) D.vdict[i] = tab.entries()[i].v;
    D.bytes_per_pass = 4 * nnz;
}
