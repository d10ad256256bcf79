// Host-only helpers of the TMA-fed KKT-apply variants (kkt_apply.cu, CTL_KKT_TMA=3|4): the record stream.
// No CUDA in here, so tests/native/ can include it directly.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <vector>

// Per row block of TR rows one 16-byte aligned record, in exactly the layout the kernels read from shared memory:
//   int      ptr[TR + 1]      entry offsets relative to the block (rows past the end repeat the last one), padded to 16 B
//   double2  mk[cnt + 1]      (m, k) value pairs, entry cnt = zero sentinel
//   double   kt[cnt + 1]      K^T values, only when kt != nullptr (non-symmetric K), padded to 16 B
//   unsigned off[cnt + 1]     byte offset of the gathered X row inside the tile (slot * row_bytes), padded to 16 B
// rec_off[b] = start of block b's record in units of 16 bytes (n_blocks + 1 entries).  Returns the longest record in bytes.
inline int kkt_build_block_records(int TR, int n_rows, const int *indptr, const uint8_t *slot, const double *m, const double *k,
                                   const double *kt, size_t row_bytes, std::vector<uint8_t> &rec, std::vector<int> &rec_off)
{
    auto up16 = [](size_t v) { return (v + 15) & ~(size_t)15; };
    const int nblk = (n_rows + TR - 1) / TR;
    const size_t hdr = up16((size_t)(TR + 1) * 4);
    rec.clear();
    rec_off.assign(nblk + 1, 0);
    int rec_max = 0;
    for (int b = 0; b < nblk; ++b) {
        const int r0 = b * TR, r1 = std::min(n_rows, r0 + TR);
        const int kb = indptr[r0], cnt = indptr[r1] - kb;
        const size_t o_mk = hdr, o_kt = o_mk + (size_t)(cnt + 1) * 16;
        const size_t o_off = kt ? o_kt + up16((size_t)(cnt + 1) * 8) : o_kt;
        const size_t bytes = o_off + up16((size_t)(cnt + 1) * 4);
        const size_t base = rec.size();
        rec.resize(base + bytes, 0);
        uint8_t *p = rec.data() + base;
        for (int i = 0; i <= TR; ++i) {
            const int v = indptr[std::min(r0 + i, r1)] - kb;
            std::memcpy(p + (size_t)i * 4, &v, 4);
        }
        for (int e = 0; e < cnt; ++e) {
            const double mk[2] = {m[kb + e], k[kb + e]};
            std::memcpy(p + o_mk + (size_t)e * 16, mk, 16);
            if (kt) std::memcpy(p + o_kt + (size_t)e * 8, &kt[kb + e], 8);
            const unsigned off = (unsigned)slot[kb + e] * (unsigned)row_bytes;
            std::memcpy(p + o_off + (size_t)e * 4, &off, 4);
        }
        rec_off[b + 1] = (int)((base + bytes) / 16);
        rec_max = std::max(rec_max, (int)bytes);
    }
    return rec_max;
}
