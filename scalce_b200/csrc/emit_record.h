// emit_record.h - one read's stream-1 record (rotated 2-bit read + end marker), host/device.
//
// output_read(read, dest, end - level, level), reads.cpp:432-461: the bases after the core, [end, L), then the bases
// before it, [0, end - level), 4 per byte MSB first, last byte zero padded; then the end marker as the low sz_meta
// bytes of int16 end (reads.cpp:128-130). Source: the read's 2-bit packed row (16 bases per u32, MSB first, zero
// filled past L, at least 2 readable words after the row). Plain C++ so that the same code is compiled into
// the kernel (emit_reads_fast.cuh) and into the CPU test that checks it base by base (tests/test_host_cpu.py).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define SCB_HD __host__ __device__ __forceinline__
#else
#define SCB_HD inline
#endif

namespace scb {

// 16 bases (32 bits, MSB first) starting at base offset b0 >= 0 of a packed row
SCB_HD uint32_t er_bits32(const uint32_t *row, int b0) {
    const int k = b0 >> 4, sh = (b0 & 15) * 2;
    const uint32_t hi = row[k], lo = row[k + 1];
#if defined(__CUDA_ARCH__)
    return __funnelshift_l(lo, hi, sh);
#else
    return sh ? ((hi << sh) | (lo >> (32 - sh))) : hi;
#endif
}

// mask that keeps the first nb (0..16) bases of a 16-base word
SCB_HD uint32_t er_keep(int nb) { return nb >= 16 ? 0xffffffffu : ~(0xffffffffu >> (2 * nb)); }

// Writes the record of one read at d (any byte alignment); returns its size. Every output word is built from both
// sources without branches (lanes of a warp hold reads with different core positions), full words leave as four
// unconditional byte stores; only the last partial word and the end marker are conditional.
SCB_HD int emit_record(const uint32_t *row, int L1, int lv, int end, int sz_meta, uint8_t *d) {
    const int tail = L1 - end, total = L1 - lv;
    const int nbytes = (total >> 2) + ((total & 3) != 0);
    const int wfull = nbytes >> 2;
    int w = 0;
    for (;; w++) {
        const int j0 = 16 * w;
        int a = tail - j0;                       // bases of this word that come from the part after the core
        a = a < 0 ? 0 : (a > 16 ? 16 : a);
        const int hs = a < 16 ? j0 + a - tail : 0;   // source offset of the first base taken from the part before the core (>= 0)
        const int s1 = a > 0 ? end + j0 : 0;     // < L1 whenever it is used; never reads past the row's pad words
        uint32_t v = (er_bits32(row, s1) & er_keep(a)) | (a < 16 ? (er_bits32(row, hs) >> (2 * a)) : 0u);   // both loads are always in range
        const int nv = total - j0;               // valid bases from j0 on
        if (nv < 16) v &= er_keep(nv < 0 ? 0 : nv);
        if (w < wfull) {
            d[4 * w + 0] = (uint8_t)(v >> 24); d[4 * w + 1] = (uint8_t)(v >> 16); d[4 * w + 2] = (uint8_t)(v >> 8); d[4 * w + 3] = (uint8_t)v;
        } else {
            const int rem = nbytes - 4 * w;      // 0..3 bytes of the read left
            for (int k = 0; k < rem; k++) d[4 * w + k] = (uint8_t)(v >> (24 - 8 * k));
            break;
        }
    }
    d[nbytes] = (uint8_t)end;
    if (sz_meta > 1) d[nbytes + 1] = (uint8_t)((uint32_t)end >> 8);
    return nbytes + sz_meta;
}

}  // namespace scb
