// emit_name.h - byte-granular stores with as few store instructions as possible, host/device.
//
// The stream-0 writer (names.cpp:48-62: one length byte + the name per read) assembles the records of a CTA in shared
// memory at arbitrary byte offsets; emit_names_st_k does that with one byte store per byte (14 per 13-byte name) and is
// bound by the shared-memory instruction queue (ncu: mio_throttle + short_scoreboard). store_bytes16 writes up to 16
// bytes held in four little-endian words with <= 3 byte stores up to the next word boundary, <= 4 word stores, <= 3 byte
// stores at the end. Plain C++: compiled into emit_names_fast_k and into the CPU test that checks every (alignment,
// length) pair (tests/test_host_cpu.py).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define SCB_HD2 __host__ __device__ __forceinline__
#else
#define SCB_HD2 inline
#endif

namespace scb {

// (lo, hi) as one 64-bit value shifted right by sh bits (0..31), low word
SCB_HD2 uint32_t en_shr(uint32_t lo, uint32_t hi, int sh) {
#if defined(__CUDA_ARCH__)
    return __funnelshift_r(lo, hi, sh);
#else
    return sh ? (lo >> sh) | (hi << (32 - sh)) : lo;
#endif
}

// d[0..n) = the first n (0..16) bytes of the little-endian words w0..w3; d may have any alignment
SCB_HD2 void store_bytes16(uint8_t *d, uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3, int n) {
    const int s = (int)((uintptr_t)d & 3);
    int head = s ? 4 - s : 0;                       // bytes up to the next word boundary
    head = head < n ? head : n;
    if (head > 0) d[0] = (uint8_t)w0;
    if (head > 1) d[1] = (uint8_t)(w0 >> 8);
    if (head > 2) d[2] = (uint8_t)(w0 >> 16);
    const int sh = head * 8;
    const uint32_t x0 = en_shr(w0, w1, sh), x1 = en_shr(w1, w2, sh), x2 = en_shr(w2, w3, sh), x3 = en_shr(w3, 0u, sh);
    const int rem = n - head, nw = rem >> 2, tail = rem & 3;
    uint32_t *dw = (uint32_t *)(d + head);          // word aligned whenever a word is stored through it
    if (nw > 0) dw[0] = x0;
    if (nw > 1) dw[1] = x1;
    if (nw > 2) dw[2] = x2;
    if (nw > 3) dw[3] = x3;
    const uint32_t xt = nw == 0 ? x0 : (nw == 1 ? x1 : (nw == 2 ? x2 : x3));
    uint8_t *dt = d + head + 4 * nw;
    if (tail > 0) dt[0] = (uint8_t)xt;
    if (tail > 1) dt[1] = (uint8_t)(xt >> 8);
    if (tail > 2) dt[2] = (uint8_t)(xt >> 16);
}

}  // namespace scb
