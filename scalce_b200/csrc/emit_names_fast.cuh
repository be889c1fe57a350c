// emit_names_fast.cuh - stream 0 ([len:u8][name] per read, names.cpp:48-62) with word stores into the staging buffer
// (opt-in: SCB_EMIT_NAMES_V2=1; written without GPU access after the A/B runs of round 1, not yet measured).
//
// emit_names_st_k spends 14 shared-memory byte stores on a 13-byte name and is throttled by the shared-memory instruction
// queue (ncu: mio_throttle 36, short_scoreboard 21 per issued instruction, 14 % issue active). Here the length byte is
// shifted in front of the first 15 name bytes in registers and the record leaves through store_bytes16 (emit_name.h:
// <= 3 + 4 + 3 stores per 16 bytes; every (alignment, length) pair checked on the CPU). Same tile layout, same flush, same
// bytes written as emit_names_st_k.
#pragma once
#include "emit2.cuh"
#include "emit_name.h"

namespace scb {

__global__ void __launch_bounds__(256) emit_names_fast_k(EmitMParams e) {
    __shared__ __align__(16) uint8_t sb[kNamesCap + 32];
    const int64_t p0 = (int64_t)blockIdx.x * 256, p1 = (p0 + 256 < e.n) ? p0 + 256 : e.n;
    const uint64_t g0 = e.offN[p0];
    const int64_t len64 = (int64_t)(e.offN[p1] - g0);
    const bool staged = len64 <= kNamesCap;
    const int64_t p = p0 + threadIdx.x;
    if (p < p1) {
        const uint64_t m = e.ms[p];
        const int64_t a = meta_name_off(m);
        const int nl = meta_namelen(m);
        const uint64_t o = e.offN[p];
        if (staged) {
            uint8_t *d = sb + (int)(g0 & 15) + (int)(o - g0);
            // first chunk: the length byte followed by up to 15 name bytes
            const int n0 = nl < 15 ? nl : 15;
            uint4 v = make_uint4(0, 0, 0, 0);
            if (n0 > 0) v = load16_unaligned(e.names + a, n0);
            store_bytes16(d, (v.x << 8) | (uint32_t)nl, __funnelshift_l(v.x, v.y, 8), __funnelshift_l(v.y, v.z, 8), __funnelshift_l(v.z, v.w, 8), n0 + 1);
            for (int k = 15; k < nl; k += 16) {       // names longer than 15 bytes: plain 16-byte chunks
                const int nbv = nl - k < 16 ? nl - k : 16;
                const uint4 w = load16_unaligned(e.names + a + k, nbv);
                store_bytes16(d + 1 + k, w.x, w.y, w.z, w.w, nbv);
            }
        } else {
            uint8_t *d = e.oN + o;
            d[0] = (uint8_t)nl;
            for (int k = 0; k < nl; k++) d[1 + k] = (uint8_t)ldg_g64(e.names + a + k);
        }
    }
    __syncthreads();
    if (staged) flush_staged(e.oN, g0, (int)len64, sb);
}

}  // namespace scb
