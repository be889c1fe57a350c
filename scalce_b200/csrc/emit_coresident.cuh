// emit_coresident.cuh - the three output kernels as CO-RESIDENT persistent grids (opt-in: SCB_EMIT_CORESIDENT=1;
// written without GPU access, not yet measured).
//
// ncu (profiles/r01_ncu_top_kernels.txt, gpurun_out/r01_top.raw.csv): the quality-row gather is DRAM bound (4.45 TB/s of
// traffic, issue 62 %), the packed-read writer is issue bound (83 % issue active, 3.2 TB/s), the name writer is latency
// bound (14 % issue active, 2.3 TB/s). Launched as full grids on three streams they still run one after the other,
// because each grid alone fills the machine. Here each kernel is a grid-stride loop over its tiles with a grid of a few
// CTAs per SM (rows 3 at <= 40 registers, reads 2, names 2: 7 x 256 threads and 63.5k of the 64k registers of an SM), so all three are resident together and the
// SM schedulers interleave DRAM-bound, issue-bound and latency-bound warps. Floor for the three together at the row
// gather's DRAM efficiency: 30.3 GB / 4.45 TB/s = 6.8 ms against 8.4 ms back to back.
// The tile bodies are copies of gather_rows16_k / emit_names_st_k / emit_reads_st_k (emit2.cuh) with blockIdx.x
// replaced by the loop variable; results are identical.
#pragma once
#include "emit2.cuh"

namespace scb {

__global__ void __launch_bounds__(256, 6) gather_rows16_loop_k(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst,
                                                            const uint32_t *__restrict__ perm, int64_t n, int L, int n_blk) {
    const int64_t total = n * (int64_t)L;
    for (int blk = blockIdx.x; blk < n_blk; blk += gridDim.x) {
        const int64_t blk0 = (int64_t)blk * (256 * kGatherChunks * 16);
        const int64_t pblk = blk0 / L;
        const uint32_t rblk = (uint32_t)(blk0 - pblk * L);
        const uint8_t *a0[kGatherChunks], *a1[kGatherChunks];
        int n0[kGatherChunks];
#pragma unroll
        for (int q = 0; q < kGatherChunks; q++) {
            const uint32_t lo = (uint32_t)(q * 256 + threadIdx.x) << 4;
            a0[q] = a1[q] = nullptr; n0[q] = 16;
            if (blk0 + lo < total) {
                const uint32_t x = rblk + lo, dp = x / (uint32_t)L;
                const int64_t p0 = pblk + dp;
                const int r0 = (int)(x - dp * (uint32_t)L);
                n0[q] = min(16, L - r0);
                a0[q] = src + (int64_t)perm[p0] * L + r0;
                if (n0[q] < 16 && p0 + 1 < n) a1[q] = src + (int64_t)perm[p0 + 1] * L;
            }
        }
        uint4 v[kGatherChunks];
#pragma unroll
        for (int q = 0; q < kGatherChunks; q++)
            if (a0[q]) {
                v[q] = load16_unaligned(a0[q], n0[q]);
                if (a1[q]) v[q] = splice16(v[q], load16_unaligned(a1[q], 16 - n0[q]), n0[q]);
            }
#pragma unroll
        for (int q = 0; q < kGatherChunks; q++) {
            const int64_t o = blk0 + ((int64_t)(q * 256 + threadIdx.x) << 4);
            if (!a0[q]) continue;
            if (o + 16 <= total) *(uint4 *)(dst + o) = v[q];
            else {
                const uint32_t w[4] = {v[q].x, v[q].y, v[q].z, v[q].w};
                for (int k = 0; k < (int)(total - o); k++) dst[o + k] = (uint8_t)(w[k >> 2] >> (8 * (k & 3)));
            }
        }
    }
}

__global__ void __launch_bounds__(256) emit_names_loop_k(EmitMParams e, int64_t n_blk) {
    __shared__ __align__(16) uint8_t sb[kNamesCap + 32];
    for (int64_t blk = blockIdx.x; blk < n_blk; blk += gridDim.x) {
        const int64_t p0 = blk * 256, p1 = (p0 + 256 < e.n) ? p0 + 256 : e.n;
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
                d[0] = (uint8_t)nl;
                for (int k = 0; k < nl; k += 16) {
                    const int nbv = nl - k < 16 ? nl - k : 16;
                    const uint4 v = load16_unaligned(e.names + a + k, nbv);
                    const uint32_t wv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                    for (int j = 0; j < 16; j++)
                        if (j < nbv) d[1 + k + j] = (uint8_t)(wv[j >> 2] >> (8 * (j & 3)));
                }
            } else {
                uint8_t *d = e.oN + o;
                d[0] = (uint8_t)nl;
                for (int k = 0; k < nl; k++) d[1 + k] = (uint8_t)ldg_g64(e.names + a + k);
            }
        }
        __syncthreads();
        if (staged) flush_staged(e.oN, g0, (int)len64, sb);
        __syncthreads();   // the staging buffer is reused by the next tile
    }
}

__global__ void __launch_bounds__(256) emit_reads_loop_k(EmitMParams e, int RPB, uint32_t NW, uint32_t inv_pws, int recmax, int64_t n_blk) {
    extern __shared__ __align__(16) uint8_t sbd[];
    const int PWs = (e.PW + kEmitRowPad) | 1;
    uint32_t *s_rows = (uint32_t *)(sbd + (((size_t)RPB * recmax + 48 + 15) & ~(size_t)15));
    for (int64_t blk = blockIdx.x; blk < n_blk; blk += gridDim.x) {
        const int64_t p0 = blk * RPB, p1 = (p0 + RPB < e.n) ? p0 + RPB : e.n;
        const int np = (int)(p1 - p0);
        const uint64_t g0 = e.offR[p0];
        const int len = (int)(e.offR[p1] - g0);
        uint8_t *sb = sbd + (int)(g0 & 15);
        {
            const uint32_t items = (uint32_t)np * (uint32_t)PWs;
            for (uint32_t t = threadIdx.x; t < items; t += 256) {
                const uint32_t pl = __umulhi(t, inv_pws), k = t - pl * (uint32_t)PWs;
                s_rows[t] = k < (uint32_t)e.PW ? ldg_g64(e.packed + (int64_t)e.perm[p0 + pl] * e.PW + k) : 0u;
            }
        }
        __syncthreads();
        for (int pl = threadIdx.x; pl < np; pl += 256) {
            const uint64_t m = e.ms[p0 + pl];
            const int lv = meta_lvl(m), end = meta_end(m);
            const int tail = e.L1 - end, total = e.L1 - lv;
            const int nbytes = sz_read(total);
            const int recsz = nbytes + e.sz_meta;
            const uint32_t *row = s_rows + (size_t)pl * PWs;
            uint8_t *d = sb + (int)(e.offR[p0 + pl] - g0);
            for (int w = 0; 4 * w < recsz; w++) {
                uint32_t v = 0;
                const int j0 = 16 * w;
                if (4 * w < nbytes) {
                    int a = tail - j0; a = a < 0 ? 0 : (a > 16 ? 16 : a);
                    int nv = total - j0; nv = nv > 16 ? 16 : nv;
                    if (a > 0) v = spk_bits32(row, end + j0) & (a == 16 ? 0xffffffffu : ~(0xffffffffu >> (2 * a)));
                    if (a < 16) v |= spk_bits32(row, j0 + a - tail) >> (2 * a);
                    if (nv < 16) v &= ~(0xffffffffu >> (2 * nv));
                }
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const int b = 4 * w + k;
                    if (b < nbytes) d[b] = (uint8_t)(v >> (24 - 8 * k));
                    else if (b < recsz) d[b] = (uint8_t)((uint32_t)end >> (8 * (b - nbytes)));
                }
            }
        }
        __syncthreads();
        flush_staged(e.oR, g0, len, sbd);
        __syncthreads();   // staging buffers are reused by the next tile
    }
}

}  // namespace scb
