// scan_big.cuh - core scan for automata that do not fit shared memory (core sets of production size: the reference
// sizes patterns[] for 5-10 M cores, reads.cpp:336, 385; 1 M cores of 12 bases are ~3 M states, 46 MB of transitions).
//
// Same warp-autonomous pipeline as scan_smem2.cuh (A pack: 16-byte cp.async tiles of 32 reads, SWAR 2-bit packing;
// B walk: one lane per read over the packed words; C pick + emit in one pass over the hits), but the completed DFA
// (reads.cpp:298-315) stays in global memory as one u32 per (state, base) and is served by L2: on random reads the walk
// sits at depth >= 9 almost always, so every base is one random 4-byte read of a table far larger than L1 - the kernel
// is bound by L2 sector throughput, and the warps per SM (as many as the per-warp staging allows) are there to keep
// that many lookups in flight. States are renumbered so that "some core ends here" is state >= H0 (one compare per
// base); hit_info[state - H0] = bucket rank | core length << 24 (ranks < 2^24).
// Reads with more hits than the per-lane queue holds take an exact slower path that dedupes against the candidates
// already written (O(hits x candidates), not the O(L) re-walk per hit of the shared-memory kernels).
#pragma once
#include "common.cuh"
#include "pipeline.cuh"
#include "scan_smem.cuh"

namespace scb {

constexpr int kHitQB = 48;       // queued hit states per read (u32 each)
constexpr int kHitQBGuard = 16;  // a 16-base word is only walked on the fast path if it cannot overflow the queue

struct ScanBigParams {
    const uint8_t *seq; int64_t n; int L;
    const uint32_t *trans;        // [ns * 4] next state
    const uint32_t *hit_info;     // [n_hit]  rank | level << 24
    uint32_t H0;
    uint8_t *lvl; uint16_t *ncand; uint64_t *cand_off; uint32_t *cand_rank; uint16_t *cand_pos;
    unsigned long long *cand_total; uint64_t cand_cap;
    int64_t n_tiles;
    uint32_t *packed; int PW;
    uint32_t inv_pw;
    int pitch;
};

__host__ __device__ inline size_t scan_big_warp_bytes(int L, int PW) {
    const size_t tile = (size_t)32 * L + 32;
    const size_t pk = (size_t)32 * scan_smem_pitch(PW) * 4, q = (size_t)32 * kHitQB * 4, hm = (((size_t)32 * PW * 2) + 15) & ~(size_t)15;
    return ((tile + pk + q + hm) + 15) & ~(size_t)15;
}

// L2 residency is what this kernel lives on: the transition table (tens of MB) is re-read 150 times per read while 7.5 GB of
// ASCII, 2 GB of packed rows and the candidate lists stream through the same L2 once. Without hints the streams evict the
// table (first B200 run: 48 G lookups/s = ~3 TB/s of 64-byte DRAM fetches, i.e. the table was served by HBM, not by L2).
// So: table loads carry an evict_last policy, the ASCII tiles an evict_first policy, and the kernel's stores are streaming
// (st.global.cs).
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint32_t ldg_nc_u32(const uint32_t *p, uint64_t pol) {
    uint32_t v;
    asm volatile("ld.global.nc.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ void cp_async16_hint(void *smem_dst, const void *gsrc, uint64_t pol) {
    uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;\n" ::"r"(d), "l"(gsrc), "l"(pol));
}
// stage_warp_tile (scan_smem.cuh) with the evict_first policy on the tile's 16-byte copies
__device__ __forceinline__ void stage_warp_tile_stream(const uint8_t *seq, int64_t n, int L, int64_t tile, uint8_t *buf, uint64_t pol) {
    const int64_t row0 = tile * 32;
    int64_t rows = n - row0;
    if (rows > 32) rows = 32;
    if (rows <= 0) return;
    const int bytes = (int)rows * L;
    const uint8_t *src = seq + row0 * L;
    const int n16 = bytes >> 4;
    if (pol) { for (int k = lane_id(); k < n16; k += 32) cp_async16_hint(buf + (k << 4), src + ((int64_t)k << 4), pol); }
    else { for (int k = lane_id(); k < n16; k += 32) cp_async16(buf + (k << 4), src + ((int64_t)k << 4)); }
    for (int k = (n16 << 4) + lane_id(); k < bytes; k += 32) buf[k] = src[k];
}
template <int HINTS>
__device__ __forceinline__ uint32_t big_ld(const uint32_t *p, uint64_t pol) {
    if (HINTS & 1) return ldg_nc_u32(p, pol);
    return __ldg(p);
}
template <int HINTS, typename T>
__device__ __forceinline__ void big_st(T *p, T v) {
    if (HINTS & 4) __stcs(p, v); else *p = v;
}
__device__ __forceinline__ void sts_u32(uint32_t saddr, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;\n" ::"r"(saddr), "r"(v) : "memory"); }

// HINTS bit 0: evict_last policy on the table loads; bit 1: evict_first policy on the ASCII tile copies; bit 2: streaming stores
template <int HINTS>
__global__ void __launch_bounds__(1024, 1) scan_big_k(ScanBigParams p) {
    extern __shared__ __align__(16) uint8_t sm[];
    // layout per warp: ASCII tile (+32) | packed tile | hit queues (u32) | hit masks
    const int L = p.L, PW = p.PW, pitch = p.pitch;
    const int w = threadIdx.x >> 5, W = blockDim.x >> 5, l = lane_id();
    uint8_t *s_tile = sm + (size_t)w * scan_big_warp_bytes(L, PW);
    uint32_t *s_pk = (uint32_t *)(s_tile + (size_t)32 * L + 32);
    uint32_t *s_q = s_pk + (size_t)32 * pitch;
    uint16_t *s_hm = (uint16_t *)(s_q + (size_t)32 * kHitQB);
    const uint32_t H0 = p.H0;
    const int full = L >> 4, tail = L & 15;
    const uint32_t *row = s_pk + (size_t)l * pitch;
    uint32_t *q = s_q + (size_t)l * kHitQB;
    uint16_t *hm = s_hm + (size_t)l * PW;
    const uint32_t q0 = (uint32_t)__cvta_generic_to_shared(q);
    const uint32_t *__restrict__ tr = p.trans;
    uint64_t c_cur = 0, c_end = 0;

    const uint64_t pol_keep = (HINTS & 1) ? l2_policy_evict_last() : 0ull, pol_stream = (HINTS & 2) ? l2_policy_evict_first() : 0ull;

    const int64_t stride = (int64_t)gridDim.x * W;
    int64_t tile = (int64_t)w * gridDim.x + blockIdx.x;
    if (tile < p.n_tiles) stage_warp_tile_stream(p.seq, p.n, L, tile, s_tile, pol_stream);
    cp_async_commit();
    for (; tile < p.n_tiles; tile += stride) {
        cp_async_wait<0>();
        __syncwarp();
        int rows = (int)((p.n - tile * 32 < 32) ? (p.n - tile * 32) : 32);
        // ---- A: pack (as scan_smem2_k) ------------------------------------------------------------------
        {
            const uint32_t nwords = (uint32_t)rows * (uint32_t)PW;
            const uint32_t *tw = (const uint32_t *)s_tile;
            uint32_t *gp = p.packed + tile * (int64_t)32 * PW;
            for (uint32_t t = l; t < nwords; t += 32) {
                const uint32_t r = PW == 1 ? t : __umulhi(t, p.inv_pw), k = t - r * (uint32_t)PW;
                const uint32_t b = r * (uint32_t)L + 16u * k;
                const uint32_t *a = tw + (b >> 2);
                const uint32_t sh = (b & 3u) * 8u;
                const uint32_t x0 = a[0], x1 = a[1], x2 = a[2], x3 = a[3], x4 = a[4];
                const uint32_t y0 = __funnelshift_r(x0, x1, sh), y1 = __funnelshift_r(x1, x2, sh), y2 = __funnelshift_r(x2, x3, sh),
                               y3 = __funnelshift_r(x3, x4, sh);
                uint32_t bad = 0;
                uint32_t wv = (pack4(y0, bad) << 24) | (pack4(y1, bad) << 16) | (pack4(y2, bad) << 8) | pack4(y3, bad);
                if (bad) wv = (pack4_masked(y0) << 24) | (pack4_masked(y1) << 16) | (pack4_masked(y2) << 8) | pack4_masked(y3);
                const int nv = L - 16 * (int)k;
                if (nv < 16) wv &= ~(0xffffffffu >> (2 * nv));
                s_pk[r * (uint32_t)pitch + k] = wv;
                big_st<HINTS>(gp + t, wv);
            }
        }
        __syncwarp();
        {
            const int64_t nxt = tile + stride;
            if (nxt < p.n_tiles) stage_warp_tile_stream(p.seq, p.n, L, nxt, s_tile, pol_stream);
            cp_async_commit();
        }
        // ---- B: walk, one L2 lookup per base ---------------------------------------------------------------
        const int64_t i = tile * 32 + l;
        const bool live = l < rows;
        int nh = 0, best = 0, cnt = 0, first_kept = 0;
        bool slow = false;
        if (live) {
            uint32_t e = 0;
            uint32_t qp = q0;
            for (int k = 0; k < full && !slow; k++) {
                if (qp - q0 > 4u * (kHitQB - kHitQBGuard)) { slow = true; break; }
                const uint32_t wv = row[k];
                uint32_t m = 0;
#pragma unroll
                for (int j = 0; j < 16; j++) {
                    const uint32_t c = (wv >> (30 - 2 * j)) & 3u;
                    e = big_ld<HINTS>(tr + ((size_t)e * 4u + c), pol_keep);
                    if (e >= H0) { m |= 1u << j; sts_u32(qp, e); qp += 4; }
                }
                hm[k] = (uint16_t)m;
            }
            if (tail && !slow) {
                if (qp - q0 > 4u * (kHitQB - kHitQBGuard)) slow = true;
                else {
                    const uint32_t wv = row[full];
                    uint32_t m = 0;
#pragma unroll
                    for (int j = 0; j < 15; j++) {
                        if (j >= tail) break;
                        const uint32_t c = (wv >> (30 - 2 * j)) & 3u;
                        e = big_ld<HINTS>(tr + ((size_t)e * 4u + c), pol_keep);
                        if (e >= H0) { m |= 1u << j; sts_u32(qp, e); qp += 4; }
                    }
                    hm[full] = (uint16_t)m;
                }
            }
            nh = (int)((qp - q0) >> 2);
        }
        // ---- candidate space before the pick, by hit count -----------------------------------------------------
        const uint32_t v = live ? (slow ? (uint32_t)L : (uint32_t)nh) : 0u;
        const uint32_t inc = warp_incl_scan(v);
        const uint32_t wtot = __shfl_sync(0xffffffffu, inc, 31);
        if (c_cur + wtot > c_end) {
            unsigned long long take = wtot > (uint32_t)kCandChunk ? wtot : (uint32_t)kCandChunk, got = 0;
            if (l == 0) got = atomicAdd(p.cand_total, take);
            got = __shfl_sync(0xffffffffu, got, 0);
            c_cur = got; c_end = got + take;
        }
        const uint64_t o = c_cur + (inc - v);
        c_cur += wtot;
        // ---- C: pick + emit (aho_search, reads.cpp:413-429, minus the running populations) ------------------------
        if (live) {
            const bool room = o + (uint64_t)v <= p.cand_cap;
            if (!slow) {
                int k = 0;
                uint32_t m = nh > 0 ? (uint32_t)hm[0] : 0u;
                for (int j = 0; j < nh; j++) {
                    while (m == 0) { k++; m = hm[k]; }
                    const int bpos = __ffs(m) - 1;
                    m &= m - 1;
                    const uint32_t info = big_ld<HINTS>(p.hit_info + (q[j] - H0), pol_keep);
                    const uint32_t r = info & 0x00ffffffu;
                    const int lv = (int)(info >> 24);
                    if (lv > best) { best = lv; cnt = 0; first_kept = j; }
                    bool drop = lv != best;
                    for (int kk = first_kept; kk < j && !drop; kk++) drop = (q[kk] == r);
                    q[j] = drop ? 0xffffffffu : r;
                    if (!drop) {
                        if (room) { big_st<HINTS>(p.cand_rank + o + cnt, r); big_st<HINTS>(p.cand_pos + o + cnt, (uint16_t)(16 * k + bpos)); }
                        cnt++;
                    }
                }
            } else {
                // more hits than the queue holds: walk again, dedupe against the candidates written so far (L slots were
                // reserved). Without room the attempt is discarded by the host and rerun with the exact size.
                uint32_t st2 = 0;
                for (int qq = 0; qq < L; qq++) {
                    st2 = big_ld<HINTS>(tr + ((size_t)st2 * 4u + pk_code(row, qq)), pol_keep);
                    if (st2 >= H0) {
                        const uint32_t info = big_ld<HINTS>(p.hit_info + (st2 - H0), pol_keep);
                        const uint32_t r = info & 0x00ffffffu;
                        const int lv = (int)(info >> 24);
                        if (lv > best) { best = lv; cnt = 0; }
                        if (lv == best) {
                            bool dup = false;
                            if (room) for (int kk = 0; kk < cnt && !dup; kk++) dup = (p.cand_rank[o + kk] == r);
                            if (!dup) {
                                if (room) { p.cand_rank[o + cnt] = r; p.cand_pos[o + cnt] = (uint16_t)qq; }
                                cnt++;
                            }
                        }
                    }
                }
            }
            p.lvl[i] = (uint8_t)best;
            p.ncand[i] = (uint16_t)cnt;
            p.cand_off[i] = o;
        }
        __syncwarp();
    }
    cp_async_wait<0>();
}

}  // namespace scb
