// scan_smem2.cuh - the shared-memory scan kernel (8.07 -> 7.63 ms at 50M x 150 bp against its predecessor scan_smem_k,
// since removed; the pipeline description in scan_smem.cuh applies).
//
// Same pipeline per warp tile (A pack, B walk: identical code), but the pick and emit phases are ONE pass over the
// hits instead of two divergent loops (SASS of scan_smem_k: ~200 warp instructions per tile in the pick loop, ~380 in
// the emit loop that walks the hit masks again, of ~3700 per tile): candidate space is reserved BEFORE the pick, by
// hit count (an upper bound of the kept candidates; the candidate arrays have holes anyway), so that each candidate
// can be written the moment it is kept; a hit of a higher level restarts the list and overwrites the slots.
// The partial last word of a read is walked with the same compile-time shifts as the full words (the variable-shift
// loop cost ~100 instructions per tile for 6 bases). Results (level, count, ordered candidates and positions per
// read) are identical; only cand_off differs.
#pragma once
#include "scan_smem.cuh"

#ifndef SCB_SCAN_BULK
#define SCB_SCAN_BULK 1    // tile staging by cp.async.bulk + mbarrier (0: per-lane 16-byte cp.async)
#endif

namespace scb {

__global__ void __launch_bounds__(1024, 1) scan_smem2_k(ScanSmemParams p) {
    extern __shared__ __align__(16) uint8_t sm[];
    // layout: trans | hit_rank | rank_level | pad16 | per warp: ASCII tile (+32) | packed tile | hit queues | hit masks
    uint16_t *s_trans = (uint16_t *)sm;
    uint32_t *s_hit = (uint32_t *)(sm + (size_t)p.ns * 8);
    uint8_t *s_lvl = (uint8_t *)(s_hit + p.n_hit);
    const int L = p.L, PW = p.PW, pitch = p.pitch;
    const int w = threadIdx.x >> 5, W = blockDim.x >> 5, l = lane_id();
    uint8_t *s_tile = sm + scan_smem_table_bytes(p.ns, p.n_hit, p.nb) + (size_t)w * scan_smem_warp_bytes(L, PW);
    uint32_t *s_pk = (uint32_t *)(s_tile + (size_t)32 * L + 32);
    uint16_t *s_q = (uint16_t *)(s_pk + (size_t)32 * pitch);
    uint16_t *s_hm = s_q + (size_t)32 * kHitQ;
#if SCB_SCAN_BULK
    uint64_t *bar = (uint64_t *)(s_tile + scan_smem_warp_bytes(L, PW) - 16);
    if (l == 0) mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    uint32_t phase = 0;
    bool pending = false;
#endif

    for (int k = threadIdx.x; k < p.ns * 2; k += blockDim.x) ((uint32_t *)s_trans)[k] = ((const uint32_t *)p.trans)[k];
    for (int k = threadIdx.x; k < p.n_hit; k += blockDim.x) s_hit[k] = p.hit_rank[k];
    for (int k = threadIdx.x; k < p.nb; k += blockDim.x) s_lvl[k] = p.rank_level[k];
    __syncthreads();                                  // the only block barrier: from here on warps run alone
    const uint32_t H4 = (uint32_t)p.H0 * 4u;
    const SmemDfa d{s_trans, s_hit, s_lvl, p.H0};
    const int full = L >> 4, tail = L & 15;
    const uint32_t *row = s_pk + (size_t)l * pitch;
    uint16_t *q = s_q + (size_t)l * kHitQ;
    uint16_t *hm = s_hm + (size_t)l * PW;
    const uint32_t q0 = (uint32_t)__cvta_generic_to_shared(q);
    const uint8_t *tb = (const uint8_t *)s_trans;
    uint64_t c_cur = 0, c_end = 0;                    // this warp's slice of the candidate arrays (uniform across lanes)

    const int64_t stride = (int64_t)gridDim.x * W;
    int64_t tile = (int64_t)w * gridDim.x + blockIdx.x;   // neighbouring CTAs take neighbouring tiles
#if SCB_SCAN_BULK
    if (tile < p.n_tiles) pending = stage_warp_tile_bulk(p, tile, s_tile, bar);
#else
    if (tile < p.n_tiles) stage_warp_tile(p, tile, s_tile);
    cp_async_commit();
#endif
    for (; tile < p.n_tiles; tile += stride) {
#if SCB_SCAN_BULK
        if (pending) { mbar_wait(bar, phase); phase ^= 1u; }
#else
        cp_async_wait<0>();
#endif
        __syncwarp();
        int rows = (int)((p.n - tile * 32 < 32) ? (p.n - tile * 32) : 32);
        // ---- A: pack ----------------------------------------------------------------------------------
        {
            const uint32_t nwords = (uint32_t)rows * (uint32_t)PW;
            const uint32_t *tw = (const uint32_t *)s_tile;
            uint32_t *gp = p.packed + tile * (int64_t)32 * PW;
            for (uint32_t t = l; t < nwords; t += 32) {
                const uint32_t r = PW == 1 ? t : __umulhi(t, p.inv_pw), k = t - r * (uint32_t)PW;   // ceil(2^32 / 1) does not fit 32 bits
                const uint32_t b = r * (uint32_t)L + 16u * k;
                const uint32_t *a = tw + (b >> 2);
                const uint32_t sh = (b & 3u) * 8u;
                const uint32_t x0 = a[0], x1 = a[1], x2 = a[2], x3 = a[3], x4 = a[4];   // the tile has 32 bytes of slack
                const uint32_t y0 = __funnelshift_r(x0, x1, sh), y1 = __funnelshift_r(x1, x2, sh), y2 = __funnelshift_r(x2, x3, sh),
                               y3 = __funnelshift_r(x3, x4, sh);
                uint32_t bad = 0;
                uint32_t wv = (pack4(y0, bad) << 24) | (pack4(y1, bad) << 16) | (pack4(y2, bad) << 8) | pack4(y3, bad);
                if (bad) wv = (pack4_masked(y0) << 24) | (pack4_masked(y1) << 16) | (pack4_masked(y2) << 8) | pack4_masked(y3);
                const int nv = L - 16 * (int)k;                   // valid bases of this word (>= 1)
                if (nv < 16) wv &= ~(0xffffffffu >> (2 * nv));
                s_pk[r * (uint32_t)pitch + k] = wv;
                gp[t] = wv;
            }
        }
        __syncwarp();
        {   // the ASCII buffer is free again: the warp's next tile streams in under phases B-D
            const int64_t nxt = tile + stride;
#if SCB_SCAN_BULK
            pending = nxt < p.n_tiles && stage_warp_tile_bulk(p, nxt, s_tile, bar);
#else
            if (nxt < p.n_tiles) stage_warp_tile(p, nxt, s_tile);
            cp_async_commit();
#endif
        }
        // ---- B: walk -----------------------------------------------------------------------------------
        const int64_t i = tile * 32 + l;
        const bool live = l < rows;
        int nh = 0, best = 0, cnt = 0, first_kept = 0;
        bool slow = false;
        if (live) {
            uint32_t e4 = 0;                                  // current state * 4
            uint32_t qp = q0;                                 // 32-bit shared address of the queue's next slot
            for (int k = 0; k < full && !slow; k++) {
                if (qp - q0 > 2u * (kHitQ - kHitQGuard)) { slow = true; break; }
                const uint32_t wv = row[k];
                uint32_t m = 0;
#pragma unroll
                for (int j = 0; j < 16; j++) {
                    const uint32_t c2 = (j < 15 ? (wv >> (29 - 2 * j)) : (wv << 1)) & 6u;      // code * 2
                    e4 = *(const uint16_t *)(tb + (e4 * 2u + c2));
                    if (e4 >= H4) { m |= 1u << j; sts_u16(qp, e4); qp += 2; }
                }
                hm[k] = (uint16_t)m;
            }
            if (tail && !slow) {
                if (qp - q0 > 2u * (kHitQ - kHitQGuard)) slow = true;
                else {
                    const uint32_t wv = row[full];
                    uint32_t m = 0;
                    // same body as the full words (compile-time shifts), left through a warp-uniform exit
#pragma unroll
                    for (int j = 0; j < 15; j++) {
                        if (j >= tail) break;
                        const uint32_t c2 = (wv >> (29 - 2 * j)) & 6u;
                        e4 = *(const uint16_t *)(tb + (e4 * 2u + c2));
                        if (e4 >= H4) { m |= 1u << j; sts_u16(qp, e4); qp += 2; }
                    }
                    hm[full] = (uint16_t)m;
                }
            }
            nh = (int)((qp - q0) >> 1);
        }
        // ---- D0: candidate space BEFORE the pick, by hit count (an upper bound of what is kept; the arrays may have
        //      holes anyway), so that the pick can write each candidate the moment it is kept ------------------------
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
        // ---- C: pick + emit in ONE pass over the hits: a cursor over the words' hit masks gives the position of hit j;
        //      a hit of a higher level restarts the list (earlier slots are overwritten), within the level first
        //      occurrences are kept (aho_search, reads.cpp:413-429, minus the running populations) ----------------------
        if (live) {
            const bool room = o + (uint64_t)v <= p.cand_cap;
            if (!slow) {
                int k = 0;
                uint32_t m = nh > 0 ? (uint32_t)hm[0] : 0u;
                for (int j = 0; j < nh; j++) {
                    while (m == 0) { k++; m = hm[k]; }            // set bits over all words == nh: k stays below PW
                    const int bpos = __ffs(m) - 1;
                    m &= m - 1;
                    const uint32_t r = s_hit[((uint32_t)q[j] >> 2) - p.H0];
                    const int lv = s_lvl[r];
                    if (lv > best) { best = lv; cnt = 0; first_kept = j; }
                    bool drop = lv != best;
                    for (int kk = first_kept; kk < j && !drop; kk++) drop = (q[kk] == r);
                    q[j] = drop ? (uint16_t)0xffffu : (uint16_t)r;
                    if (!drop) {
                        if (room) { p.cand_rank[o + cnt] = r; p.cand_pos[o + cnt] = (uint16_t)(16 * k + bpos); }
                        cnt++;
                    }
                }
            } else {
                // more hits than the queue holds: full walk with inline dedupe, L slots were reserved
                uint32_t st2 = 0;
                for (int qq = 0; qq < L; qq++) {
                    st2 = dfa_step(d, st2, pk_code(row, qq));
                    if (st2 >= (uint32_t)p.H0) {
                        const uint32_t r = s_hit[st2 - p.H0];
                        const int lv = s_lvl[r];
                        if (lv > best) { best = lv; cnt = 0; }
                        if (lv == best && !seen_before_smem(row, qq, r, d)) {
                            if (room) { p.cand_rank[o + cnt] = r; p.cand_pos[o + cnt] = (uint16_t)qq; }
                            cnt++;
                        }
                    }
                }
            }
            p.lvl[i] = (uint8_t)best;
            p.ncand[i] = (uint16_t)cnt;
            p.cand_off[i] = o;
        }
        __syncwarp();   // packed tile, queues and masks are reused by the warp's next iteration
    }
    cp_async_wait<0>();
}

}  // namespace scb
