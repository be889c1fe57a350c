// scan_smem.cuh - core scan with the automaton resident in shared memory.
//
// Persistent kernel, one CTA per SM; every WARP runs its own pipeline over tiles of 32 reads and never
// waits for another warp (no block barriers after the table is loaded), so the stalls of one warp's phase
// are covered by the other warps' phases. Per warp tile:
//   A  pack   the tile's 32 ASCII rows arrive by 16-byte cp.async (the warp's NEXT tile streams in while this
//             one is walked); the lanes turn them into 2-bit packed words, 16 bases per word, SWAR on four
//             bytes at a time (bytes that are not ACGT/acgt become A, const.cpp:47-49). The words go to
//             global memory (coalesced: a tile's packed rows are one contiguous run) and to a shared-memory
//             copy with an odd row pitch (conflict-free for phase B).
//   B  walk   one lane per read: one 32-bit shared load per 16 bases, then per base
//             shift+mask -> address -> LDS.U16 of the u16 transition table (entries = next state * 4) ->
//             compare. States are renumbered so that "some core ends here" is state >= H0; a hit sets a
//             bit in the word's position mask and appends the state to a per-read u16 queue - three
//             predicated instructions, no divergence.
//   C  pick   per read, one pass over its hits: bucket rank, maximum core level, ordered distinct
//             candidates of that level (everything aho_search, reads.cpp:413-429, needs except the
//             running populations).
//   D  emit   candidate space comes from a per-warp bump allocator refilled from one global counter in
//             chunks (no block-wide scan); lists, counts, offsets, level out.
// Reads with more hits than the queue holds (very dense core sets) take a slower exact path.
#pragma once
#include "common.cuh"
#include "pipeline.cuh"

namespace scb {

constexpr int kHitQ = 32;      // queued hit states per read (u16 each)
constexpr int kHitQGuard = 16; // a 16-base word is only walked on the fast path if it cannot overflow the queue

struct ScanSmemParams {
    const uint8_t *seq; int64_t n; int L;
    const uint16_t *trans; const uint32_t *hit_rank; const uint8_t *rank_level;
    int ns, n_hit, nb, H0, R;
    uint8_t *lvl; uint16_t *ncand; uint64_t *cand_off; uint32_t *cand_rank; uint16_t *cand_pos;
    unsigned long long *cand_total; uint64_t cand_cap;   // bump counter over the candidate arrays (holes allowed)
    int64_t n_tiles;              // tiles of 32 reads
    uint32_t *packed; int PW;     // 2-bit packed copy of every read, PW = ceil(L/16) words
    uint32_t inv_pw;              // ceil(2^32 / PW)
    int pitch;                    // shared-memory row pitch of the packed tile in words (odd)
};

// shared-memory footprint (host + device agree through these); W = warps per CTA
constexpr int kCandChunk = 4096;   // candidate slots a warp takes from the global counter at a time
__host__ __device__ inline size_t scan_smem_table_bytes(int ns, int n_hit, int nb) {
    return (((size_t)ns * 8 + (size_t)n_hit * 4 + (size_t)nb) + 15) & ~(size_t)15;
}
__host__ __device__ inline int scan_smem_pitch(int PW) { return PW | 1; }
__host__ __device__ inline size_t scan_smem_warp_bytes(int L, int PW) {
    const size_t tile = (size_t)32 * L + 32;                        // 32*L is a multiple of 16
    const size_t pk = (size_t)32 * scan_smem_pitch(PW) * 4, q = (size_t)32 * kHitQ * 2, hm = (((size_t)32 * PW * 2) + 15) & ~(size_t)15;
    return ((tile + pk + q + hm) + 15) & ~(size_t)15;
}
__host__ __device__ inline size_t scan_smem_total(int ns, int n_hit, int nb, int W, int L, int PW) {
    return scan_smem_table_bytes(ns, n_hit, nb) + (size_t)W * scan_smem_warp_bytes(L, PW);
}

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc) {
    uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// one warp stages its tile (32 rows, or fewer at the end of the input) into its own buffer
__device__ __forceinline__ void stage_warp_tile(const ScanSmemParams &p, int64_t tile, uint8_t *buf) {
    const int64_t row0 = tile * 32;
    int64_t rows = p.n - row0;
    if (rows > 32) rows = 32;
    if (rows <= 0) return;
    const int bytes = (int)rows * p.L;
    const uint8_t *src = p.seq + row0 * p.L;
    const int n16 = bytes >> 4;
    for (int k = lane_id(); k < n16; k += 32) cp_async16(buf + (k << 4), src + ((int64_t)k << 4));
    for (int k = (n16 << 4) + lane_id(); k < bytes; k += 32) buf[k] = src[k];
}

__device__ __forceinline__ void sts_u16(uint32_t saddr, uint32_t v) {
    asm volatile("st.shared.u16 [%0], %1;\n" ::"r"(saddr), "h"((uint16_t)v) : "memory");
}

// four ASCII bases (one per byte, first base in the low byte) -> 8 bits, first base in the top two bits.
// *bad gets a non-zero bit for any byte that is not one of ACGTacgt.
__device__ __forceinline__ uint32_t pack4(uint32_t w, uint32_t &bad) {
    const uint32_t x = w | 0x20202020u;
    const uint32_t cw = (x >> 1) & 0x03030303u;                    // a 0, c 1, t 2, g 3
    const uint32_t t2 = (cw >> 1) & ~cw & 0x01010101u;             // 1 in the bytes that hold t
    const uint32_t r = 0x61616161u + (cw << 1) + ((t2 << 4) - t2); // the letter each code stands for
    bad |= r ^ x;
    const uint32_t code = cw ^ ((cw >> 1) & 0x01010101u);          // a 0, c 1, g 2, t 3
    return (code * 0x40100401u) >> 24;
}

// same, exact for any byte: codes of bytes that are not ACGTacgt are forced to 0 (A)
__device__ __forceinline__ uint32_t pack4_masked(uint32_t w) {
    const uint32_t x = w | 0x20202020u;
    const uint32_t cw = (x >> 1) & 0x03030303u;
    const uint32_t t2 = (cw >> 1) & ~cw & 0x01010101u;
    const uint32_t r = 0x61616161u + (cw << 1) + ((t2 << 4) - t2);
    const uint32_t df = r ^ x;
    const uint32_t nz = (df | ((df & 0x7f7f7f7fu) + 0x7f7f7f7fu)) & 0x80808080u;   // 0x80 in the bytes where df != 0
    uint32_t code = cw ^ ((cw >> 1) & 0x01010101u);
    code &= ~((nz >> 7) | (nz >> 6));
    return (code * 0x40100401u) >> 24;
}

// the DFA over 2-bit codes, for the rare paths (queue overflow)
struct SmemDfa {
    const uint16_t *trans; const uint32_t *hit_rank; const uint8_t *rank_level; int H0;
};
__device__ __forceinline__ uint32_t pk_code(const uint32_t *row, int q) { return (row[q >> 4] >> (30 - 2 * (q & 15))) & 3u; }
__device__ __forceinline__ uint32_t dfa_step(const SmemDfa &d, uint32_t st, uint32_t c) { return (uint32_t)d.trans[(st << 2) | c] >> 2; }
__device__ __forceinline__ bool seen_before_smem(const uint32_t *row, int p, uint32_t r, const SmemDfa &d) {
    uint32_t st = 0;
    for (int q = 0; q < p; q++) {
        st = dfa_step(d, st, pk_code(row, q));
        if (st >= (uint32_t)d.H0 && d.hit_rank[st - d.H0] == r) return true;
    }
    return false;
}

__global__ void __launch_bounds__(1024, 1) scan_smem_k(ScanSmemParams p) {
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
    if (tile < p.n_tiles) stage_warp_tile(p, tile, s_tile);
    cp_async_commit();
    for (; tile < p.n_tiles; tile += stride) {
        cp_async_wait<0>();
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
            if (nxt < p.n_tiles) stage_warp_tile(p, nxt, s_tile);
            cp_async_commit();
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
                    for (int j = 0; j < tail; j++) {
                        const uint32_t c2 = ((wv >> (30 - 2 * j)) & 3u) * 2u;
                        e4 = *(const uint16_t *)(tb + (e4 * 2u + c2));
                        if (e4 >= H4) { m |= 1u << j; sts_u16(qp, e4); qp += 2; }
                    }
                    hm[full] = (uint16_t)m;
                }
            }
            nh = (int)((qp - q0) >> 1);
            // ---- C: pick ------------------------------------------------------------------------------
            if (!slow) {
                // one pass: a hit of a higher level restarts the list; within the level keep first occurrences.
                // q[j] becomes the bucket rank (n_buckets < 2^14 here) or 0xffff for a dropped hit.
                for (int j = 0; j < nh; j++) {
                    const uint32_t r = s_hit[((uint32_t)q[j] >> 2) - p.H0];
                    const int lv = s_lvl[r];
                    if (lv > best) { best = lv; cnt = 0; first_kept = j; }
                    bool drop = lv != best;
                    for (int k = first_kept; k < j && !drop; k++) drop = (q[k] == r);
                    q[j] = drop ? (uint16_t)0xffffu : (uint16_t)r;
                    cnt += drop ? 0 : 1;
                }
            } else {
                // more hits than the queue holds: full walk with inline dedupe
                uint32_t st2 = 0;
                for (int qq = 0; qq < L; qq++) {
                    st2 = dfa_step(d, st2, pk_code(row, qq));
                    if (st2 >= (uint32_t)p.H0) {
                        const uint32_t r = s_hit[st2 - p.H0];
                        const int lv = s_lvl[r];
                        if (lv > best) { best = lv; cnt = 0; }
                        if (lv == best && !seen_before_smem(row, qq, r, d)) cnt++;
                    }
                }
            }
        }
        // ---- D: candidate space from the warp's slice (refilled in chunks from the global counter) -------
        const uint32_t v = live ? (uint32_t)cnt : 0u;
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
        if (live) {
            p.lvl[i] = (uint8_t)best;
            p.ncand[i] = (uint16_t)cnt;
            p.cand_off[i] = o;
            if (o + (uint64_t)cnt <= p.cand_cap) {
                if (!slow) {
                    int c2 = 0, j = 0;
                    for (int k = 0; k < PW && c2 < cnt; k++) {
                        uint32_t m = hm[k];
                        while (m) {
                            const int bpos = __ffs(m) - 1;
                            m &= m - 1;
                            const uint32_t r = (j >= first_kept) ? (uint32_t)q[j] : 0xffffu;
                            j++;
                            if (r != 0xffffu) { p.cand_rank[o + c2] = r; p.cand_pos[o + c2] = (uint16_t)(16 * k + bpos); c2++; }
                        }
                    }
                } else {
                    uint32_t st2 = 0; int c2 = 0;
                    for (int qq = 0; qq < L; qq++) {
                        st2 = dfa_step(d, st2, pk_code(row, qq));
                        if (st2 >= (uint32_t)p.H0) {
                            const uint32_t r = s_hit[st2 - p.H0];
                            if ((int)s_lvl[r] == best && !seen_before_smem(row, qq, r, d)) { p.cand_rank[o + c2] = r; p.cand_pos[o + c2] = (uint16_t)qq; c2++; }
                        }
                    }
                }
            }
        }
        __syncwarp();   // packed tile, queues and masks are reused by the warp's next iteration
    }
    cp_async_wait<0>();
}

}  // namespace scb
