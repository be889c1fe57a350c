// scan_smem.cuh - core scan with the automaton resident in shared memory.
//
// Persistent kernel, one CTA per SM, one thread per read of a tile. The tile's ASCII rows are
// staged into shared memory with 16-byte cp.async (coalesced, double buffered: the next tile
// streams in while the current one is walked); the DFA is a u16 transition table
// trans[state][4] whose states are renumbered so that "some core ends here" is a single compare
// (state >= H0); only then are the rank / level tables consulted.
// Per read it emits, in ONE pass: the maximum core level, the ordered list of distinct
// candidates of that level (bucket rank, position) and their count - everything aho_search
// (reads.cpp:413-429) needs except the running populations. Candidate space is handed out per
// tile with one atomicAdd; cand_off[i] records where read i's list starts.
#pragma once
#include "common.cuh"
#include "pipeline.cuh"

namespace scb {

constexpr int kHitCap = 64;   // queued hits per read before the slow path (power of two)

struct ScanSmemParams {
    const uint8_t *seq; int64_t n; int L;
    const uint16_t *trans; const uint32_t *hit_rank; const uint8_t *rank_level;
    int ns, n_hit, nb, H0, R;
    uint8_t *lvl; uint16_t *ncand; uint64_t *cand_off; uint32_t *cand_rank; uint16_t *cand_pos;
    unsigned long long *cand_total; uint64_t cand_cap;
    int64_t n_tiles;
    uint32_t *packed; int PW;     // 2-bit packed copy of every read, PW = ceil(L/16) words
};

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc) {
    uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

struct SmemDfa {
    const uint16_t *trans; const uint32_t *hit_rank; const uint8_t *rank_level; int H0;
};

// one transition on a NORMALISED byte (see norm4): entries are next-state*4, columns in permuted order
__device__ __forceinline__ uint32_t dfa_step(const SmemDfa &d, uint32_t st, uint8_t x) {
    return (uint32_t)d.trans[(st << 2) | (((uint32_t)x & 6u) >> 1)] >> 2;
}
__device__ __forceinline__ bool seen_before_smem(const uint8_t *s, int p, uint32_t r, const SmemDfa &d) {
    uint32_t st = 0;
    for (int q = 0; q < p; q++) {
        st = dfa_step(d, st, s[q]);
        if (st >= (uint32_t)d.H0 && d.hit_rank[st - d.H0] == r) return true;
    }
    return false;
}

__device__ __forceinline__ void stage_tile(const ScanSmemParams &p, int64_t tile, uint8_t *buf) {
    const int64_t row0 = tile * p.R;
    int64_t rows = p.n - row0;
    if (rows > p.R) rows = p.R;
    if (rows <= 0) return;
    const int64_t bytes = rows * p.L;
    const uint8_t *src = p.seq + row0 * p.L;
    const int64_t n16 = bytes >> 4;
    for (int64_t k = threadIdx.x; k < n16; k += blockDim.x) cp_async16(buf + (k << 4), src + (k << 4));
    for (int64_t k = (n16 << 4) + threadIdx.x; k < bytes; k += blockDim.x) buf[k] = src[k];
}

// In-place normalisation of a staged tile: every byte that is not one of ACGTacgt becomes 'A'
// (getval maps all of those to 0, const.cpp:47-49). After it, bits 1-2 of a byte are a permuted
// 2-bit base code (A 0, C 1, T 2, G 3) that indexes the permuted transition table directly.
__device__ __forceinline__ uint32_t norm4(uint32_t w) {
    const uint32_t x = w | 0x20202020u;
    // per byte: 0x80 where the byte equals the pattern (exact zero-byte test on x ^ pattern)
    auto eq = [](uint32_t x, uint32_t pat) {
        const uint32_t t = x ^ pat;
        return ~(((t & 0x7f7f7f7fu) + 0x7f7f7f7fu) | t | 0x7f7f7f7fu);
    };
    const uint32_t ok = eq(x, 0x61616161u) | eq(x, 0x63636363u) | eq(x, 0x67676767u) | eq(x, 0x74747474u);
    const uint32_t keep = (ok >> 7) * 0xffu;                 // 0xff in valid bytes
    return (w & keep) | (0x41414141u & ~keep);
}

__global__ void __launch_bounds__(1024, 1) scan_smem_k(ScanSmemParams p) {
    extern __shared__ __align__(16) uint8_t sm[];
    // layout: trans | hit_rank | rank_level | pad16 | tile0 | tile1 | scan scratch
    // trans holds next-state ids pre-multiplied by 4 and indexed by the PERMUTED base code
    uint16_t *s_trans = (uint16_t *)sm;
    uint32_t *s_hit = (uint32_t *)(sm + (size_t)p.ns * 8);
    uint8_t *s_lvl = (uint8_t *)(s_hit + p.n_hit);
    const size_t off = ((size_t)p.ns * 8 + (size_t)p.n_hit * 4 + (size_t)p.nb + 15) & ~(size_t)15;
    const size_t tile_bytes = (((size_t)p.R * p.L) + 15) & ~(size_t)15;
    uint32_t *s_scan = (uint32_t *)(sm + off + 2 * tile_bytes);   // [33]
    __shared__ unsigned long long s_base;

    for (int k = threadIdx.x; k < p.ns * 2; k += blockDim.x) ((uint32_t *)s_trans)[k] = ((const uint32_t *)p.trans)[k];
    for (int k = threadIdx.x; k < p.n_hit; k += blockDim.x) s_hit[k] = p.hit_rank[k];
    for (int k = threadIdx.x; k < p.nb; k += blockDim.x) s_lvl[k] = p.rank_level[k];
    const uint32_t H4 = (uint32_t)p.H0 * 4u;
    SmemDfa d{s_trans, s_hit, s_lvl, p.H0};

    int cur = 0;
    int64_t tile = blockIdx.x;
    if (tile < p.n_tiles) stage_tile(p, tile, sm + off);
    cp_async_commit();
    for (; tile < p.n_tiles; tile += gridDim.x, cur ^= 1) {
        uint8_t *tb = sm + off + (size_t)cur * tile_bytes;
        const int64_t nxt = tile + gridDim.x;
        if (nxt < p.n_tiles) stage_tile(p, nxt, sm + off + (size_t)(cur ^ 1) * tile_bytes);
        cp_async_commit();
        cp_async_wait<1>();
        __syncthreads();
        {   // normalise the tile (coalesced, branch free)
            int64_t rows = p.n - tile * p.R; if (rows > p.R) rows = p.R;
            const int nw4 = (int)((rows * p.L + 15) >> 4);
            uint4 *t4 = (uint4 *)tb;
            for (int k = threadIdx.x; k < nw4; k += blockDim.x) {
                uint4 v = t4[k];
                v.x = norm4(v.x); v.y = norm4(v.y); v.z = norm4(v.z); v.w = norm4(v.w);
                t4[k] = v;
            }
        }
        __syncthreads();

        const int64_t i = tile * p.R + threadIdx.x;
        const bool live = i < p.n;
        // hot loop: one table lookup per base; a hit (some core ends here) only queues (state, pos)
        uint32_t hits[kHitCap];
        int nh = 0, best = 0, cnt = 0;
        const uint8_t *s = tb + (size_t)threadIdx.x * p.L;
        if (live) {
            uint32_t e4 = 0;                                  // current state * 4
            uint32_t *prow = p.packed + i * (int64_t)p.PW;
            const int full = p.L >> 4;
            for (int k = 0; k < full; k++) {                  // 16 bases per packed word
                uint32_t acc = 0;
#pragma unroll
                for (int j = 0; j < 16; j++) {
                    const uint32_t x = s[16 * k + j];
                    const uint32_t c2 = x & 6u;                                  // permuted code * 2
                    acc = acc * 4u + (c2 >> 1);
                    e4 = *(const uint16_t *)((const uint8_t *)s_trans + (e4 * 2u + c2));
                    if (e4 >= H4) { hits[nh & (kHitCap - 1)] = (e4 << 14) | (uint32_t)(16 * k + j); nh++; }
                }
                prow[k] = acc ^ ((acc >> 1) & 0x55555555u);   // permuted codes (A0 C1 T2 G3) -> A0 C1 G2 T3
            }
            if (p.L & 15) {
                uint32_t acc = 0;
                for (int q = full << 4; q < p.L; q++) {
                    const uint32_t c2 = (uint32_t)s[q] & 6u;
                    acc = acc * 4u + (c2 >> 1);
                    e4 = *(const uint16_t *)((const uint8_t *)s_trans + (e4 * 2u + c2));
                    if (e4 >= H4) { hits[nh & (kHitCap - 1)] = (e4 << 14) | (uint32_t)q; nh++; }
                }
                acc = acc ^ ((acc >> 1) & 0x55555555u);
                prow[full] = acc << (2 * (16 - (p.L & 15)));
            }
            if (nh <= kHitCap) {
                // max level, then keep the first occurrence of each bucket of that level
                for (int j = 0; j < nh; j++) {
                    uint32_t r = s_hit[(hits[j] >> 16) - p.H0];
                    int lv = s_lvl[r];
                    best = lv > best ? lv : best;
                    hits[j] = (r << 16) | (hits[j] & 0xffffu);
                }
                for (int j = 0; j < nh; j++) {
                    uint32_t r = hits[j] >> 16;
                    bool drop = (int)s_lvl[r] != best;
                    for (int k = 0; k < j && !drop; k++) drop = (hits[k] != 0xffffffffu) && ((hits[k] >> 16) == r);
                    if (drop) hits[j] = 0xffffffffu; else cnt++;
                }
            } else {
                // more hits than the queue holds (very dense core sets): full walk with inline dedupe
                uint32_t st2 = 0;
                for (int q = 0; q < p.L; q++) {
                    st2 = dfa_step(d, st2, s[q]);
                    if (st2 >= (uint32_t)p.H0) {
                        uint32_t r = s_hit[st2 - p.H0];
                        int lv = s_lvl[r];
                        if (lv > best) { best = lv; cnt = 0; }
                        if (lv == best && !seen_before_smem(s, q, r, d)) cnt++;
                    }
                }
            }
        }
        // candidate space for the tile: block scan of counts + one atomicAdd
        (void)0;
        uint32_t v = live ? (uint32_t)cnt : 0u;
        uint32_t inc = warp_incl_scan(v);
        const int w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
        if (lane_id() == 31) s_scan[w] = inc;
        __syncthreads();
        if (w == 0) {
            uint32_t x = (int)lane_id() < nw ? s_scan[lane_id()] : 0u;
            uint32_t xi = warp_incl_scan(x);
            s_scan[lane_id()] = xi - x;
            if (lane_id() == 31) {
                s_scan[32] = xi;
                s_base = xi ? atomicAdd(p.cand_total, (unsigned long long)xi) : 0ull;
            }
        }
        __syncthreads();
        if (live) {
            const uint64_t o = s_base + s_scan[w] + (inc - v);
            p.lvl[i] = (uint8_t)best;
            p.ncand[i] = (uint16_t)cnt;
            p.cand_off[i] = o;
            if (o + (uint64_t)cnt <= p.cand_cap) {
                if (nh <= kHitCap) {
                    int c2 = 0;
                    for (int j = 0; j < nh; j++)
                        if (hits[j] != 0xffffffffu) { p.cand_rank[o + c2] = hits[j] >> 16; p.cand_pos[o + c2] = (uint16_t)(hits[j] & 0xffffu); c2++; }
                } else {
                    uint32_t st2 = 0; int c2 = 0;
                    for (int q = 0; q < p.L; q++) {
                        st2 = dfa_step(d, st2, s[q]);
                        if (st2 >= (uint32_t)p.H0) {
                            uint32_t r = s_hit[st2 - p.H0];
                            if ((int)s_lvl[r] == best && !seen_before_smem(s, q, r, d)) { p.cand_rank[o + c2] = r; p.cand_pos[o + c2] = (uint16_t)q; c2++; }
                        }
                    }
                }
            }
        }
        __syncthreads();   // the tile buffer and s_scan are reused next iteration
    }
    cp_async_wait<0>();
}

}  // namespace scb
