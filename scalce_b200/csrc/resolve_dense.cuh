// resolve_dense.cuh - the stateful tie-break (reads.cpp:420-421 + 246) made parallel, exactly.
//
// Sequential rule: read i goes to the FIRST candidate with the largest lifetime population
// cnt_i(b) = #{j < i assigned to b}; then cnt(b)++. The sequential answer is the unique fixed point
// of "re-decide every read from the prefix counts of the current assignment". This engine iterates
// to that fixed point over geometrically growing blocks of the input (block k+1 is as long as
// everything before it): inside a block almost all decisions have margins far larger than the
// block can perturb, so a handful of rounds suffice; total work is a small multiple of one pass.
//
// One persistent cooperative kernel, one CTA per SM. Per round:
//   P  each CTA turns the per-subtile histograms of its chunk into per-subtile start counts
//   D  each warp sweeps one subtile in input order, 32 reads per step, with the bucket populations
//      of "everything before" in shared memory (cnt[nb+1], u32)
//   --grid sync--  column scan of the chunk totals  --grid sync--  converged?
// Dense = one u32 counter per bucket per warp in shared memory, so it needs 4*(nb+1)*W <= ~200 KB.
#pragma once
#include <cooperative_groups.h>

#include "common.cuh"

namespace scb {
namespace cg = cooperative_groups;

constexpr int kRdMaxWarps = 16;
constexpr int kRdRegCands = 8;          // candidates kept in registers per read
constexpr uint32_t kNoSel = 0xffffu;
constexpr int kRdMaxRounds = 1 << 16;

struct RdParams {
    int64_t n;
    const uint16_t *ncand; const uint64_t *cand_off; const uint32_t *cand_rank;
    uint16_t *sel;                // [n] chosen candidate slot, kNoSel = undecided / no candidate
    uint32_t *base;               // [nb1] populations before the current block (absolute)
    uint32_t *H, *S;              // [max_subtiles][nb1] subtile histograms / start counts
    uint32_t *Csum, *Cpre;        // [grid][nb1] chunk totals / exclusive prefix over chunks
    uint32_t *changed;            // [kRdMaxRounds]
    const int64_t *blk;           // [nblk+1] block boundaries
    int nblk, nb1, W;
    int *status;                  // 0 ok, 1 round cap hit
    int *rounds_out;
};

__device__ __forceinline__ uint32_t lanemask_ge() {
    uint32_t m;
    asm("mov.u32 %0, %%lanemask_ge;" : "=r"(m));
    return m;
}

struct RdLane {   // one read's state held by one lane
    int nc; uint64_t off; uint32_t so; uint32_t r[kRdRegCands];
};

__device__ __forceinline__ void rd_load(const RdParams &p, int64_t i, int64_t hi, RdLane &x) {
    x.nc = 0; x.off = 0; x.so = kNoSel;
#pragma unroll
    for (int k = 0; k < kRdRegCands; k++) x.r[k] = 0;
    if (i < hi) {
        x.nc = p.ncand[i];
        x.off = p.cand_off[i];
        x.so = p.sel[i];
#pragma unroll
        for (int k = 0; k < kRdRegCands; k++)
            if (k < x.nc) x.r[k] = p.cand_rank[x.off + k];
    }
}
__device__ __forceinline__ uint32_t rd_rank(const RdParams &p, const RdLane &x, int k) {
    uint32_t v = 0;
#pragma unroll
    for (int q = 0; q < kRdRegCands; q++) v = (k == q) ? x.r[q] : v;
    if (k >= kRdRegCands) v = p.cand_rank[x.off + k];
    return v;
}

// sweep reads [lo, hi) in order with populations cnt[] (shared memory, one array per warp)
__device__ __forceinline__ uint32_t rd_sweep(const RdParams &p, int64_t lo, int64_t hi, uint32_t *cnt) {
    const uint32_t l = lane_id();
    uint32_t changed = 0;
    RdLane cur, nxt;
    rd_load(p, lo + l, hi, cur);
    for (int64_t g = lo; g < hi; g += 32) {
        rd_load(p, g + 32 + l, hi, nxt);   // prefetch the next step while this one is decided
        const int64_t i = g + l;
        const int nc = cur.nc;
        const bool has_old = nc > 0 && cur.so != kNoSel;
        const uint32_t a_old = has_old ? rd_rank(p, cur, (int)cur.so) : 0xffffffffu;
        // pass 1: first arg-max on the populations before this step
        uint32_t best_c = 0, sum1 = 0; int best_k = 0;
#pragma unroll
        for (int k = 0; k < kRdRegCands; k++)
            if (k < nc) { uint32_t c = cnt[cur.r[k]]; sum1 += c; if (k == 0 || c > best_c) { best_c = c; best_k = k; } }
        for (int k = kRdRegCands; k < nc; k++) { uint32_t c = cnt[p.cand_rank[cur.off + k]]; sum1 += c; if (c > best_c) { best_c = c; best_k = k; } }
        __syncwarp();
        if (has_old) atomicAdd(&cnt[a_old], 1u);      // old assignments of this step become visible
        __syncwarp();
        uint32_t sum2 = 0;
#pragma unroll
        for (int k = 0; k < kRdRegCands; k++)
            if (k < nc) sum2 += cnt[cur.r[k]];
        for (int k = kRdRegCands; k < nc; k++) sum2 += cnt[p.cand_rank[cur.off + k]];
        // somebody else in this step sits in one of my candidate buckets -> count only the earlier lanes
        bool flagged = nc > 0 && (sum2 - sum1) != (has_old ? 1u : 0u);
        uint32_t m = __ballot_sync(0xffffffffu, flagged);
        while (m) {
            const int src = __ffs(m) - 1;
            m &= m - 1;
            const int nk = __shfl_sync(0xffffffffu, nc, src);
            uint32_t bc = 0; int bk = 0;
            for (int k = 0; k < nk; k++) {
                uint32_t mine = (l == (uint32_t)src) ? rd_rank(p, cur, k) : 0u;
                uint32_t rk = __shfl_sync(0xffffffffu, mine, src);
                uint32_t bal = __ballot_sync(0xffffffffu, has_old && a_old == rk);
                if (l == (uint32_t)src) {
                    uint32_t c = cnt[rk] - __popc(bal & lanemask_ge());
                    if (k == 0 || c > bc) { bc = c; bk = k; }
                }
            }
            if (l == (uint32_t)src) best_k = bk;
        }
        __syncwarp();
        if (nc > 0) {
            uint32_t a_new = rd_rank(p, cur, best_k);
            if (!has_old) { atomicAdd(&cnt[a_new], 1u); p.sel[i] = (uint16_t)best_k; changed++; }
            else if ((uint32_t)best_k != cur.so) { atomicSub(&cnt[a_old], 1u); atomicAdd(&cnt[a_new], 1u); p.sel[i] = (uint16_t)best_k; changed++; }
        }
        __syncwarp();
        cur = nxt;
    }
    return changed;
}

__global__ void __launch_bounds__(kRdMaxWarps * 32, 1) resolve_dense_k(RdParams p) {
    extern __shared__ uint32_t sm_cnt[];   // [W][nb1]
    cg::grid_group grid = cg::this_grid();
    const int W = p.W, nb1 = p.nb1;
    const int w = threadIdx.x >> 5, l = lane_id();
    const int ncta = gridDim.x, c = blockIdx.x;
    const int total_warps = ncta * W;
    int round = 0;
    for (int b = 0; b < p.nblk; b++) {
        const int64_t n0 = p.blk[b], n1 = p.blk[b + 1];
        const int64_t len = n1 - n0;
        int64_t ts = (len + total_warps - 1) / total_warps;
        ts = ((ts + 31) / 32) * 32;
        if (ts < 32) ts = 32;
        const int ns = (int)((len + ts - 1) / ts);            // subtiles in this block (<= total_warps)
        const int k = (ns + ncta - 1) / ncta;                 // subtiles per CTA (<= W)
        const int nact = (ns + k - 1) / k;                    // CTAs with work
        const int t_lo = c * k, t_hi = min(ns, t_lo + k);
        bool first = true;
        while (true) {
            // ---- P: start counts of my subtiles -------------------------------------------------
            if (c < nact) {
                for (int col = threadIdx.x; col < nb1; col += blockDim.x) {
                    uint32_t run = p.base[col] + (first ? 0u : p.Cpre[(size_t)c * nb1 + col]);
                    for (int t = t_lo; t < t_hi; t++) {
                        p.S[(size_t)t * nb1 + col] = run;
                        if (!first) run += p.H[(size_t)t * nb1 + col];
                    }
                }
            }
            __syncthreads();
            // ---- D: sweep ------------------------------------------------------------------------------
            uint32_t ch = 0;
            const int t = t_lo + w;
            if (c < nact && w < k && t < t_hi) {
                uint32_t *cnt = sm_cnt + (size_t)w * nb1;
                for (int col = l; col < nb1; col += 32) cnt[col] = p.S[(size_t)t * nb1 + col];
                __syncwarp();
                const int64_t lo = n0 + (int64_t)t * ts, hi = min(n1, lo + ts);
                ch = rd_sweep(p, lo, hi, cnt);
                __syncwarp();
                for (int col = l; col < nb1; col += 32) p.H[(size_t)t * nb1 + col] = cnt[col] - p.S[(size_t)t * nb1 + col];
            }
            ch = __reduce_add_sync(0xffffffffu, ch);
            if (l == 0 && ch) atomicAdd(&p.changed[round], ch);
            __syncthreads();
            if (c < nact) {
                for (int col = threadIdx.x; col < nb1; col += blockDim.x) {
                    uint32_t s = 0;
                    for (int tt = t_lo; tt < t_hi; tt++) s += p.H[(size_t)tt * nb1 + col];
                    p.Csum[(size_t)c * nb1 + col] = s;
                }
            }
            __threadfence();
            grid.sync();
            // ---- column scan over chunk totals; on convergence fold the block into base ---------------
            const uint32_t chg = *((volatile uint32_t *)&p.changed[round]);
            const bool done = (chg == 0);
            for (int64_t col = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; col < nb1; col += (int64_t)gridDim.x * blockDim.x) {
                uint32_t run = 0;
                for (int cc = 0; cc < nact; cc++) {
                    uint32_t v = p.Csum[(size_t)cc * nb1 + col];
                    p.Cpre[(size_t)cc * nb1 + col] = run;
                    run += v;
                }
                if (done) p.base[col] += run;
            }
            __threadfence();
            grid.sync();
            round++;
            first = false;
            if (done) break;
            if (round >= kRdMaxRounds - 1) { if (blockIdx.x == 0 && threadIdx.x == 0) *p.status = 1; return; }
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) { *p.rounds_out = round; }
}

// asg / end from the converged slots; base -> lifetime counts
__global__ void resolve_finalize_k(int64_t n, const uint16_t *__restrict__ ncand, const uint64_t *__restrict__ cand_off,
                                   const uint32_t *__restrict__ cand_rank, const uint16_t *__restrict__ cand_pos,
                                   const uint16_t *__restrict__ sel, int nb, uint32_t *__restrict__ asg, uint16_t *__restrict__ endv,
                                   unsigned long long *root_count) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool root = false;
    if (i < n) {
        int nc = ncand[i];
        if (nc == 0) { asg[i] = (uint32_t)nb; endv[i] = 0; root = true; }
        else {
            uint64_t o = cand_off[i] + sel[i];
            asg[i] = cand_rank[o];
            endv[i] = (uint16_t)(cand_pos[o] + 1);
        }
    }
    uint32_t m = __ballot_sync(0xffffffffu, root);
    if (lane_id() == 0 && m) atomicAdd(root_count, (unsigned long long)__popc(m));
}
__global__ void life_to_base_k(const unsigned long long *life, uint32_t *base, int nb1) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nb1) base[i] = (uint32_t)life[i];
}
__global__ void base_to_life_k(const uint32_t *base, unsigned long long *life, int nb) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nb) life[i] = base[i];   // root (index nb) is accumulated by resolve_finalize_k
}

}  // namespace scb
