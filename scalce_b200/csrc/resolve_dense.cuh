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
// After the second round of a block most subtiles only REPLAY the short list of reads whose decision margin is
// small ("fragile reads" below); everything else provably keeps its decision.
// Dense = two u32 words per bucket per warp in shared memory (population + lane tags), so it needs
// 8*(nb+1)*W <= ~200 KB.
#pragma once
#include <cooperative_groups.h>

#include "common.cuh"
#include "prims.cuh"

namespace scb {
namespace cg = cooperative_groups;

constexpr int kRdMaxWarps = 16;
constexpr int kRdRegCands = 8;          // candidates kept in registers per read
constexpr uint32_t kNoSel = 0xffffu;
constexpr int kRdMaxRounds = 1 << 16;
constexpr int kRdBatch = 6;             // 128-bit histogram rows fetched together in the P / E phases

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
    int pitch;                    // row pitch of every [..][nb1] array here and in shared memory: nb1 rounded up to 4 (128-bit rows)
    int *status;                  // 0 ok, 1 round cap hit
    int *rounds_out;
    unsigned long long *tstamps;  // optional [rounds][8] globaltimer stamps of CTA 0 (profiling aid)
    // reads that precede blk[0] in the job (earlier flushes, and in a sharded run the lower ranks' shards):
    // only the first-round extrapolation looks at it
    int64_t g0;
    // 0: all blocks, each iterated to its fixed point (one GPU owns the whole input order)
    // 1: ONE round over block 0 from extrapolated start counts; 2: ONE round from the histograms the previous
    //    launch left in H / Cpre. Modes 1-2 are the sharded run: `base` is supplied per round by the caller
    //    (populations of everything before this shard under the current global assignment) and is not
    //    modified; tot_out[0..nb1) receives this shard's bucket histogram, tot_out[nb1] the changed count.
    int mode;
    uint32_t *tot_out;
    // ---- incremental rounds (see "fragile reads" below); incr_T = 0 switches them off ----
    uint32_t *S0, *H0;            // [max_subtiles][pitch] start counts / histogram of each subtile at its last FULL sweep
    uint32_t *fr_buf;             // fragile records of subtile [lo, hi): words fr_buf[2*lo .. 2*hi)
    uint32_t *fr_idx;             // word offset of each record of the subtile, fr_idx[lo ..]
    uint32_t *fr_used;            // [max_subtiles] records in the list; kFrNone = no valid list
    int incr_T;                   // decision-margin threshold
    unsigned long long *incr_stat;   // optional [4]: full sweeps, incremental sweeps, incremental sweeps redone in full, records visited
    // DEFER instances only (see resolve_dense_k): stale[t] = 1 while subtile t runs on replays although its margin bound failed;
    // n_stale[round] = such subtiles after the round
    uint32_t *stale, *n_stale;
};

__device__ __forceinline__ uint32_t lanemask_ge() {
    uint32_t m;
    asm("mov.u32 %0, %%lanemask_ge;" : "=r"(m));
    return m;
}

__device__ __forceinline__ unsigned long long gtimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define RD_STAMP(k) do { if (p.tstamps && blockIdx.x == 0 && threadIdx.x == 0 && round < 4096) p.tstamps[round * 8 + (k)] = gtimer(); } while (0)

struct RdLane {   // one read's state held by one lane
    int nc; uint64_t off; uint32_t so; uint32_t r[kRdRegCands];
};

// meta (count, list offset, old slot) is fetched two steps ahead, the candidates one step ahead, so
// neither dependent global load stalls the step being decided
__device__ __forceinline__ void rd_load_meta(const RdParams &p, int64_t i, int64_t hi, RdLane &x) {
    x.nc = 0; x.off = 0; x.so = kNoSel;
    if (i < hi) { x.nc = p.ncand[i]; x.off = p.cand_off[i]; x.so = p.sel[i]; }
}
__device__ __forceinline__ void rd_load_cands(const RdParams &p, RdLane &x) {
#pragma unroll
    for (int k = 0; k < kRdRegCands; k++) x.r[k] = (k < x.nc) ? p.cand_rank[x.off + k] : 0u;
}
__device__ __forceinline__ uint32_t rd_rank(const RdParams &p, const RdLane &x, int k) {
    uint32_t v = 0;
#pragma unroll
    for (int q = 0; q < kRdRegCands; q++) v = (k == q) ? x.r[q] : v;
    if (k >= kRdRegCands) v = p.cand_rank[x.off + k];
    return v;
}

// proportional extrapolation of a bucket population `off` reads into a block that starts at n0
__device__ __forceinline__ uint32_t rd_guess(uint32_t b0, int64_t off, int64_t n0) {
    return n0 > 0 ? (uint32_t)(((unsigned long long)b0 * (unsigned long long)off) / (unsigned long long)n0) : 0u;
}

// ---- fragile reads -------------------------------------------------------------------------------------
// A decision has a MARGIN: the largest change of any pairwise difference of its candidates' counts that it
// survives (first arg-max: candidates listed before the winner must stay strictly smaller, later ones not
// larger). Between two rounds the count a read sees for bucket b changes by (change of its subtile's start
// count) + (net effect of the earlier reads of the subtile that changed), so a pairwise difference moves by at
// most 2 * (Dmax + E), Dmax = largest start-count change of the subtile, E = reads of the subtile that differ
// from the last full sweep. A full sweep therefore lists, in order, the reads whose margin is <= T ("fragile")
// together with what they saw (per candidate: bucket, count of earlier reads of the subtile in it); as long as
// 2 * (Dmax + E) <= T every other read provably keeps its decision, and a round only has to replay the list:
// counts = new start count + recorded in-subtile prefix + the deviations of earlier fragile reads. When the
// bound fails the subtile is swept in full again (which rebuilds the list).
// Record: w0 = index in subtile (16) | n candidates (8) << 16 | decision of the full sweep (8) << 24;
//         w1 = current decision; then two words per candidate: bucket rank, count seen at the full sweep.
// fr_idx holds the word offset of every record so that a replay can take 32 records at a time, one per lane.
constexpr uint32_t kFrNone = 0xffffffffu;
constexpr uint32_t kRdInf = 0x7fffffffu;

struct RdFrag {   // fragile-record sink of a full sweep (warp-uniform fields)
    uint32_t *buf, *idx; uint32_t words, recs, cap; int T; int64_t lo;   // recs = kFrNone: list abandoned (does not fit / cannot be encoded)
};

// one step: decide the 32 reads held in `cur` (one per lane) exactly as if they were processed in
// lane order, while the candidate lists of the next step and the meta of the one after stream in
__device__ __forceinline__ void rd_step(const RdParams &p, int64_t g, int64_t hi, RdLane &cur, RdLane &nxt, RdLane &nn,
                                        uint32_t *cnt, uint32_t *tag, uint32_t &changed, RdFrag &fr) {
    const uint32_t l = lane_id();
    const uint32_t lt = lanemask_lt();
    const uint32_t bit = 1u << l;
    rd_load_cands(p, nxt);                       // candidates of step g+1 (its meta arrived a step ago)
    rd_load_meta(p, g + 64 + l, hi, nn);         // meta of step g+2
    const int nc = cur.nc;
    int my_k = (nc > 0 && cur.so != kNoSel) ? (int)cur.so : -1;
    uint32_t a_cur = my_k >= 0 ? rd_rank(p, cur, my_k) : 0u;
    if (my_k >= 0) atomicOr(&tag[a_cur], bit);
    __syncwarp();
    // iterate inside the step until no lane changes: lane j is final after at most j+1 passes, in
    // practice after one or two
    uint32_t margin = kRdInf;
    while (true) {
        uint32_t best_c = 0; int best_k = 0;
        uint32_t mA = kRdInf, mB = kRdInf;       // gap to the runner-up listed before / after the winner
#pragma unroll
        for (int k = 0; k < kRdRegCands; k++)
            if (k < nc) {
                uint32_t c = cnt[cur.r[k]] + __popc(tag[cur.r[k]] & lt);
                if (k == 0) { best_c = c; best_k = 0; }
                else if (c > best_c) { mA = c - best_c; mB = kRdInf; best_c = c; best_k = k; }   // first arg-max, strict > (reads.cpp:420-421)
                else mB = min(mB, best_c - c);
            }
        for (int k = kRdRegCands; k < nc; k++) {
            uint32_t rk = p.cand_rank[cur.off + k];
            uint32_t c = cnt[rk] + __popc(tag[rk] & lt);
            if (c > best_c) { mA = c - best_c; mB = kRdInf; best_c = c; best_k = k; }
            else mB = min(mB, best_c - c);
        }
        margin = min(mA - 1u, mB);               // mA >= 1 whenever it is set
        const bool chg = nc > 0 && best_k != my_k;
        __syncwarp();
        if (chg) {
            if (my_k >= 0) atomicAnd(&tag[a_cur], ~bit);
            my_k = best_k;
            a_cur = rd_rank(p, cur, best_k);
            atomicOr(&tag[a_cur], bit);
        }
        if (!__any_sync(0xffffffffu, chg)) break;
    }
    if (fr.buf && fr.recs != kFrNone) {
        // list the fragile reads of this step, in lane (= input) order, with the counts they just saw
        const bool frag = nc >= 2 && margin <= (uint32_t)fr.T;
        const uint32_t fm = __ballot_sync(0xffffffffu, frag);
        if (fm) {
            const int64_t il = g + l - fr.lo;
            const bool bad = frag && (nc > 255 || il > 0xffff);
            const uint32_t need = frag ? 2u + 2u * (uint32_t)nc : 0u;
            const uint32_t inc = warp_incl_scan(need);
            const uint32_t total = __shfl_sync(0xffffffffu, inc, 31);
            if (__any_sync(0xffffffffu, bad) || fr.words + total > fr.cap) fr.recs = kFrNone;
            else {
                if (frag) {
                    const uint32_t ow = fr.words + (inc - need);
                    uint32_t *o = fr.buf + ow;
                    fr.idx[fr.recs + __popc(fm & lt)] = ow;
                    o[0] = (uint32_t)il | ((uint32_t)nc << 16) | ((uint32_t)my_k << 24);
                    o[1] = (uint32_t)my_k;
                    for (int k = 0; k < nc; k++) {
                        const uint32_t rk = rd_rank(p, cur, k);
                        o[2 + 2 * k] = rk;
                        o[3 + 2 * k] = cnt[rk] + __popc(tag[rk] & lt);
                    }
                }
                fr.words += total;
                fr.recs += __popc(fm);
            }
        }
    }
    if (nc > 0) {
        atomicAdd(&cnt[a_cur], 1u);              // reads.cpp:246
        tag[a_cur] = 0;
        if ((uint32_t)my_k != cur.so) { p.sel[g + l] = (uint16_t)my_k; changed++; }
    }
    __syncwarp();
}

// sweep reads [lo, hi) in order with populations cnt[] and lane tags tag[] (shared memory, one pair
// of arrays per warp): cnt[b] = population of b before the current step; tag[b] = bit set of the
// lanes of the current step assigned to b, so read i sees cnt[b] + popc(tag[b] & lanes_below(i)).
__device__ __forceinline__ uint32_t rd_sweep(const RdParams &p, int64_t lo, int64_t hi, uint32_t *cnt, uint32_t *tag, RdFrag &fr) {
    const uint32_t l = lane_id();
    uint32_t changed = 0;
    RdLane x0, x1, x2;
    rd_load_meta(p, lo + l, hi, x0);
    rd_load_cands(p, x0);
    rd_load_meta(p, lo + 32 + l, hi, x1);
    // three register sets rotate roles (deciding / candidates in flight / meta in flight) so that no
    // register is copied while its load is outstanding
    for (int64_t g = lo; g < hi; g += 96) {
        rd_step(p, g, hi, x0, x1, x2, cnt, tag, changed, fr);
        if (g + 32 < hi) rd_step(p, g + 32, hi, x1, x2, x0, cnt, tag, changed, fr);
        if (g + 64 < hi) rd_step(p, g + 64, hi, x2, x0, x1, cnt, tag, changed, fr);
    }
    return changed;
}

// replay of a subtile's fragile list (see above), 32 records at a time, one per lane. row[] holds on entry
// (new start count - start count of the full sweep) per bucket and accumulates the deviations of the reads replayed so
// far, so a candidate's count now = count recorded at the full sweep + row[bucket]. Lanes evaluate in parallel; a lane
// whose decision deviates from the full sweep changes what later lanes see, so after each such lane (in order) the
// later lanes are evaluated again - one extra pass per deviating read, and deviating reads are few.
// Returns the decisions that changed since the previous round (lane 0 only); *E = reads that now deviate.
__device__ __forceinline__ uint32_t rd_replay(const RdParams &p, int64_t lo, uint32_t *frw, const uint32_t *fidx, uint32_t nrec, uint32_t *row,
                                              uint32_t *E) {
    const uint32_t l = lane_id();
    uint32_t changed = 0, e = 0;
    for (uint32_t base = 0; base < nrec; base += 32) {
        const bool valid = base + l < nrec;
        uint32_t *rec = frw + (valid ? fidx[base + l] : 0u);
        const uint32_t hdr = valid ? rec[0] : 0u;
        const uint32_t nc = (hdr >> 16) & 255u, d0 = hdr >> 24;
        const uint32_t curk = valid ? (rec[1] & 255u) : 0u;
        uint32_t newk = d0, done_upto = 0;
        while (true) {
            if (valid && l >= done_upto) {
                uint32_t bc = 0;
                for (uint32_t k = 0; k < nc; k++) {
                    const uint32_t c = rec[3 + 2 * k] + row[rec[2 + 2 * k]];
                    if (k == 0 || c > bc) { bc = c; newk = k; }     // first arg-max, strict >
                }
            }
            const uint32_t applied = done_upto >= 32u ? 0xffffffffu : ((1u << done_upto) - 1u);
            const uint32_t dm = __ballot_sync(0xffffffffu, valid && newk != d0) & ~applied;
            if (!dm) break;
            const uint32_t j = __ffs(dm) - 1;                       // its decision is final: everything before it is applied
            if (l == j) { row[rec[2 + 2 * d0]] -= 1u; row[rec[2 + 2 * newk]] += 1u; }
            done_upto = j + 1;
            __syncwarp();
        }
        if (valid) {
            e += newk != d0 ? 1u : 0u;
            if (newk != curk) {
                rec[1] = newk;
                p.sel[lo + (hdr & 0xffffu)] = (uint16_t)newk;
                changed++;
            }
        }
        __syncwarp();
    }
    *E = __reduce_add_sync(0xffffffffu, e);
    changed = __reduce_add_sync(0xffffffffu, changed);
    return l == 0 ? changed : 0u;
}

__device__ __forceinline__ uint4 add4(uint4 a, uint4 b) { return make_uint4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ uint4 sub4(uint4 a, uint4 b) { return make_uint4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
__device__ __forceinline__ uint4 guess4(uint4 b0, int64_t off, int64_t n0) {
    return make_uint4(rd_guess(b0.x, off, n0), rd_guess(b0.y, off, n0), rd_guess(b0.z, off, n0), rd_guess(b0.w, off, n0));
}

// a warp's shared-memory row (+/-)= a global row, 128 bits per lane, four rows of loads in flight
__device__ __forceinline__ void rd_row_sub(uint32_t *row, const uint32_t *g, int Q) {
    uint4 *r4 = (uint4 *)row; const uint4 *g4 = (const uint4 *)g;
    for (int q0 = lane_id(); q0 < Q; q0 += 128) {
        uint4 v[4];
#pragma unroll
        for (int u = 0; u < 4; u++) v[u] = q0 + 32 * u < Q ? g4[q0 + 32 * u] : make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int u = 0; u < 4; u++) if (q0 + 32 * u < Q) r4[q0 + 32 * u] = sub4(r4[q0 + 32 * u], v[u]);
    }
    __syncwarp();
}
__device__ __forceinline__ void rd_row_add(uint32_t *row, const uint32_t *g, int Q) {
    uint4 *r4 = (uint4 *)row; const uint4 *g4 = (const uint4 *)g;
    for (int q0 = lane_id(); q0 < Q; q0 += 128) {
        uint4 v[4];
#pragma unroll
        for (int u = 0; u < 4; u++) v[u] = q0 + 32 * u < Q ? g4[q0 + 32 * u] : make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int u = 0; u < 4; u++) if (q0 + 32 * u < Q) r4[q0 + 32 * u] = add4(r4[q0 + 32 * u], v[u]);
    }
    __syncwarp();
}

// DEFER = true (the default since it won its A/B run: 9.03 -> 8.43 ms at 50M x 150): a subtile whose margin bound fails is NOT swept
// in full in the same round. Otherwise it is, and the whole
// round - two grid syncs, ~1775 other warps that only replay - waits for that one sequential sweep; tools/sim_resolve.c counts such
// straggler sweeps in rounds 3-8 of every large block and, with a simple time model that reproduces the measured 9 ms, attributes ~15 %
// of the kernel to them. Deferred: the subtile keeps the replay's result (fragile reads exact, the others possibly outdated) and is
// marked stale; when a round changes nothing anywhere, the stale subtiles are swept in full in the next round (which rebuilds their
// lists), and the block is finished by a quiet round without stale subtiles. A subtile whose bound holds again is not stale: the bound
// only compares the current state with the last full sweep. Exactness as before: at termination every read was re-decided exactly.
template <bool DEFER>
__global__ void __launch_bounds__(kRdMaxWarps * 32, 1) resolve_dense_k(RdParams p) {
    extern __shared__ __align__(16) uint32_t sm_cnt[];   // [W][pitch] populations, then [W][pitch] per-step lane tags
    const int W = p.W, nb1 = p.nb1, P = p.pitch, Q = p.pitch >> 2;   // Q = 128-bit quads per row
    uint32_t *sm_tag = sm_cnt + (size_t)W * P;
    cg::grid_group grid = cg::this_grid();
    const int w = threadIdx.x >> 5, l = lane_id();
    const int ncta = gridDim.x, c = blockIdx.x;
    const int total_warps = ncta * W;
    uint4 *sm_cnt4 = (uint4 *)sm_cnt;
    const uint4 *base4 = (const uint4 *)p.base;
    uint4 *H4 = (uint4 *)p.H, *Csum4 = (uint4 *)p.Csum, *Cpre4 = (uint4 *)p.Cpre;
    uint4 *S04 = (uint4 *)p.S0, *H04 = (uint4 *)p.H0;
    __shared__ uint32_t s_dmax[kRdMaxWarps];     // largest start-count change of each of my subtiles since its last full sweep
    __shared__ int s_incr[kRdMaxWarps];          // 1 = the subtile's round was an incremental replay (E phase: H = H0 + deviations)
    for (int k = threadIdx.x; k < W * P; k += blockDim.x) sm_tag[k] = 0;
    __syncthreads();
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
        bool first = p.mode != 2;
        bool verify = false;                     // DEFER: this round sweeps the stale subtiles in full
        while (true) {
                        // ---- P: start counts of my subtiles, straight into the warps' shared-memory counters ----
            // first round of a block: no histogram of the block exists yet; start from the populations
            // before the block, extrapolated proportionally to the subtile's position (only a guess:
            // it shortens convergence, the fixed point does not depend on it). Four buckets per thread
            // and 128-bit loads / stores: the phase is bound by the number of memory requests.
            RD_STAMP(0);
            const int kl = (c < nact) ? (t_hi - t_lo) : 0;
            const bool incr_on = p.incr_T > 0 && !first && ts <= 0xffff;
            if (threadIdx.x < kRdMaxWarps) { s_dmax[threadIdx.x] = 0; s_incr[threadIdx.x] = 0; }
            __syncthreads();
            for (int q = threadIdx.x; q < Q && kl > 0; q += blockDim.x) {
                const uint4 b0 = base4[q];
                if (first) {
                    for (int tl = 0; tl < kl; tl++) sm_cnt4[(size_t)tl * Q + q] = add4(b0, guess4(b0, (int64_t)(t_lo + tl) * ts, p.g0 + n0));
                } else {
                    uint4 run = add4(b0, Cpre4[(size_t)c * Q + q]);
                    for (int tb = 0; tb < kl; tb += kRdBatch) {   // loads of a batch are in flight together
                        uint4 hv[kRdBatch], sv[kRdBatch];
#pragma unroll
                        for (int u = 0; u < kRdBatch; u++) {
                            hv[u] = tb + u < kl ? H4[(size_t)(t_lo + tb + u) * Q + q] : make_uint4(0, 0, 0, 0);
                            sv[u] = (incr_on && tb + u < kl) ? S04[(size_t)(t_lo + tb + u) * Q + q] : make_uint4(0, 0, 0, 0);
                        }
#pragma unroll
                        for (int u = 0; u < kRdBatch; u++)
                            if (tb + u < kl) {
                                sm_cnt4[(size_t)(tb + u) * Q + q] = run;
                                if (incr_on) {
                                    const uint4 d = sub4(run, sv[u]);
                                    const uint32_t m = max(max(abs((int)d.x), abs((int)d.y)), max(abs((int)d.z), abs((int)d.w)));
                                    if (m > s_dmax[tb + u]) atomicMax(&s_dmax[tb + u], m);
                                }
                                run = add4(run, hv[u]);
                            }
                    }
                }
            }
            __syncthreads();
            RD_STAMP(1);
            // ---- D: sweep ------------------------------------------------------------------------------
            uint32_t ch = 0;
            if (w < kl) {
                const int t = t_lo + w;
                const int64_t lo = n0 + (int64_t)t * ts, hi = min(n1, lo + ts);
                uint32_t *cnt = sm_cnt + (size_t)w * P, *tagw = sm_tag + (size_t)w * P;
                uint32_t *frw = p.incr_T > 0 ? p.fr_buf + 2 * lo : nullptr;
                uint32_t *fidx = p.incr_T > 0 ? p.fr_idx + lo : nullptr;
                uint32_t *S0row = p.S0 + (size_t)t * P;
                bool replayed = false;
                uint32_t stat_full = 0, stat_incr = 0, stat_redo = 0, visited = 0;
                if (incr_on) {
                    const uint32_t nrec = p.fr_used[t];
                    const uint32_t dmax = s_dmax[w];
                    const bool sweep_now = DEFER && verify && p.stale[t] != 0;   // a quiet round was seen: stale subtiles are swept in full now
                    if (nrec != kFrNone && !sweep_now && (DEFER || 2u * dmax <= (uint32_t)p.incr_T)) {
                        uint32_t E = 0;
                        rd_row_sub(cnt, S0row, Q);                               // row = new start - start of the full sweep
                        ch = rd_replay(p, lo, frw, fidx, nrec, cnt, &E);
                        visited = nrec;
                        const bool bound_ok = 2u * (dmax + E) <= (uint32_t)p.incr_T;
                        if (bound_ok || DEFER) {
                            rd_row_add(cnt, S0row, Q);                           // back to new start + deviations (what the E phase expects)
                            replayed = true; stat_incr = 1;
                            if (DEFER && l == 0) {
                                p.stale[t] = bound_ok ? 0u : 1u;
                                if (!bound_ok) atomicAdd(&p.n_stale[round], 1u);
                            }
                        } else {
                            // the bound broke: sweep in full. First take the deviations out again: row -> new start counts
                            for (uint32_t r0 = l; r0 < nrec; r0 += 32) {
                                const uint32_t *rec = frw + fidx[r0];
                                const uint32_t hdr = rec[0], d0 = hdr >> 24, ck = rec[1] & 255u;
                                if (ck != d0) { atomicAdd(&cnt[rec[2 + 2 * d0]], 1u); atomicSub(&cnt[rec[2 + 2 * ck]], 1u); }
                            }
                            __syncwarp();
                            rd_row_add(cnt, S0row, Q);
                            stat_redo = 1;
                        }
                    }
                }
                if (!replayed) {
                    RdFrag fr;
                    const bool list = p.incr_T > 0 && ts <= 0xffff && !first;    // the round after a guess always sweeps in full: no list yet
                    fr.buf = list ? frw : nullptr; fr.idx = fidx;
                    fr.words = 0; fr.recs = 0; fr.cap = (uint32_t)(2 * (hi - lo)); fr.T = p.incr_T; fr.lo = lo;
                    if (list) {   // the start counts of this sweep: what the recorded counts are relative to
                        for (int q = l; q < Q; q += 32) ((uint4 *)S0row)[q] = ((const uint4 *)cnt)[q];
                        __syncwarp();
                    }
                    ch += rd_sweep(p, lo, hi, cnt, tagw, fr);
                    if (p.incr_T > 0 && l == 0) p.fr_used[t] = list ? fr.recs : kFrNone;
                    if (DEFER && l == 0) p.stale[t] = 0u;
                    stat_full = 1;
                }
                if (l == 0) s_incr[w] = replayed ? 1 : 0;
                if (p.incr_stat && l == 0) {
                    if (stat_full) atomicAdd(&p.incr_stat[0], 1ull);
                    if (stat_incr) atomicAdd(&p.incr_stat[1], 1ull);
                    if (stat_redo) atomicAdd(&p.incr_stat[2], 1ull);
                    if (visited) atomicAdd(&p.incr_stat[3], (unsigned long long)visited);
                }
            }
            ch = __reduce_add_sync(0xffffffffu, ch);
            if (l == 0 && ch) atomicAdd(&p.changed[round], ch);
            __syncthreads();
            RD_STAMP(2);
            // ---- new subtile histograms = final counters - start counts; chunk total ------------------
            for (int q = threadIdx.x; q < Q && kl > 0; q += blockDim.x) {
                const uint4 b0 = base4[q];
                uint4 tot = make_uint4(0, 0, 0, 0);
                if (first) {
                    for (int tl = 0; tl < kl; tl++) {
                        const uint4 hn = sub4(sm_cnt4[(size_t)tl * Q + q], add4(b0, guess4(b0, (int64_t)(t_lo + tl) * ts, p.g0 + n0)));
                        H4[(size_t)(t_lo + tl) * Q + q] = hn;
                        if (p.incr_T > 0) H04[(size_t)(t_lo + tl) * Q + q] = hn;
                        tot = add4(tot, hn);
                    }
                } else {
                    uint4 run = add4(b0, Cpre4[(size_t)c * Q + q]);
                    for (int tb = 0; tb < kl; tb += kRdBatch) {
                        uint4 hv[kRdBatch];
#pragma unroll
                        for (int u = 0; u < kRdBatch; u++) hv[u] = tb + u < kl ? H4[(size_t)(t_lo + tb + u) * Q + q] : make_uint4(0, 0, 0, 0);
#pragma unroll
                        for (int u = 0; u < kRdBatch; u++)
                            if (tb + u < kl) {
                                uint4 hn = sub4(sm_cnt4[(size_t)(tb + u) * Q + q], run);   // full sweep: the histogram; replay: the deviations
                                if (s_incr[tb + u]) hn = add4(hn, H04[(size_t)(t_lo + tb + u) * Q + q]);
                                else if (p.incr_T > 0) H04[(size_t)(t_lo + tb + u) * Q + q] = hn;
                                run = add4(run, hv[u]);
                                H4[(size_t)(t_lo + tb + u) * Q + q] = hn;
                                tot = add4(tot, hn);
                            }
                    }
                }
                Csum4[(size_t)c * Q + q] = tot;
            }
            RD_STAMP(3);
            __threadfence();
            grid.sync();
            RD_STAMP(4);
            // ---- column scan over chunk totals; on convergence fold the block into base ---------------
            const uint32_t chg = *((volatile uint32_t *)&p.changed[round]);
            const uint32_t nst = DEFER ? *((volatile uint32_t *)&p.n_stale[round]) : 0u;
            const bool done = (chg == 0) && nst == 0;
            if (DEFER) verify = chg == 0 && nst != 0;
            {   // one warp per quad of columns; lanes stride over the chunks, shuffle scan across lanes
                const int gw = blockIdx.x * W + w, tw = gridDim.x * W;
                for (int q = gw; q < Q; q += tw) {
                    // lane l owns chunks 5l .. 5l+4 (grid <= 160 CTAs): serial prefix inside the lane, one shuffle
                    // scan of the lane totals across the warp
                    uint4 v[5];
#pragma unroll
                    for (int u = 0; u < 5; u++) { const int cc = l * 5 + u; v[u] = cc < nact ? Csum4[(size_t)cc * Q + q] : make_uint4(0, 0, 0, 0); }
                    uint4 mine = make_uint4(0, 0, 0, 0);
#pragma unroll
                    for (int u = 0; u < 5; u++) { const uint4 t = v[u]; v[u] = mine; mine = add4(mine, t); }   // v[u] = exclusive inside the lane
                    const uint4 inc = make_uint4(warp_incl_scan(mine.x), warp_incl_scan(mine.y), warp_incl_scan(mine.z), warp_incl_scan(mine.w));
                    const uint4 ex = sub4(inc, mine);
#pragma unroll
                    for (int u = 0; u < 5; u++) { const int cc = l * 5 + u; if (cc < nact) Cpre4[(size_t)cc * Q + q] = add4(ex, v[u]); }
                    uint4 carry;
                    carry.x = __shfl_sync(0xffffffffu, inc.x, 31); carry.y = __shfl_sync(0xffffffffu, inc.y, 31);
                    carry.z = __shfl_sync(0xffffffffu, inc.z, 31); carry.w = __shfl_sync(0xffffffffu, inc.w, 31);
                    if (l == 0) {
                        const uint32_t cv[4] = {carry.x, carry.y, carry.z, carry.w};
                        for (int j = 0; j < 4; j++) {
                            const int col = 4 * q + j;
                            if (col >= nb1) break;
                            if (p.mode != 0) p.tot_out[col] = cv[j];
                            else if (done) p.base[col] += cv[j];
                        }
                    }
                }
            }
            RD_STAMP(5);
            __threadfence();
            grid.sync();
            RD_STAMP(6);
            if (p.tstamps && blockIdx.x == 0 && threadIdx.x == 0 && round < 4096) p.tstamps[round * 8 + 7] = (unsigned long long)len;
                        if (p.mode != 0) {
                if (blockIdx.x == 0 && threadIdx.x == 0) { p.tot_out[nb1] = chg; *p.rounds_out = 1; }
                return;
            }
            round++;
            first = false;
            if (done) break;
            if (round >= kRdMaxRounds - 1) { if (blockIdx.x == 0 && threadIdx.x == 0) *p.status = 1; return; }
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) { *p.rounds_out = round; }
}

// asg / end from the converged slots; base -> lifetime counts
constexpr int kFinPer = 4;   // reads per thread: the dependent loads of four reads are in flight together
__global__ void __launch_bounds__(256) resolve_finalize_k(int64_t n, const uint16_t *__restrict__ ncand, const uint64_t *__restrict__ cand_off,
                                   const uint32_t *__restrict__ cand_rank, const uint16_t *__restrict__ cand_pos,
                                   const uint16_t *__restrict__ sel, int nb, uint32_t *__restrict__ asg, uint16_t *__restrict__ endv,
                                   unsigned long long *root_count) {
    const int64_t i0 = (int64_t)blockIdx.x * (256 * kFinPer) + threadIdx.x;
    int nc[kFinPer]; uint64_t o[kFinPer];
#pragma unroll
    for (int q = 0; q < kFinPer; q++) {
        const int64_t i = i0 + q * 256;
        nc[q] = -1; o[q] = 0;
        if (i < n) { nc[q] = ncand[i]; if (nc[q] > 0) o[q] = cand_off[i] + sel[i]; }
    }
    uint32_t r[kFinPer]; uint32_t ps[kFinPer];
#pragma unroll
    for (int q = 0; q < kFinPer; q++) { r[q] = 0; ps[q] = 0; if (nc[q] > 0) { r[q] = cand_rank[o[q]]; ps[q] = cand_pos[o[q]]; } }
    uint32_t roots = 0;
#pragma unroll
    for (int q = 0; q < kFinPer; q++) {
        const int64_t i = i0 + q * 256;
        if (nc[q] < 0) continue;
        if (nc[q] == 0) { asg[i] = (uint32_t)nb; endv[i] = 0; roots++; }
        else { asg[i] = r[q]; endv[i] = (uint16_t)(ps[q] + 1); }
    }
    roots = __reduce_add_sync(0xffffffffu, roots);
    if (lane_id() == 0 && roots) atomicAdd(root_count, (unsigned long long)roots);
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
