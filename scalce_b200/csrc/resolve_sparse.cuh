// resolve_sparse.cuh - the stateful tie-break (reads.cpp:420-421 + 246) for core sets of production size
// (the reference sizes patterns[] for 5-10 M cores, reads.cpp:336, 385), where the dense engine's per-warp
// population rows (resolve_dense.cuh: 8 bytes of shared memory per bucket per warp) no longer fit.
//
// Same fixed point, different data structure. The sequential answer is the unique fixed point of "re-decide every
// read from the prefix counts of the current assignment" (DESIGN.md section 5). Here the prefix counts come from a
// bucket-major copy of the candidate lists:
//   pair        = (read i, its k-th candidate bucket b); pairs are numbered read-major (p = doff[i] + k)
//   sorted view = the pairs stably sorted by bucket (one LSD radix sort per flush), so a bucket's pairs are
//                 contiguous and in input order: the number of earlier reads currently assigned to b that
//                 pair (b, i) sees is a SEGMENTED exclusive prefix sum of the flag "read selects this pair"
// One round = flags + per-tile tails (sp_flags_k) -> segmented scan of the tile tails (sp_tilescan_k) ->
// per-pair counts scattered back to read-major order (sp_counts_k) -> every read re-decides (sp_decide_k).
// A round that changes nothing proves the assignment is the sequential one. All kernels stream; the two random
// accesses per pair (flag gather, count scatter) are what a round costs, so they are only made for DIRTY buckets: a
// bucket whose flags changed in the previous round, or whose population before the local reads changed (sharded
// rounds). In a clean bucket neither the flags nor the counts its pairs see can have changed, and the set of dirty
// buckets roughly halves per round (measured on the CPU model: 28 %, 14 %, 7 %, ... of the pairs).
// One GPU owning the input order walks it in BLOCKS of reads (Gauss-Seidel across blocks, Jacobi inside): the sort key is
// (block, bucket), a block's pairs are one contiguous range of the sorted view, and a block is iterated to its fixed point from
// the populations every earlier block left behind. A decision depends on earlier reads only, so chains of dependent decisions
// are cut at the block boundaries: a block (128 K reads, doubling up to 4 M) needs 6-11 rounds where the whole flush needed 30,
// and every round touches only that block's pairs (50 M reads x 1 M cores: 122 -> 23.5 ms).
// The same round serves the sharded run (one global round per call, populations of the lower ranks supplied by
// the caller) - include/scalce_b200.h "Sharded run".
#pragma once
#include "common.cuh"
#include "prims.cuh"

namespace scb {

constexpr int kSpThreads = 256;
constexpr int kSpItems = 8;                       // consecutive sorted pairs per thread
constexpr int kSpTile = kSpThreads * kSpItems;    // 2048 pairs per CTA

// ---- set-up: read-major pair arrays ----------------------------------------------------------------------------
// doff[i] = exclusive prefix of ncand (dense numbering: the scan's candidate arrays have holes)
__global__ void __launch_bounds__(256) sp_pairs_k(int64_t n, const uint16_t *__restrict__ ncand, const uint64_t *__restrict__ cand_off,
                                                  const uint32_t *__restrict__ cand_rank, const uint64_t *__restrict__ doff,
                                                  uint32_t *__restrict__ key, uint32_t *__restrict__ val, uint32_t *__restrict__ pread,
                                                  uint16_t *__restrict__ sel, const int64_t *__restrict__ blk_first, int nblk, int rank_bits) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int lo = 0, hi = nblk - 1;                                         // last block whose first read is <= i
    while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (blk_first[mid] <= i) lo = mid; else hi = mid - 1; }
    const uint32_t blk = (uint32_t)lo << rank_bits;                    // input-order block of the read: the sort groups (block, bucket)
    const int nc = ncand[i];
    // start of the iteration: the first candidate. Any start converges to the same fixed point; a warm start by candidate-pair
    // populations ("the bucket most reads can choose") was measured and needed MORE rounds (36 vs 30 at 50M reads x 1M cores).
    sel[i] = nc > 0 ? (uint16_t)0 : (uint16_t)0xffffu;
    const uint64_t src = cand_off[i], d = doff[i];
    for (int k = 0; k < nc; k++) {
        key[d + k] = blk | cand_rank[src + k];
        val[d + k] = (uint32_t)(d + k);
        pread[d + k] = (uint32_t)i;
    }
}

// sorted view: read and candidate slot of the pair at sorted position s (the sorted keys are sb, the sorted values sval)
__global__ void __launch_bounds__(256) sp_post_k(int64_t M, const uint32_t *__restrict__ sval,
                                                 const uint32_t *__restrict__ pread, const uint64_t *__restrict__ doff,
                                                 uint32_t *__restrict__ sread, uint16_t *__restrict__ sk) {
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= M) return;
    const uint32_t p = sval[s], i = pread[p];
    sread[s] = i;
    sk[s] = (uint16_t)((uint64_t)p - doff[i]);
}

// ---- segmented scans --------------------------------------------------------------------------------------------
// (v, r): v = flags counted in the trailing segment of an interval, r = 1 if that segment starts inside the interval.
// Inclusive scan: out[t] = r[t] ? v[t] : v[t] + out[t-1].
__device__ __forceinline__ void sp_seg_warp_scan(uint32_t &v, uint32_t &r) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t ov = __shfl_up_sync(0xffffffffu, v, d), orr = __shfl_up_sync(0xffffffffu, r, d);
        if (lane_id() >= (uint32_t)d) { if (!r) v += ov; r |= orr; }
    }
}
// block-wide (THREADS a multiple of 32, <= 1024): returns the inclusive value of the calling thread given `cin` = value
// flowing in from before the block; *r_out = 1 if a segment start lies in [block start, this thread]. sm: 2 * 32 words.
template <int THREADS>
__device__ __forceinline__ uint32_t sp_seg_block_scan(uint32_t v, uint32_t r, uint32_t cin, uint32_t *sm, uint32_t *r_out) {
    const int w = threadIdx.x >> 5;
    constexpr int NW = THREADS / 32;
    sp_seg_warp_scan(v, r);
    if (lane_id() == 31) { sm[w] = v; sm[32 + w] = r; }
    __syncthreads();
    if (w == 0) {
        uint32_t wv = lane_id() < NW ? sm[lane_id()] : 0u, wr = lane_id() < NW ? sm[32 + lane_id()] : 0u;
        sp_seg_warp_scan(wv, wr);
        if (lane_id() < NW) { sm[lane_id()] = wv; sm[32 + lane_id()] = wr; }
    }
    __syncthreads();
    uint32_t pv = cin, pr = 0;                    // what flows into this warp
    if (w > 0) { pv = sm[w - 1]; pr = sm[32 + w - 1]; if (!pr) pv += cin; }
    if (!r) v += pv;
    r |= pr;
    if (r_out) *r_out = r;
    __syncthreads();
    return v;
}

struct SpRound {
    int64_t M;                       // pairs
    const uint32_t *sb, *sread, *sval;
    const uint16_t *sk;
    const uint16_t *sel;
    uint8_t *fbyte;                  // [ceil(M / 8)] flags of 8 consecutive sorted pairs
    uint32_t *tail, *treset;         // [tiles]
    const uint32_t *dirty;           // [nb1] round stamp: bucket b is dirty in round r iff dirty[b] == r
    uint32_t round;
    uint32_t all;                    // 1: first round over these pairs, every bucket counts as dirty
    uint32_t rank_mask;              // sb = input block << rank_bits | bucket rank
    uint8_t *tile_clean;             // [tiles] 1 = no pair of the tile belongs to a dirty bucket this round: nothing of the tile changes
    uint32_t *n_dirty_tiles;         // counter of the round
};

// flags of the current assignment + per tile: flags in the tile's trailing segment, and whether that segment starts
// inside the tile (or at its first pair)
__global__ void __launch_bounds__(kSpThreads) sp_flags_k(SpRound p) {
    __shared__ uint32_t sm[kSpThreads / 32];
    const int64_t tbeg = (int64_t)blockIdx.x * kSpTile;
    const int64_t tend = tbeg + kSpTile < p.M ? tbeg + kSpTile : p.M;
    const int64_t base = tbeg + (int64_t)threadIdx.x * kSpItems;
    uint32_t b[kSpItems];
#pragma unroll
    for (int j = 0; j < kSpItems; j++) b[j] = (base + j < p.M) ? p.sb[base + j] : 0xffffffffu;
    uint32_t dm0 = 0;
#pragma unroll
    for (int j = 0; j < kSpItems; j++) dm0 |= ((base + j < p.M) && (p.all || p.dirty[b[j] & p.rank_mask] == p.round)) ? (1u << j) : 0u;
    // a tile without a pair of a dirty bucket keeps its flags, its tail and (sp_counts_k) its counts: late rounds touch few tiles
    if (!__syncthreads_or(dm0 != 0)) {
        if (threadIdx.x == 0) p.tile_clean[blockIdx.x] = 1;
        return;
    }
    if (threadIdx.x == 0) { p.tile_clean[blockIdx.x] = 0; atomicAdd(p.n_dirty_tiles, 1u); }
    const uint32_t klast = p.sb[tend - 1];
    uint32_t rd[kSpItems], kk[kSpItems];
#pragma unroll
    for (int j = 0; j < kSpItems; j++) {
        const bool ok = ((dm0 >> j) & 1u) != 0;
        rd[j] = ok ? p.sread[base + j] : 0u;
        kk[j] = ok ? (uint32_t)p.sk[base + j] : 0x10000u;    // never equals a slot
    }
    // only pairs of dirty buckets look their read's selection up again (the random access of this kernel); the others keep
    // the flag of the round before
    const uint32_t dm = dm0;
    const uint32_t oldbits = (base < p.M && dm != 0xffu) ? (uint32_t)p.fbyte[base >> 3] : 0u;
    uint32_t s[kSpItems];
#pragma unroll
    for (int j = 0; j < kSpItems; j++) s[j] = ((dm >> j) & 1u) ? (uint32_t)p.sel[rd[j]] : 0x20000u;   // independent random 2-byte loads
    uint32_t bits = 0, cl = 0;
#pragma unroll
    for (int j = 0; j < kSpItems; j++) {
        const uint32_t f = ((dm >> j) & 1u) ? (s[j] == kk[j] ? 1u : 0u) : ((oldbits >> j) & 1u);
        bits |= f << j;
        cl += (f && b[j] == klast) ? 1u : 0u;
    }
    if (base < p.M) p.fbyte[base >> 3] = (uint8_t)bits;
    cl = __reduce_add_sync(0xffffffffu, cl);
    if (lane_id() == 0) sm[threadIdx.x >> 5] = cl;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int w = 0; w < kSpThreads / 32; w++) t += sm[w];
        p.tail[blockIdx.x] = t;
        const bool cont = p.sb[tbeg] == klast && tbeg > 0 && p.sb[tbeg - 1] == klast;   // the trailing segment began in an earlier tile
        p.treset[blockIdx.x] = cont ? 0u : 1u;
    }
}

// X[t] = flags of tile t's trailing segment counted from that segment's start (possibly many tiles back); one CTA
__global__ void __launch_bounds__(1024) sp_tilescan_k(const uint32_t *__restrict__ tail, const uint32_t *__restrict__ treset, int64_t tiles,
                                                      uint32_t *__restrict__ X) {
    __shared__ uint32_t sm[64];
    __shared__ uint32_t s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int64_t t0 = 0; t0 < tiles; t0 += 1024) {
        const int64_t t = t0 + threadIdx.x;
        const uint32_t v = t < tiles ? tail[t] : 0u, r = t < tiles ? treset[t] : 0u;
        const uint32_t cin = s_carry;
        const uint32_t y = sp_seg_block_scan<1024>(v, r, cin, sm, (uint32_t *)nullptr);
        if (t < tiles) X[t] = y;
        if (threadIdx.x == 1023) s_carry = y;   // padding elements (v = 0, r = 0) pass the value through
        __syncthreads();
    }
}

struct SpCounts {
    int64_t M;
    const uint32_t *sb, *sval;
    const uint8_t *fbyte;
    const uint32_t *X;               // tile scan
    const uint32_t *base;            // [nb1] populations before the local reads
    uint32_t *cnt;                   // [M] read-major: what pair p sees
    uint32_t *fold;                  // != null: instead of scattering counts, write base + segment total at every segment end
    const uint32_t *dirty; uint32_t round, all, rank_mask;   // counts are only scattered for dirty buckets (see SpRound)
    const uint8_t *tile_clean;               // tiles sp_flags_k found clean are skipped
    const uint32_t *sread; uint8_t *ractive; uint8_t stamp;   // reads that receive a new count are marked for sp_decide_k
};

__global__ void __launch_bounds__(kSpThreads) sp_counts_k(SpCounts p) {
    __shared__ uint32_t sm[64];
    if (!p.fold && p.tile_clean[blockIdx.x]) return;        // block-uniform, before any barrier
    const int64_t tbeg = (int64_t)blockIdx.x * kSpTile;
    const int64_t base = tbeg + (int64_t)threadIdx.x * kSpItems;
    uint32_t b[kSpItems + 1];
#pragma unroll
    for (int j = 0; j <= kSpItems; j++) b[j] = (base + j < p.M) ? p.sb[base + j] : 0xffffffffu;
    const uint32_t prevkey = base > 0 && base < p.M ? p.sb[base - 1] : 0xfffffffeu;       // differs from every key (keys < 2^31)
    const uint32_t bits = base < p.M ? (uint32_t)p.fbyte[base >> 3] : 0u;
    uint32_t ex[kSpItems];
    uint32_t run = 0, lead = 0xffu, any = 0;       // lead bit j: no segment start in [0, j]
#pragma unroll
    for (int j = 0; j < kSpItems; j++) {
        const bool start = (j == 0) ? (b[0] != prevkey) : (b[j] != b[j - 1]);
        if (start) { run = 0; any = 1; lead &= (1u << j) - 1u; }
        ex[j] = run;
        run += (bits >> j) & 1u;
    }
    uint32_t cin = 0;
    if (blockIdx.x > 0 && tbeg < p.M && p.sb[tbeg - 1] == p.sb[tbeg]) cin = p.X[blockIdx.x - 1];
    // inclusive over threads; the exclusive value of this thread = inclusive value of the thread before it
    const uint32_t y = sp_seg_block_scan<kSpThreads>(run, any, cin, sm, (uint32_t *)nullptr);
    __shared__ uint32_t ys[kSpThreads];
    ys[threadIdx.x] = y;
    __syncthreads();
    const uint32_t pre = threadIdx.x == 0 ? cin : ys[threadIdx.x - 1];
#pragma unroll
    for (int j = 0; j < kSpItems; j++) {
        if (base + j >= p.M) break;
        const uint32_t c = p.base[b[j] & p.rank_mask] + ex[j] + (((lead >> j) & 1u) ? pre : 0u);
        if (p.fold) {
            if (b[j + 1] != b[j]) p.fold[b[j] & p.rank_mask] = c + ((bits >> j) & 1u);    // last pair of its (block, bucket) segment: every bucket at most once per range
        } else if (p.all || p.dirty[b[j] & p.rank_mask] == p.round) {
            p.cnt[p.sval[base + j]] = c;
            if (p.ractive) p.ractive[p.sread[base + j]] = p.stamp;     // late rounds only: one more random store per pair does not pay while most tiles are dirty
        }
    }
}

// every read re-decides: first arg-max over its ordered candidates, strict > (reads.cpp:420-421).
// hist != null (sharded rounds): the local bucket histogram follows the decisions (full = count everything, else only changes)
__global__ void __launch_bounds__(256) sp_decide_k(int64_t n, const uint16_t *__restrict__ ncand, const uint64_t *__restrict__ doff,
                                                   const uint32_t *__restrict__ cnt, const uint64_t *__restrict__ cand_off,
                                                   const uint32_t *__restrict__ cand_rank, uint16_t *__restrict__ sel,
                                                   uint32_t *__restrict__ changed, uint32_t *__restrict__ hist, int full,
                                                   uint32_t *__restrict__ dirty, uint32_t next_round, const uint8_t *__restrict__ ractive, int stamp /* < 0: all reads */) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t chg = 0;
    // a read whose counts were not rewritten this round decides as before (a stale stamp from 256 rounds ago only costs a re-evaluation)
    if (i < n && (stamp < 0 || ractive[i] == (uint8_t)stamp)) {
        const int nc = ncand[i];
        if (nc > 0) {
            const uint64_t d = doff[i];
            uint32_t bc = cnt[d]; int bk = 0;
            for (int k = 1; k < nc; k++) {
                const uint32_t c = cnt[d + k];
                if (c > bc) { bc = c; bk = k; }
            }
            const int old = sel[i];
            const uint64_t src = cand_off[i];
            if (bk != old) {
                sel[i] = (uint16_t)bk; chg = 1;
                dirty[cand_rank[src + old]] = next_round;      // both buckets' flags changed (plain stores: every writer stores the same value)
                dirty[cand_rank[src + bk]] = next_round;
            }
            if (hist) {
                if (full) atomicAdd(&hist[cand_rank[src + bk]], 1u);
                else if (chg) { atomicSub(&hist[cand_rank[src + old]], 1u); atomicAdd(&hist[cand_rank[src + bk]], 1u); }
            }
        }
    }
    chg = __reduce_add_sync(0xffffffffu, chg);
    if (lane_id() == 0 && chg) atomicAdd(changed, chg);
}

// sharded rounds: the populations before the local reads change from round to round; a bucket whose value moved is dirty
__global__ void sp_mark_base_k(const uint32_t *__restrict__ base, uint32_t *__restrict__ base_prev, int nb1, uint32_t *__restrict__ dirty, uint32_t round) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nb1) return;
    const uint32_t v = base[i];
    if (v != base_prev[i]) { base_prev[i] = v; dirty[i] = round; }
}

__global__ void sp_copy_tot_k(const uint32_t *__restrict__ hist, const uint32_t *__restrict__ changed, int nb1, uint32_t *__restrict__ tot) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nb1) tot[i] = hist[i];
    if (i == nb1) tot[nb1] = *changed;
}

}  // namespace scb
