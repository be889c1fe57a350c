// prims.cuh - device-wide building blocks written for this path: exclusive scan over a functor,
// stable LSD radix sort of (u64 key, u32 value) pairs, stream compaction helpers.
// No CUB/Thrust: these are part of the hot path (histogram + stable counting-sort scatter).
#pragma once
#include "common.cuh"

namespace scb {

// =================================================================================================
// exclusive scan  out[i] = sum_{j<i} f(j)     (3 kernels: tile reduce, scan of tile sums, apply)
// =================================================================================================
constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;

template <typename T>
__device__ __forceinline__ T warp_incl_scan(T v) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        T o = __shfl_up_sync(0xffffffffu, v, d);
        if (lane_id() >= (uint32_t)d) v += o;
    }
    return v;
}

// block-wide exclusive scan of one value per thread; returns exclusive prefix, total in *total
template <typename T, int THREADS>
__device__ __forceinline__ T block_excl_scan(T v, T *smem /* THREADS/32 + 1 */, T *total) {
    T inc = warp_incl_scan(v);
    const int w = threadIdx.x >> 5;
    if (lane_id() == 31) smem[w] = inc;
    __syncthreads();
    if (w == 0) {
        T s = (lane_id() < THREADS / 32) ? smem[lane_id()] : T(0);
        T si = warp_incl_scan(s);
        if (lane_id() < THREADS / 32) smem[lane_id()] = si - s;
        if (lane_id() == THREADS / 32 - 1) smem[THREADS / 32] = si;
    }
    __syncthreads();
    T r = inc - v + smem[w];
    if (total) *total = smem[THREADS / 32];
    __syncthreads();
    return r;
}

template <typename T, typename F>
__global__ void __launch_bounds__(kScanThreads) scan_reduce_k(F f, int64_t n, T *tile_sums) {
    __shared__ T sm[kScanThreads / 32 + 1];
    int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
    T s = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; k++)
        if (base + k < n) s += f(base + k);
    T tot;
    block_excl_scan<T, kScanThreads>(s, sm, &tot);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = tot;
}

template <typename T>
__global__ void __launch_bounds__(1024) scan_sums_k(T *tile_sums, int64_t nt, T *total_out, T init) {
    __shared__ T sm[1024 / 32 + 1];
    __shared__ T carry;
    if (threadIdx.x == 0) carry = init;
    __syncthreads();
    for (int64_t b = 0; b < nt; b += 1024) {
        int64_t i = b + threadIdx.x;
        T v = i < nt ? tile_sums[i] : T(0);
        T tot;
        T ex = block_excl_scan<T, 1024>(v, sm, &tot);
        if (i < nt) tile_sums[i] = ex + carry;
        __syncthreads();
        if (threadIdx.x == 0) carry += tot;
        __syncthreads();
    }
    if (threadIdx.x == 0 && total_out) *total_out = carry;
}

template <typename T, typename F>
__global__ void __launch_bounds__(kScanThreads) scan_apply_k(F f, int64_t n, const T *tile_sums, T *out) {
    __shared__ T sm[kScanThreads / 32 + 1];
    int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
    T v[kScanItems];
    T s = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; k++) {
        v[k] = (base + k < n) ? f(base + k) : T(0);
        s += v[k];
    }
    T ex = block_excl_scan<T, kScanThreads>(s, sm, (T *)nullptr) + tile_sums[blockIdx.x];
#pragma unroll
    for (int k = 0; k < kScanItems; k++) {
        if (base + k < n) out[base + k] = ex;
        ex += v[k];
    }
}

inline int64_t scan_tiles(int64_t n) { return cdiv(n > 0 ? n : 1, kScanTile); }

// out[0..n) exclusive; total (device pointer, may be null) receives init + sum. tile_ws: scan_tiles(n) T's.
template <typename T, typename F>
void exclusive_scan(F f, int64_t n, T *out, T *total_dev, T *tile_ws, cudaStream_t st, T init = T(0)) {
    int64_t nt = scan_tiles(n);
    SCB_LAUNCH((scan_reduce_k<T, F>), (unsigned)nt, kScanThreads, 0, st, f, n, tile_ws);
    SCB_LAUNCH((scan_sums_k<T>), 1, 1024, 0, st, tile_ws, nt, total_dev, init);
    if (n > 0) SCB_LAUNCH((scan_apply_k<T, F>), (unsigned)nt, kScanThreads, 0, st, f, n, tile_ws, out);
}

template <typename TIn, typename TOut>
struct LoadAs {
    const TIn *p;
    __device__ __forceinline__ TOut operator()(int64_t i) const { return (TOut)p[i]; }
};

// =================================================================================================
// stable LSD radix sort of (u64 or u32 key, u32 val), 8-bit digits over key bits [bit_lo, bit_hi)
//   pass = digit histogram per tile -> exclusive scan (digit-major) -> stable scatter
// Each warp owns a contiguous run of kSortItems*32 elements and ranks them 32 at a time with
// match.any, so order inside a digit is input order (stability).
// =================================================================================================
constexpr int kSortThreads = 256;
constexpr int kSortWarps = kSortThreads / 32;
#ifndef SCB_SORT_ITEMS
#define SCB_SORT_ITEMS 12
#endif
constexpr int kSortItems = SCB_SORT_ITEMS;   // tile = 3072 elements: 36 KB of staging + 10 KB of counters stays under 48 KB static smem
constexpr int kSortTile = kSortThreads * kSortItems;

template <typename K>
__global__ void __launch_bounds__(kSortThreads) sort_hist_k(const K *keys, int64_t n, int shift, uint32_t mask,
                                                            uint32_t *hist /*[256][tiles]*/, int64_t tiles) {
    __shared__ uint32_t h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    int64_t base = (int64_t)blockIdx.x * kSortTile;
#pragma unroll 4
    for (int k = 0; k < kSortItems; k++) {
        int64_t i = base + (int64_t)k * kSortThreads + threadIdx.x;
        if (i < n) atomicAdd(&h[(uint32_t)(keys[i] >> shift) & mask], 1u);
    }
    __syncthreads();
    hist[(int64_t)threadIdx.x * tiles + blockIdx.x] = h[threadIdx.x];
}

// Scatter pass. Ranks are computed per warp with match.any (stable), the tile is first reordered by
// digit in shared memory, then written out so that each digit's elements leave as one contiguous run.
#ifndef SCB_SORT_MINBLOCKS
#define SCB_SORT_MINBLOCKS 4   // 64 registers (8 bytes of spill), 4 CTAs per SM: sort 3.94 -> 3.44 ms at 50M x 150 (3 CTAs at 80 registers before; 8-item tiles with 4 or 5 CTAs: 3.72 / 3.64)
#endif
template <typename K>
__global__ void __launch_bounds__(kSortThreads, SCB_SORT_MINBLOCKS) sort_scatter_k(const K *__restrict__ keys, const uint32_t *__restrict__ vals,
                                                               K *__restrict__ keys_out, uint32_t *__restrict__ vals_out, int64_t n,
                                                               int shift, uint32_t mask, const uint32_t *__restrict__ hist_scanned, int64_t tiles) {
    __shared__ uint32_t cnt[kSortWarps][256];
    __shared__ uint32_t gbase[256];     // global start of this tile's run of digit d, minus its start inside the tile
    __shared__ uint32_t tbase[256];     // start of digit d inside the digit-ordered tile
    __shared__ uint32_t sc[kSortThreads / 32 + 1];
    __shared__ K skey[kSortTile];
    __shared__ uint32_t sval[kSortTile];
    const int w = threadIdx.x >> 5, l = lane_id();
    for (int d = l; d < 256; d += 32) cnt[w][d] = 0;
    __syncwarp();
    const int64_t tbeg = (int64_t)blockIdx.x * kSortTile;
    const int64_t wbase = tbeg + (int64_t)w * (kSortItems * 32);
    K key[kSortItems];
    uint32_t rank[kSortItems];
#pragma unroll
    for (int k = 0; k < kSortItems; k++) {
        int64_t i = wbase + k * 32 + l;
        key[k] = i < n ? keys[i] : (K)~(K)0;
    }
#pragma unroll
    for (int k = 0; k < kSortItems; k++) {
        int64_t i = wbase + k * 32 + l;
        bool ok = i < n;
        uint32_t d = ok ? ((uint32_t)(key[k] >> shift) & mask) : 0xffffffffu;  // invalid lanes never match valid ones
        uint32_t m = __match_any_sync(0xffffffffu, d);
        uint32_t leader = __ffs(m) - 1;
        uint32_t b = 0;
        if (ok && l == leader) {
            b = cnt[w][d];
            cnt[w][d] = b + __popc(m);
        }
        b = __shfl_sync(0xffffffffu, b, leader);
        rank[k] = b + __popc(m & lanemask_lt());
        __syncwarp();
    }
    __syncthreads();
    {   // digit = threadIdx.x: exclusive scan over warps, then over digits
        uint32_t run = 0;
#pragma unroll
        for (int ww = 0; ww < kSortWarps; ww++) {
            uint32_t c = cnt[ww][threadIdx.x];
            cnt[ww][threadIdx.x] = run;
            run += c;
        }
        uint32_t tb = block_excl_scan<uint32_t, kSortThreads>(run, sc, (uint32_t *)nullptr);
        tbase[threadIdx.x] = tb;
        gbase[threadIdx.x] = hist_scanned[(int64_t)threadIdx.x * tiles + blockIdx.x] - tb;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kSortItems; k++) {
        int64_t i = wbase + k * 32 + l;
        if (i < n) {
            uint32_t d = (uint32_t)(key[k] >> shift) & mask;
            uint32_t t = tbase[d] + cnt[w][d] + rank[k];
            skey[t] = key[k];
            sval[t] = vals[i];          // loaded late: keeps the values out of the registers while ranks are computed
        }
    }
    __syncthreads();
    const int64_t rem = n - tbeg;
    const int cntTile = (int)(rem < (int64_t)kSortTile ? rem : (int64_t)kSortTile);
    for (int t = threadIdx.x; t < cntTile; t += kSortThreads) {
        const K kk = skey[t];
        uint32_t d = (uint32_t)(kk >> shift) & mask;
        uint32_t p = gbase[d] + (uint32_t)t;
        keys_out[p] = kk;
        vals_out[p] = sval[t];
    }
}

struct SortWs {
    uint32_t *hist = nullptr;      // 256 * tiles
    uint32_t *tile_ws = nullptr;   // scan_tiles(256 * tiles)
    static int64_t hist_elems(int64_t n) { return 256 * cdiv(n > 0 ? n : 1, kSortTile); }
};

// Sorts in place logically: on return *keys / *vals point at the buffer holding the result
// (ping-pong with *keys_alt / *vals_alt).
template <typename K>
inline void radix_sort_pairs(K **keys, uint32_t **vals, K **keys_alt, uint32_t **vals_alt, int64_t n,
                             int bit_lo, int bit_hi, const SortWs &ws, cudaStream_t st) {
    if (n <= 1) return;
    int64_t tiles = cdiv(n, kSortTile);
    for (int lo = bit_lo; lo < bit_hi; lo += 8) {
        int bits = bit_hi - lo < 8 ? bit_hi - lo : 8;
        uint32_t mask = (1u << bits) - 1u;
        SCB_LAUNCH(sort_hist_k<K>, (unsigned)tiles, kSortThreads, 0, st, *keys, n, lo, mask, ws.hist, tiles);
        exclusive_scan<uint32_t>(LoadAs<uint32_t, uint32_t>{ws.hist}, 256 * tiles, ws.hist, (uint32_t *)nullptr, ws.tile_ws, st);
        SCB_LAUNCH(sort_scatter_k<K>, (unsigned)tiles, kSortThreads, 0, st, *keys, *vals, *keys_alt, *vals_alt, n, lo, mask,
                   ws.hist, tiles);
        K *tk = *keys; *keys = *keys_alt; *keys_alt = tk;
        uint32_t *tv = *vals; *vals = *vals_alt; *vals_alt = tv;
    }
}

}  // namespace scb
