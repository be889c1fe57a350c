// emit_offsets.cuh - the output side's three prefix sums in one pass (default since the end of round 1: -0.36 ms at
// 50M x 150 bp; SCB_EMIT_FUSED_SCAN=0 selects the generic scans).
//
// emit_order (api.cu) needs, per emitted read p: the segment number (running count of (chunk, bucket) heads), the
// byte offset of its name record in stream 0 (names.cpp:48-62: length byte + name) and of its read record in
// stream 1 (reads.cpp:128-130: packed rotated read + end marker). The default path gathers the per-read metadata
// word into output order (gather_meta_k) and then runs three generic exclusive scans (prims.cuh: 3 kernels each,
// every input read twice). Here: reduce (with the gather fused in) -> one 3-row scan of the tile sums -> apply.
// Inside a tile of 2048 reads all three running sums fit in ONE u64 (12 + 21 + 31 bits), so a tile costs a single
// block scan; tile bases are kept unpacked in u64.
#pragma once
#include "common.cuh"
#include "emit2.cuh"
#include "prims.cuh"

namespace scb {

constexpr int kEoHeadBits = 12;   // <= 2048 heads per tile
constexpr int kEoNameBits = 21;   // <= 2048 * 256 name-record bytes per tile
constexpr int kEoReadShift = kEoHeadBits + kEoNameBits;   // <= 2048 * 514 read-record bytes per tile above this
static_assert(kScanTile <= 2048, "packed per-tile sums assume at most 2048 reads per tile");

__host__ __device__ __forceinline__ uint64_t eo_pack(uint32_t head, uint32_t namelen, uint32_t lvl, int L1, int sz_meta, int use_names) {
    const uint64_t nb = use_names ? (uint64_t)namelen + 1 : 0;
    const uint64_t rb = (uint64_t)(sz_read(L1 - (int)lvl) + sz_meta);
    return (uint64_t)head | (nb << kEoHeadBits) | (rb << kEoReadShift);
}
__host__ __device__ __forceinline__ uint32_t eo_heads(uint64_t v) { return (uint32_t)(v & ((1u << kEoHeadBits) - 1u)); }
__host__ __device__ __forceinline__ uint64_t eo_name_bytes(uint64_t v) { return (v >> kEoHeadBits) & ((1ull << kEoNameBits) - 1ull); }
__host__ __device__ __forceinline__ uint64_t eo_read_bytes(uint64_t v) { return v >> kEoReadShift; }

// pass 1: ms[p] = meta_in[perm[p]] (the one random small gather of the output side) + per-tile sums, rows [3][nt]
__global__ void __launch_bounds__(kScanThreads) emit_off_reduce_k(const uint64_t *__restrict__ meta_in, const uint32_t *__restrict__ perm, KeyHead kh,
                                                                  int64_t n, int L1, int sz_meta, int use_names, uint64_t *__restrict__ ms,
                                                                  uint64_t *__restrict__ ts, int64_t nt) {
    __shared__ uint64_t sm[kScanThreads / 32 + 1];
    const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
    uint32_t pi[kScanItems];
#pragma unroll
    for (int k = 0; k < kScanItems; k++) pi[k] = (base + k < n) ? perm[base + k] : 0u;
    uint64_t m[kScanItems];
#pragma unroll
    for (int k = 0; k < kScanItems; k++) m[k] = (base + k < n) ? (uint64_t)ldg_g64((const int64_t *)meta_in + pi[k]) : 0ull;
    uint64_t s = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; k++)
        if (base + k < n) {
            ms[base + k] = m[k];
            s += eo_pack(kh(base + k), (uint32_t)meta_namelen(m[k]), (uint32_t)meta_lvl(m[k]), L1, sz_meta, use_names);
        }
    uint64_t tot;
    block_excl_scan<uint64_t, kScanThreads>(s, sm, &tot);
    if (threadIdx.x == 0) {
        ts[blockIdx.x] = eo_heads(tot);
        ts[nt + blockIdx.x] = eo_name_bytes(tot);
        ts[2 * nt + blockIdx.x] = eo_read_bytes(tot);
    }
}

// pass 2: exclusive scan of each row of tile sums (one CTA per row); totals -> tot3[row]
__global__ void __launch_bounds__(1024) emit_off_sums_k(uint64_t *ts, int64_t nt, uint64_t *tot3) {
    __shared__ uint64_t sm[1024 / 32 + 1];
    __shared__ uint64_t carry;
    uint64_t *row = ts + (int64_t)blockIdx.x * nt;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int64_t b = 0; b < nt; b += 1024) {
        const int64_t i = b + threadIdx.x;
        const uint64_t v = i < nt ? row[i] : 0ull;
        uint64_t tot;
        const uint64_t ex = block_excl_scan<uint64_t, 1024>(v, sm, &tot);
        if (i < nt) row[i] = ex + carry;
        __syncthreads();
        if (threadIdx.x == 0) carry += tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) tot3[blockIdx.x] = carry;
}

// pass 3: tile-local packed exclusive scan + tile bases -> segment numbers and stream offsets; entry [n] = totals
__global__ void __launch_bounds__(kScanThreads) emit_off_apply_k(const uint64_t *__restrict__ ms, KeyHead kh, int64_t n, int L1, int sz_meta, int use_names,
                                                                 const uint64_t *__restrict__ ts, int64_t nt, const uint64_t *__restrict__ tot3,
                                                                 uint32_t *__restrict__ hsum, uint64_t *__restrict__ offN, uint64_t *__restrict__ offR) {
    __shared__ uint64_t sm[kScanThreads / 32 + 1];
    const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
    uint64_t v[kScanItems];
    uint64_t s = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; k++) {
        v[k] = 0;
        if (base + k < n) {
            const uint64_t m = ms[base + k];
            v[k] = eo_pack(kh(base + k), (uint32_t)meta_namelen(m), (uint32_t)meta_lvl(m), L1, sz_meta, use_names);
        }
        s += v[k];
    }
    uint64_t ex = block_excl_scan<uint64_t, kScanThreads>(s, sm, (uint64_t *)nullptr);
    const uint32_t bh = (uint32_t)ts[blockIdx.x];
    const uint64_t bn = ts[nt + blockIdx.x], br = ts[2 * nt + blockIdx.x];
#pragma unroll
    for (int k = 0; k < kScanItems; k++) {
        if (base + k < n) {
            hsum[base + k] = bh + eo_heads(ex);
            offN[base + k] = bn + eo_name_bytes(ex);
            offR[base + k] = br + eo_read_bytes(ex);
        }
        ex += v[k];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) { hsum[n] = (uint32_t)tot3[0]; offN[n] = tot3[1]; offR[n] = tot3[2]; }
}

}  // namespace scb
