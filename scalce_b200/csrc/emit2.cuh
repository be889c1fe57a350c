// emit2.cuh - output side of the transform on 2-bit packed reads.
//   packed row: PW = ceil(L/16) u32 words per read, base 16k+j of word k at bits 31-2j (MSB first),
//   zero filled past L - written by the scan kernel (or pack_reads_k on the fallback path).
//   keys      : prefix of the in-bucket sort key straight from the packed row
//   segments  : runs of equal (chunk, bucket) in the sorted order -> per-segment table
//   emit      : 16 lanes per read: names, rotated packed read + end marker, qualities (word copies
//               with funnel-shifted sources), mate 2
//   meta      : one record per segment (reads.cpp:160-176)
#pragma once
#include "common.cuh"
#include "pipeline.cuh"
#include "scan_smem.cuh"

namespace scb {

__device__ __forceinline__ uint32_t pk_word(const uint32_t *__restrict__ row, int k, int PW) { return (k >= 0 && k < PW) ? row[k] : 0u; }
// 32 bases (64 bits, MSB first) starting at base offset b0; zero past the row
__device__ __forceinline__ uint64_t pk_bits64(const uint32_t *__restrict__ row, int PW, int b0) {
    const int k = b0 >> 4, sh = (b0 & 15) * 2;
    const uint32_t w0 = pk_word(row, k, PW), w1 = pk_word(row, k + 1, PW), w2 = pk_word(row, k + 2, PW);
    const uint32_t hi = __funnelshift_l(w1, w0, sh), lo = __funnelshift_l(w2, w1, sh);
    return ((uint64_t)hi << 32) | lo;
}
// 16 bases (32 bits, MSB first) starting at base offset b0 >= 0; unguarded: the caller masks what lies
// past the logical end, and every row is followed by readable words
__device__ __forceinline__ uint32_t pk_bits32(const uint32_t *__restrict__ row, int b0) {
    const int k = b0 >> 4, sh = (b0 & 15) * 2;
    return __funnelshift_l(ldg_g64(row + k + 1), ldg_g64(row + k), sh);
}
// `nbases` (1..32) bases of the in-bucket sort key of a read from key offset `from`:
// key = s[end..L) right-padded with A (reads.cpp:547-559)
__device__ __forceinline__ uint64_t pk_key_bits(const uint32_t *__restrict__ row, int PW, int end, int from, int nbases) {
    return pk_bits64(row, PW, end + from) >> (64 - 2 * nbases);
}

// fallback packer (global-table scan path): one thread per (read, word)
__global__ void pack_reads_k(const uint8_t *__restrict__ seq, int64_t n, int L, int PW, uint32_t *__restrict__ packed) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * PW) return;
    int64_t i = t / PW;
    int k = (int)(t - i * PW);
    const uint8_t *s = seq + i * (int64_t)L;
    uint32_t w = 0;
#pragma unroll
    for (int j = 0; j < 16; j++) {
        int q = 16 * k + j;
        w = (w << 2) | (q < L ? base_code(s[q]) : 0u);
    }
    packed[t] = w;
}

__global__ void build_keys_pk_k(const uint32_t *__restrict__ packed, int PW, int64_t n, const uint32_t *__restrict__ asg,
                                const uint16_t *__restrict__ endv, const uint32_t *__restrict__ chunk, int nb, int root_pos,
                                int seg_bits, int pb, uint64_t *__restrict__ keys, uint32_t *__restrict__ vals) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t seg = (uint64_t)(chunk ? chunk[i] : 0u) * (uint64_t)(nb + 1) + bucket_ord(asg[i], nb, root_pos);
    uint64_t kb = pk_key_bits(packed + i * (int64_t)PW, PW, endv[i], 0, pb);
    keys[i] = (seg_bits ? (seg << (64 - seg_bits)) : 0ull) | (kb << (64 - seg_bits - 2 * pb));
    vals[i] = (uint32_t)i;
}

__global__ void tie_rekey_pk_k(const uint32_t *__restrict__ packed, int PW, const uint16_t *__restrict__ endv,
                               const uint32_t *__restrict__ c_idx, const uint32_t *__restrict__ c_grp, int64_t m, int grp_bits,
                               int from, int nbases, uint64_t *__restrict__ keys, uint32_t *__restrict__ vals) {
    int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= m) return;
    uint32_t i = c_idx[c];
    uint64_t kb = pk_key_bits(packed + (int64_t)i * PW, PW, endv[i], from, nbases);
    keys[c] = (grp_bits ? ((uint64_t)c_grp[c] << (64 - grp_bits)) : 0ull) | (kb << (64 - grp_bits - 2 * nbases));
    vals[c] = i;
}

// ---- tie groups of up to 32 elements: one warp orders a whole group on the full remaining key -----------------
// The compact tie list is in position order, so a group (run of equal sorted keys) is a contiguous range of it and
// its members are already in input-index order. Each lane holds one member; 32 key bases at a time every lane
// compares its chunk with every other member's (shuffles) until all pairs are decided or the key is exhausted;
// fully equal keys keep input order (the reference's radix sort is stable, reads.cpp:571-587).
__global__ void tie_group_starts_k(const uint32_t *__restrict__ c_grp, int64_t t, uint32_t G, uint32_t *__restrict__ gstart) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c == 0) gstart[G] = (uint32_t)t;
    if (c >= t) return;
    if (c == 0 || c_grp[c] != c_grp[c - 1]) gstart[c_grp[c]] = (uint32_t)c;
}
__global__ void __launch_bounds__(256) tie_small_groups_k(const uint32_t *__restrict__ packed, int PW, int L1, const uint16_t *__restrict__ endv,
                                                          const uint32_t *__restrict__ gstart, uint32_t G, const uint32_t *__restrict__ c_pos,
                                                          const uint32_t *__restrict__ c_idx, int consumed, uint32_t *__restrict__ perm,
                                                          uint8_t *__restrict__ big /* [t]: 1 = member of a group left to the iterative rounds */,
                                                          uint32_t *__restrict__ mid_list, uint32_t *__restrict__ mid_count, uint32_t mid_max) {
    const uint32_t g = (uint32_t)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (g >= G) return;
    const uint32_t l = lane_id();
    const uint32_t gs = gstart[g], size = gstart[g + 1] - gs;
    if (size > 32) {
        // 33..mid_max members: queued for one CTA each (tie_mid_groups_k); beyond that: iterative radix rounds
        const uint8_t b = size > mid_max ? 1 : 0;
        for (uint32_t k = l; k < size; k += 32) big[gs + k] = b;
        if (!b && l == 0) mid_list[atomicAdd(mid_count, 1u)] = g;
        return;
    }
    const bool on = l < size;
    const uint32_t idx = on ? c_idx[gs + l] : 0u, pos = on ? c_pos[gs + l] : 0u;
    if (on) big[gs + l] = 0;
    const uint32_t *row = packed + (int64_t)idx * PW;
    const int end = on ? (int)endv[idx] : 0;
    const uint32_t members = size == 32 ? 0xffffffffu : ((1u << size) - 1u);
    uint32_t undecided = on ? (members & ~(1u << l)) : 0u, less = 0;   // less: members that sort before this lane's
    for (int from = consumed; from < L1; from += 32) {
        if (!__any_sync(0xffffffffu, undecided != 0)) break;
        const uint64_t mine = on ? pk_key_bits(row, PW, end, from, 32) : 0ull;
        for (uint32_t k = 0; k < size; k++) {
            const uint64_t other = __shfl_sync(0xffffffffu, mine, (int)k);
            if (undecided & (1u << k)) {
                if (other < mine) { less |= 1u << k; undecided &= ~(1u << k); }
                else if (other > mine) undecided &= ~(1u << k);
            }
        }
    }
    less |= undecided & lanemask_lt();                 // equal keys: input order
    const uint32_t rank = __popc(less);
    const uint32_t dst = __shfl_sync(0xffffffffu, pos, (int)rank);   // the group's positions are ascending along the lanes
    if (on) perm[dst] = idx;
}
// groups of 33..kTieMidMax members: one CTA refines the whole group in shared memory, 32 key bases per round:
// rank = number of members with a smaller (label, chunk) plus equal ones that come earlier; label = the rank of the
// first equal member, so members that are still tied share a label in the next round.
constexpr int kTieMidMax = 1024;
__global__ void __launch_bounds__(256) tie_mid_groups_k(const uint32_t *__restrict__ packed, int PW, int L1, const uint16_t *__restrict__ endv,
                                                        const uint32_t *__restrict__ gstart, const uint32_t *__restrict__ mid_list,
                                                        const uint32_t *__restrict__ mid_count, const uint32_t *__restrict__ c_pos,
                                                        const uint32_t *__restrict__ c_idx, int consumed, uint32_t *__restrict__ perm) {
    __shared__ uint64_t s_key[kTieMidMax];
    __shared__ uint32_t s_lab[kTieMidMax], s_new[kTieMidMax];
    __shared__ int s_tied;
    constexpr int PER = kTieMidMax / 256;
    const uint32_t count = *mid_count;
    for (uint32_t q = blockIdx.x; q < count; q += gridDim.x) {
        const uint32_t g = mid_list[q];
        const uint32_t gs = gstart[g], size = gstart[g + 1] - gs;
        uint32_t idx[PER], rank[PER];
        int end[PER];
#pragma unroll
        for (int u = 0; u < PER; u++) {
            const uint32_t j = threadIdx.x + 256 * u;
            idx[u] = j < size ? c_idx[gs + j] : 0u;
            end[u] = j < size ? (int)endv[idx[u]] : 0;
            rank[u] = j;
            if (j < size) s_lab[j] = 0;
        }
        for (int from = consumed; from < L1; from += 32) {
            if (threadIdx.x == 0) s_tied = 0;
#pragma unroll
            for (int u = 0; u < PER; u++) {
                const uint32_t j = threadIdx.x + 256 * u;
                if (j < size) s_key[j] = pk_key_bits(packed + (int64_t)idx[u] * PW, PW, end[u], from, 32);
            }
            __syncthreads();
            bool tied = false;
#pragma unroll
            for (int u = 0; u < PER; u++) {
                const uint32_t j = threadIdx.x + 256 * u;
                if (j >= size) continue;
                const uint32_t lj = s_lab[j];
                const uint64_t kj = s_key[j];
                uint32_t less = 0, eqb = 0, eqa = 0;
                for (uint32_t k = 0; k < size; k++) {
                    const uint32_t lk = s_lab[k];
                    const uint64_t kk = s_key[k];
                    const bool eq = lk == lj && kk == kj;
                    less += (lk < lj || (lk == lj && kk < kj)) ? 1u : 0u;
                    eqb += (eq && k < j) ? 1u : 0u;
                    eqa += (eq && k > j) ? 1u : 0u;
                }
                s_new[j] = less;
                rank[u] = less + eqb;
                tied |= (eqb + eqa) != 0;
            }
            if (tied) s_tied = 1;
            __syncthreads();
#pragma unroll
            for (int u = 0; u < PER; u++) {
                const uint32_t j = threadIdx.x + 256 * u;
                if (j < size) s_lab[j] = s_new[j];
            }
            const bool more = s_tied != 0;
            __syncthreads();
            if (!more) break;
        }
#pragma unroll
        for (int u = 0; u < PER; u++) {
            const uint32_t j = threadIdx.x + 256 * u;
            if (j < size) perm[c_pos[gs + rank[u]]] = idx[u];   // the group's positions are ascending along the compact list
        }
        __syncthreads();
    }
}

// what is left for the iterative rounds: members of large groups, keyed by their group number
__global__ void tie_big_compact_k(const uint8_t *__restrict__ big, const uint32_t *__restrict__ bpos, const uint32_t *__restrict__ c_pos,
                                  const uint32_t *__restrict__ c_idx, const uint32_t *__restrict__ c_grp, int64_t t,
                                  uint64_t *__restrict__ keys, uint32_t *__restrict__ idx, uint32_t *__restrict__ pos) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= t || !big[c]) return;
    const uint32_t o = bpos[c];
    keys[o] = (uint64_t)c_grp[c];
    idx[o] = c_idx[c];
    pos[o] = c_pos[c];
}

// ---- segments ----------------------------------------------------------------------------------------
struct KeyHead {   // 1 where the segment id (top `bits` bits of the sorted key, 0 bits = one segment) changes
    const uint64_t *keys; int shift, bits;
    __device__ __forceinline__ uint32_t operator()(int64_t p) const {
        if (p == 0) return 1u;
        if (bits == 0) return 0u;
        return ((keys[p] >> shift) != (keys[p - 1] >> shift)) ? 1u : 0u;
    }
};
struct SegTab {
    uint32_t *pos;      // [n_seg+1] first output position (pos[n_seg] = n)
    uint32_t *rank;     // [n_seg] bucket rank (nb = root)
    uint32_t *chunk;    // [n_seg]
    uint32_t *recsz;    // [n_seg] bytes per read in stream 1
};
__global__ void seg_table_k(KeyHead kh, const uint32_t *__restrict__ hsum, int64_t n, const uint32_t *__restrict__ perm,
                            const uint32_t *__restrict__ asg, const uint32_t *__restrict__ chunk, const uint8_t *__restrict__ rank_level,
                            int nb, int L1, int sz_meta, SegTab t, uint32_t n_seg) {
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p == 0) t.pos[n_seg] = (uint32_t)n;
    if (p >= n || !kh(p)) return;
    uint32_t m = hsum[p];
    uint32_t i = perm[p];
    uint32_t r = asg[i];
    t.pos[m] = (uint32_t)p;
    t.rank[m] = r;
    t.chunk[m] = chunk ? chunk[i] : 0u;
    int lv = r == (uint32_t)nb ? 0 : rank_level[r];
    t.recsz[m] = (uint32_t)(sz_read(L1 - lv) + sz_meta);
}
// ---- emit ---------------------------------------------------------------------------------------------
// ---- stream 2 / 5: qualities are a pure row gather: out row p <- in row perm[p] -----------------------
// One thread per 16 output bytes (the output is dense, so stores are aligned STG.128 and fully
// coalesced); a chunk may straddle two rows. Sources are unaligned: two aligned 16-byte loads and a
// byte shift. Requires L >= 16.
// 16 bytes starting `sh` (0..15) bytes into the 32-byte window (v0, v1): word select in two predicated levels,
// then four funnel shifts - branch free, so lanes of a warp that copy different rows do not diverge
__device__ __forceinline__ uint4 window16(uint4 v0, uint4 v1, int sh) {
    const bool b0 = (sh & 4) != 0, b1 = (sh & 8) != 0;
    // level 1: skip one word if bit 2 of sh is set
    const uint32_t u0 = b0 ? v0.y : v0.x, u1 = b0 ? v0.z : v0.y, u2 = b0 ? v0.w : v0.z, u3 = b0 ? v1.x : v0.w,
                   u4 = b0 ? v1.y : v1.x, u5 = b0 ? v1.z : v1.y, u6 = b0 ? v1.w : v1.z;
    // level 2: skip two words if bit 3 is set
    const uint32_t w0 = b1 ? u2 : u0, w1 = b1 ? u3 : u1, w2 = b1 ? u4 : u2, w3 = b1 ? u5 : u3, w4 = b1 ? u6 : u4;
    const int bs = (sh & 3) * 8;
    return make_uint4(__funnelshift_r(w0, w1, bs), __funnelshift_r(w1, w2, bs), __funnelshift_r(w2, w3, bs), __funnelshift_r(w3, w4, bs));
}
__device__ __forceinline__ uint4 load16_unaligned(const uint8_t *__restrict__ a, int nbytes) {
    const uintptr_t A = (uintptr_t)a;
    const uint4 *a0 = (const uint4 *)(A & ~(uintptr_t)15);
    const int sh = (int)(A & 15);
    uint4 v0 = ldg_g64(a0), v1 = make_uint4(0, 0, 0, 0);
    if (sh + nbytes > 16) v1 = ldg_g64(a0 + 1);               // only when it holds a needed byte (never past the buffer's last granule)
    return window16(v0, v1, sh);
}
// dst bytes [0, k) from x, bytes [k, 16) from the first 16-k bytes of y (k in 1..15): y is moved up by k bytes
// with the same window trick (32-byte window = 16 zero bytes, then y; start 16-k), then merged under a byte mask
__device__ __forceinline__ uint4 splice16(uint4 x, uint4 y, int k) {
    const uint4 ys = window16(make_uint4(0, 0, 0, 0), y, 16 - k);
    // mask of the bytes taken from x: the low k bytes of the 16
    const int kb = k * 8;                                       // 8..120 bits
    const uint32_t m0 = kb >= 32 ? 0xffffffffu : (0xffffffffu >> (32 - kb));
    const uint32_t m1 = kb >= 64 ? 0xffffffffu : (kb <= 32 ? 0u : (0xffffffffu >> (64 - kb)));
    const uint32_t m2 = kb >= 96 ? 0xffffffffu : (kb <= 64 ? 0u : (0xffffffffu >> (96 - kb)));
    const uint32_t m3 = kb <= 96 ? 0u : (0xffffffffu >> (128 - kb));
    return make_uint4((x.x & m0) | (ys.x & ~m0), (x.y & m1) | (ys.y & ~m1), (x.z & m2) | (ys.z & ~m2), (x.w & m3) | (ys.w & ~m3));
}
// A logical row array whose rows [lo, hi) live somewhere else: the sharded run with chunk ownership leaves a rank's OWN quality /
// mate-2 rows in its input arrays (no self-copy into the receive arrays) and only receives the other ranks' rows around them.
// `own` is pre-offset so that row r of that range is at own + r * L; lo == hi: one plain array.
struct RowSrc {
    const uint8_t *p, *own;
    int64_t lo, hi;
    // SEG = false: the caller knows lo == hi (one GPU, bucket ranges) and the two compares per row lookup are compiled out
    // (they cost the one-GPU emit 0.15 ms of 9.85 when they were unconditional)
    template <bool SEG>
    __host__ __device__ __forceinline__ const uint8_t *row(int64_t r, int L) const { return ((SEG && r >= lo && r < hi) ? own : p) + r * (int64_t)L; }
    bool split() const { return hi > lo; }
};
#ifndef SCB_GATHER_CHUNKS
#define SCB_GATHER_CHUNKS 3
#endif
constexpr int kGatherChunks = SCB_GATHER_CHUNKS;   // independent 16-byte chunks per thread; 3 measured best (4: 11.7 ms emit, 3: 10.6, 2: 10.8)   // independent 16-byte chunks per thread (memory-level parallelism)
template <bool SEG>
__global__ void __launch_bounds__(256) gather_rows16_k(const RowSrc src, uint8_t *__restrict__ dst,
                                                       const uint32_t *__restrict__ perm, int64_t n, int L) {
    const int64_t total = n * (int64_t)L;
    // one 64-bit division per thread (the block's first byte), 32-bit arithmetic after that
    const int64_t blk0 = (int64_t)blockIdx.x * (256 * kGatherChunks * 16);
    const int64_t pblk = blk0 / L;
    const uint32_t rblk = (uint32_t)(blk0 - pblk * L);
    const uint8_t *a0[kGatherChunks], *a1[kGatherChunks];
    int n0[kGatherChunks];
#pragma unroll
    for (int q = 0; q < kGatherChunks; q++) {     // all row lookups first ...
        const uint32_t lo = (uint32_t)(q * 256 + threadIdx.x) << 4;   // byte offset inside the block's range
        a0[q] = a1[q] = nullptr; n0[q] = 16;
        if (blk0 + lo < total) {
            const uint32_t x = rblk + lo, dp = x / (uint32_t)L;
            const int64_t p0 = pblk + dp;
            const int r0 = (int)(x - dp * (uint32_t)L);
            n0[q] = min(16, L - r0);
            a0[q] = src.row<SEG>(perm[p0], L) + r0;
            if (n0[q] < 16 && p0 + 1 < n) a1[q] = src.row<SEG>(perm[p0 + 1], L);
        }
    }
    uint4 v[kGatherChunks];
#pragma unroll
    for (int q = 0; q < kGatherChunks; q++)       // ... then all row loads ...
        if (a0[q]) {
            v[q] = load16_unaligned(a0[q], n0[q]);
            if (a1[q]) v[q] = splice16(v[q], load16_unaligned(a1[q], 16 - n0[q]), n0[q]);
        }
#pragma unroll
    for (int q = 0; q < kGatherChunks; q++) {     // ... then the stores
        const int64_t o = blk0 + ((int64_t)(q * 256 + threadIdx.x) << 4);
        if (!a0[q]) continue;
        if (o + 16 <= total) *(uint4 *)(dst + o) = v[q];        // dst is 16-byte aligned (allocation base)
        else {
            const uint32_t w[4] = {v[q].x, v[q].y, v[q].z, v[q].w};
            for (int k = 0; k < (int)(total - o); k++) dst[o + k] = (uint8_t)(w[k >> 2] >> (8 * (k & 3)));
        }
    }
}
__global__ void gather_rows_small_k(const RowSrc src, uint8_t *__restrict__ dst, const uint32_t *__restrict__ perm,
                                    int64_t n, int L) {
    const int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= n * (int64_t)L) return;
    const int64_t p = o / L;
    dst[o] = src.row<true>(perm[p], L)[o - p * L];
}

// ---- per-read metadata word ---------------------------------------------------------------------------
// Everything small the output side needs about a read, in one u64 so that output order costs ONE random
// 8-byte gather per read: bits 0-35 name offset (64 GiB of names per flush), 36-43 name length, 44-51 core level,
// 52-63 end marker (reads up to the reference's line cap, const.h:87).
__host__ __device__ __forceinline__ uint64_t meta_pack(uint64_t name_off, uint32_t namelen, uint32_t lvl, uint32_t end) {
    return (name_off & ((1ull << 36) - 1)) | ((uint64_t)(namelen & 0xffu) << 36) | ((uint64_t)(lvl & 0xffu) << 44) | ((uint64_t)(end & 0xfffu) << 52);
}
__device__ __forceinline__ int64_t meta_name_off(uint64_t m) { return (int64_t)(m & ((1ull << 36) - 1)); }
__device__ __forceinline__ int meta_namelen(uint64_t m) { return (int)((m >> 36) & 0xffu); }
__device__ __forceinline__ int meta_lvl(uint64_t m) { return (int)((m >> 44) & 0xffu); }
__device__ __forceinline__ int meta_end(uint64_t m) { return (int)(m >> 52); }

__global__ void build_meta_k(int64_t n, const int64_t *__restrict__ name_off, const uint8_t *__restrict__ lvl,
                             const uint16_t *__restrict__ endv, uint64_t *__restrict__ meta_in) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t a = 0; uint32_t nl = 0;
    if (name_off) { a = (uint64_t)name_off[i]; nl = (uint32_t)(name_off[i + 1] - name_off[i]); }
    meta_in[i] = meta_pack(a, nl, lvl[i], endv[i]);
}
__global__ void gather_meta_k(const uint64_t *__restrict__ meta_in, const uint32_t *__restrict__ perm, int64_t n,
                              uint64_t *__restrict__ meta_s) {
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p < n) meta_s[p] = (uint64_t)ldg_g64((const int64_t *)meta_in + perm[p]);
}
struct NameRecM {  // bytes of stream 0 for the p-th emitted read
    const uint64_t *ms;
    __device__ __forceinline__ uint64_t operator()(int64_t p) const { return (uint64_t)meta_namelen(ms[p]) + 1; }
};
struct ReadRecM {  // bytes of stream 1: packed rotated read + end marker (reads.cpp:128-130)
    const uint64_t *ms; int L1, sz_meta;
    __device__ __forceinline__ uint64_t operator()(int64_t p) const { return (uint64_t)(sz_read(L1 - meta_lvl(ms[p])) + sz_meta); }
};

struct EmitMParams {
    const uint8_t *names; const uint32_t *packed; int PW;
    const uint32_t *perm; const uint64_t *ms; const uint64_t *offN, *offR;
    int64_t n; int L1, sz_meta;
    uint8_t *oN, *oR;
};

// ---- dense writers: records of a CTA are assembled in shared memory, then leave as aligned 16-byte stores ----
// sbuf[(g0 & 15) + i] holds stream byte g0 + i for i in [0, len); gbase is the (16-byte aligned) start of the stream
__device__ __forceinline__ void flush_staged(uint8_t *__restrict__ gbase, uint64_t g0, int len, const uint8_t *sbuf) {
    const int mis = (int)(g0 & 15);
    const int nch = (mis + len + 15) >> 4;
    uint8_t *ga = gbase + (g0 - mis);
    for (int c = threadIdx.x; c < nch; c += blockDim.x) {
        const int lo = c * 16 - mis;
        if (lo >= 0 && lo + 16 <= len) *(uint4 *)(ga + 16 * c) = *(const uint4 *)(sbuf + 16 * c);
        else
            for (int k = 0; k < 16; k++) { const int i = lo + k; if (i >= 0 && i < len) ga[16 * c + k] = sbuf[16 * c + k]; }
    }
}

// stream 0: [len:u8][name] per read (names.cpp:48-62), 256 reads per CTA
constexpr int kNamesCap = 24 * 1024;
__global__ void __launch_bounds__(256) emit_names_st_k(EmitMParams e) {
    __shared__ __align__(16) uint8_t sb[kNamesCap + 32];
    const int64_t p0 = (int64_t)blockIdx.x * 256, p1 = (p0 + 256 < e.n) ? p0 + 256 : e.n;
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
            d[0] = (uint8_t)nl;                                                   // names.cpp:58
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
}

// stream 1: rotated 2-bit reads + end marker; RPB reads per CTA. Each read's packed row, metadata word and
// stream offset are fetched ONCE into shared memory (8-byte loads; rows are 8-byte aligned or the row pitch is
// odd and 4-byte loads are used), then one work item per (read, 4 record bytes) rotates out of shared memory.
constexpr int kEmitRowPad = 3;   // words readable past a staged row (pk_bits32 looks up to 2 words ahead)
__device__ __forceinline__ uint32_t spk_bits32(const uint32_t *row, int b0) {
    const int k = b0 >> 4, sh = (b0 & 15) * 2;
    return __funnelshift_l(row[k + 1], row[k], sh);
}

// ---- stream 4: mate 2 is packed without rotation (output_read(read2, dest, 0, 0), compress.cpp:696) ----
// One thread per output word (16 bases): consecutive threads take consecutive 16-byte pieces of the same (randomly placed)
// row. The piece is fetched as five aligned 32-bit words and packed four bases at a time (pack4, scan_smem.cuh: the scan's
// SWAR packer, same code table as base_code); the byte-wise path only serves the last few bytes of the input array
// (an aligned fetch would cross its end). Before: sixteen byte loads + base_code per thread, 12 ms of the 16.6 ms emit stage
// of the paired 25M x 2 x 150 bp configuration.
template <bool SEG>
__global__ void __launch_bounds__(256) emit_reads2_k(const RowSrc seq2, const uint32_t *__restrict__ perm, int64_t n, int L2,
                                                     uint8_t *__restrict__ oR2) {
    const int nb2 = sz_read(L2), nw = (nb2 + 3) >> 2;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * nw) return;
    const int64_t p = t / nw;
    const int w = (int)(t - p * nw);
    const uint32_t src_row = perm[p];
    const uint8_t *s = seq2.row<SEG>(src_row, L2) + 16 * w;
    const bool in_own = SEG && (int64_t)src_row >= seq2.lo && (int64_t)src_row < seq2.hi;
    const uint8_t *arr_end = in_own ? seq2.own + seq2.hi * (int64_t)L2 : seq2.p + n * (int64_t)L2;   // last byte + 1 of the array the row lives in
    const int nv = L2 - 16 * w < 16 ? L2 - 16 * w : 16;            // bases of this word (>= 1)
    uint32_t v = 0;
    const uintptr_t a0 = (uintptr_t)s & ~(uintptr_t)3;
    if (a0 + 20 <= (uintptr_t)arr_end) {
        const uint32_t *a = (const uint32_t *)a0;
        const uint32_t sh = ((uint32_t)(uintptr_t)s & 3u) * 8u;
        const uint32_t x0 = __ldg(a), x1 = __ldg(a + 1), x2 = __ldg(a + 2), x3 = __ldg(a + 3), x4 = __ldg(a + 4);
        const uint32_t y0 = __funnelshift_r(x0, x1, sh), y1 = __funnelshift_r(x1, x2, sh), y2 = __funnelshift_r(x2, x3, sh), y3 = __funnelshift_r(x3, x4, sh);
        uint32_t bad = 0;
        v = (pack4(y0, bad) << 24) | (pack4(y1, bad) << 16) | (pack4(y2, bad) << 8) | pack4(y3, bad);
        if (bad) v = (pack4_masked(y0) << 24) | (pack4_masked(y1) << 16) | (pack4_masked(y2) << 8) | pack4_masked(y3);   // also: the bytes past the row's end
        if (nv < 16) v &= ~(0xffffffffu >> (2 * nv));
    } else {
        for (int j = 0; j < 16; j++) v = (v << 2) | (j < nv ? base_code(s[j]) : 0u);
    }
    uint8_t *d = oR2 + p * (int64_t)nb2 + 4 * w;
    const int nbw = nb2 - 4 * w;
    if (nbw >= 4 && (((uintptr_t)d) & 3) == 0) { *(uint32_t *)d = __byte_perm(v, 0, 0x0123); return; }
    if (nbw >= 4 && (((uintptr_t)d) & 1) == 0) { *(uint16_t *)d = (uint16_t)__byte_perm(v, 0, 0x4423); *(uint16_t *)(d + 2) = (uint16_t)__byte_perm(v, 0, 0x4401); return; }
    d[0] = (uint8_t)(v >> 24);
    if (nbw > 1) d[1] = (uint8_t)(v >> 16);
    if (nbw > 2) d[2] = (uint8_t)(v >> 8);
    if (nbw > 3) d[3] = (uint8_t)v;
}

// one meta record per segment: int32 id, int32 core, int64 tN, tR, tQ [, tR2, tQ2]   reads.cpp:160-176
__global__ void meta2_k(SegTab t, int64_t n_seg, const uint64_t *__restrict__ offN, const uint64_t *__restrict__ offR, const int32_t *__restrict__ rank_node_id,
                        const int32_t *__restrict__ rank_core, int nb, int L1, int L2, int use_names, int use_quals, int paired,
                        uint8_t *__restrict__ meta, int64_t *__restrict__ chunk_first /*[2][n_chunks] or null*/, int n_chunks) {
    int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= n_seg) return;
    const int64_t p0 = t.pos[m], p1 = t.pos[m + 1];
    const uint32_t r = t.rank[m];
    const int32_t id = r == (uint32_t)nb ? SCB_ROOT_ID_DEV : rank_node_id[r];
    const int32_t core = r == (uint32_t)nb ? SCB_ROOT_ID_DEV : rank_core[r];
    const int64_t cnt = p1 - p0;
    const int nlen = 3 + 2 * paired;
    uint8_t *d = meta + m * (int64_t)(8 + 8 * nlen);
    int64_t v[5];
    v[0] = use_names ? (int64_t)(offN[p1] - offN[p0]) : 0;
    v[1] = (int64_t)(offR[p1] - offR[p0]);
    v[2] = use_quals ? cnt * L1 : 0;
    v[3] = cnt * sz_read(L2);
    v[4] = use_quals ? cnt * L2 : 0;
    memcpy(d, &id, 4); memcpy(d + 4, &core, 4);
    for (int k = 0; k < nlen; k++) memcpy(d + 8 + 8 * k, &v[k], 8);
    if (chunk_first) {
        const uint32_t c = t.chunk[m];
        if (m == 0 || t.chunk[m - 1] != c) { chunk_first[c] = p0; chunk_first[n_chunks + c] = m; }
    }
}

}  // namespace scb
