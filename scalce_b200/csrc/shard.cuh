// shard.cuh - kernels of the sharded (multi-GPU) run, SURVEY.md 8(e).
//
// The global input is the concatenation of the ranks' shards in rank order. Each rank scans its shard,
// the tie-break is resolved jointly (resolve_dense.cuh, one global round per launch), flush-chunk ids are
// numbered along the global order, and then reads are exchanged by BUCKET RANGE so that every rank owns
// a contiguous slice of the bucket emission order (all flush chunks of its buckets). What travels per read:
//   aux word (bucket rank, end marker, name length, flush chunk), the 2-bit packed row, the quality row,
//   the name bytes and, for pairs, mate 2's bases and qualities.
// Senders keep input order inside every destination and receivers concatenate sources in rank order, so
// the received reads are in global input order and the local stable sort + emit is the one-GPU code.
#pragma once
#include "common.cuh"
#include "emit2.cuh"
#include "pipeline.cuh"

namespace scb {

// ---- aux word: bits 0-23 bucket rank, 24-35 end marker, 36-43 name length, 44-63 flush chunk ------------
constexpr uint32_t kAuxMaxBuckets = 1u << 24;
constexpr uint32_t kAuxMaxChunks = 1u << 20;
__host__ __device__ __forceinline__ uint64_t aux_pack(uint32_t asg, uint32_t end, uint32_t namelen, uint32_t chunk) {
    return (uint64_t)(asg & 0xffffffu) | ((uint64_t)(end & 0xfffu) << 24) | ((uint64_t)(namelen & 0xffu) << 36) | ((uint64_t)(chunk & 0xfffffu) << 44);
}
__host__ __device__ __forceinline__ uint32_t aux_asg(uint64_t a) { return (uint32_t)(a & 0xffffffu); }
__host__ __device__ __forceinline__ uint32_t aux_end(uint64_t a) { return (uint32_t)((a >> 24) & 0xfffu); }
__host__ __device__ __forceinline__ uint32_t aux_namelen(uint64_t a) { return (uint32_t)((a >> 36) & 0xffu); }
__host__ __device__ __forceinline__ uint32_t aux_chunk(uint64_t a) { return (uint32_t)(a >> 44); }

// ---- flush chunks along the GLOBAL order ---------------------------------------------------------------
// S[0..n] = exclusive prefix of rd.sz + 40 over the local shard. The running sum enters the shard at
// `carry_in` (bytes already in the open chunk). bounds[k] = local index where the (k+1)-th new chunk of this
// shard starts (may equal n: the flush fell on the shard's last read). compress.cpp:702, 708-713.
__global__ void chunk_bounds_carry_k(const uint64_t *__restrict__ S, int64_t n, uint64_t B, uint64_t carry_in,
                                     uint32_t *bounds, int cap, unsigned long long *out /* [0] n_bounds, [1] carry_out */) {
    int c = 0;
    int64_t start = 0;
    uint64_t carry = carry_in;
    while (true) {
        const uint64_t base = S[start];
        if (carry + (S[n] - base) < B) { carry += S[n] - base; break; }
        int64_t lo = start, hi = n - 1;   // smallest i with carry + S[i+1] - base >= B
        while (lo < hi) {
            const int64_t mid = (lo + hi) >> 1;
            if (carry + (S[mid + 1] - base) >= B) hi = mid; else lo = mid + 1;
        }
        start = lo + 1;
        if (c < cap) bounds[c] = (uint32_t)start;
        c++;
        carry = 0;
        if (start >= n) break;
    }
    out[0] = (unsigned long long)c;
    out[1] = carry;
}
__global__ void chunk_ids_global_k(const uint32_t *__restrict__ bounds, int n_bounds, uint32_t chunk_in, int64_t n, uint32_t *__restrict__ chunk) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int lo = 0, hi = n_bounds;   // number of bounds <= i
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if ((int64_t)bounds[mid] <= i) lo = mid + 1; else hi = mid;
    }
    chunk[i] = chunk_in + (uint32_t)lo;
}

// ---- bucket histogram in emission order (the input of the bucket-range split) ---------------------------
template <bool SMEM>
__global__ void __launch_bounds__(512) bucket_hist_k(const uint32_t *__restrict__ asg, int64_t n, int nb, int root_pos, uint32_t *__restrict__ hist) {
    extern __shared__ uint32_t sh[];
    if (SMEM) {
        for (int k = threadIdx.x; k <= nb; k += blockDim.x) sh[k] = 0;
        __syncthreads();
    }
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t o = bucket_ord(asg[i], nb, root_pos);
        atomicAdd(SMEM ? &sh[o] : &hist[o], 1u);
    }
    if (SMEM) {
        __syncthreads();
        for (int k = threadIdx.x; k <= nb; k += blockDim.x)
            if (sh[k]) atomicAdd(&hist[k], sh[k]);
    }
}

// ---- destination of every read: last g with split[g] <= emission order -----------------------------------
constexpr int kMaxRanks = 64;
struct SplitTab { uint32_t s[kMaxRanks + 1]; int G; };
__global__ void dest_keys_k(const uint32_t *__restrict__ asg, int64_t n, int nb, int root_pos, SplitTab t,
                            uint64_t *__restrict__ keys, uint32_t *__restrict__ vals) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t o = bucket_ord(asg[i], nb, root_pos);
    int g = 0;
    for (int k = 1; k < t.G; k++) g += (t.s[k] <= o) ? 1 : 0;
    keys[i] = (uint64_t)g;
    vals[i] = (uint32_t)i;
}
// chunk ownership: destination = owner of the read's flush chunk. Owners never decrease along the input order, so the keys
// come out sorted and the send order is the input order (no sort pass).
__global__ void dest_keys_chunk_k(const uint32_t *__restrict__ chunk, int64_t n, const uint8_t *__restrict__ owner,
                                  uint64_t *__restrict__ keys, uint32_t *__restrict__ vals) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    keys[i] = (uint64_t)owner[chunk[i]];
    vals[i] = (uint32_t)i;
}
// first[g] = first position of the sorted destination keys holding a value >= g, g = 0..G
__global__ void dest_bounds_k(const uint64_t *__restrict__ keys, int64_t n, int G, int64_t *__restrict__ first) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g > G) return;
    int64_t lo = 0, hi = n;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (keys[mid] >= (uint64_t)g) hi = mid; else lo = mid + 1;
    }
    first[g] = lo;
}

// ---- send side -------------------------------------------------------------------------------------------
__global__ void pack_aux_k(const uint32_t *__restrict__ perm, int64_t n, const uint32_t *__restrict__ asg, const uint16_t *__restrict__ endv,
                           const int64_t *__restrict__ name_off, const uint32_t *__restrict__ chunk, uint64_t *__restrict__ aux) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const uint32_t i = perm[j];
    const uint32_t nl = name_off ? (uint32_t)(name_off[i + 1] - name_off[i]) : 0u;
    aux[j] = aux_pack(asg[i], endv[i], nl, chunk ? chunk[i] : 0u);
}
struct AuxNameLen {
    const uint64_t *aux;
    __device__ __forceinline__ uint64_t operator()(int64_t j) const { return (uint64_t)aux_namelen(aux[j]); }
};
// names in send order without the length byte; one thread per read
__global__ void pack_names_k(const uint32_t *__restrict__ perm, int64_t n, const int64_t *__restrict__ name_off,
                             const uint8_t *__restrict__ names, const uint64_t *__restrict__ off_out, uint8_t *__restrict__ out) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const uint32_t i = perm[j];
    const int64_t a = name_off[i];
    const int nl = (int)(name_off[i + 1] - a);
    uint8_t *d = out + off_out[j];
    for (int k = 0; k < nl; k++) d[k] = (uint8_t)ldg_g64(names + a + k);
}
// name-byte offsets at the destination boundaries
__global__ void gather_u64_k(const uint64_t *__restrict__ src, const int64_t *__restrict__ idx, int m, int64_t *__restrict__ out) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < m) out[k] = (int64_t)src[idx[k]];
}

// ---- fused pack + send: row gathers that write straight into the owner's receive buffer ---------------------
// dst is peer memory mapped over NVLink (or this GPU's own receive buffer) at an arbitrary byte offset, so the
// 16-byte chunks are aligned in DESTINATION address space: chunk c covers stream bytes [16c - mis, 16c - mis + 16)
// with mis = dst & 15; the (at most two) partial chunks at the ends use byte stores. Requires L >= 16.
__global__ void __launch_bounds__(256) gather_rows16_to_k(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst,
                                                          const uint32_t *__restrict__ perm, int64_t n, int L) {
    const int mis = (int)((uintptr_t)dst & 15);
    const int64_t total = n * (int64_t)L;
    const int64_t nchunks = (total + mis + 15) >> 4;
    for (int64_t cblk = (int64_t)blockIdx.x * (256 * kGatherChunks); cblk < nchunks; cblk += (int64_t)gridDim.x * (256 * kGatherChunks)) {
    const int64_t s_blk = cblk * 16 - mis;
    const int64_t sb = s_blk < 0 ? 0 : s_blk;          // first stream byte of the block, clamped for the division
    const int64_t pblk = sb / L;
    const uint32_t rblk = (uint32_t)(sb - pblk * L);
    const uint8_t *a0[kGatherChunks], *a1[kGatherChunks];
    int n0[kGatherChunks];
    int64_t s0[kGatherChunks];
#pragma unroll
    for (int q = 0; q < kGatherChunks; q++) {
        const int64_t c = cblk + q * 256 + threadIdx.x;
        a0[q] = a1[q] = nullptr; n0[q] = 16;
        s0[q] = c * 16 - mis;
        if (c < nchunks && s0[q] >= 0 && s0[q] + 16 <= total) {
            const uint32_t x = rblk + (uint32_t)(s0[q] - sb), dp = x / (uint32_t)L;
            const int64_t p0 = pblk + dp;
            const int r0 = (int)(x - dp * (uint32_t)L);
            n0[q] = min(16, L - r0);
            a0[q] = src + (int64_t)perm[p0] * L + r0;
            if (n0[q] < 16) a1[q] = src + (int64_t)perm[p0 + 1] * L;   // p0 + 1 < n: the chunk lies inside the stream
        }
    }
    uint4 v[kGatherChunks];
#pragma unroll
    for (int q = 0; q < kGatherChunks; q++)
        if (a0[q]) {
            v[q] = load16_unaligned(a0[q], n0[q]);
            if (a1[q]) v[q] = splice16(v[q], load16_unaligned(a1[q], 16 - n0[q]), n0[q]);
        }
#pragma unroll
    for (int q = 0; q < kGatherChunks; q++) {
        const int64_t c = cblk + q * 256 + threadIdx.x;
        if (c >= nchunks) continue;
        if (a0[q]) { *(uint4 *)(dst + s0[q]) = v[q]; continue; }
        for (int k = 0; k < 16; k++) {                 // partial chunk at either end of the stream
            const int64_t sx = s0[q] + k;
            if (sx < 0 || sx >= total) continue;
            const int64_t pr = sx / L;
            dst[sx] = src[(int64_t)perm[pr] * L + (sx - pr * L)];
        }
    }
    }   // grid-stride loop (the grid is capped when the send overlaps the receive side's sort)
}
// rows of W 32-bit words (packed reads of short inputs): one thread per word
__global__ void gather_words_to_k(const uint32_t *__restrict__ src, uint32_t *__restrict__ dst, const uint32_t *__restrict__ perm, int64_t n, int W) {
    const int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= n * (int64_t)W) return;
    const int64_t pr = o / W;
    dst[o] = src[(int64_t)perm[pr] * W + (o - pr * W)];
}

// ---- receive side ------------------------------------------------------------------------------------------
__global__ void unpack_aux_k(const uint64_t *__restrict__ aux, int64_t n, int nb, const uint8_t *__restrict__ rank_level,
                             uint32_t *__restrict__ asg, uint16_t *__restrict__ endv, uint8_t *__restrict__ lvl, uint32_t *__restrict__ chunk) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const uint64_t a = aux[j];
    const uint32_t r = aux_asg(a);
    asg[j] = r;
    endv[j] = (uint16_t)aux_end(a);
    lvl[j] = r == (uint32_t)nb ? (uint8_t)0 : rank_level[r];
    if (chunk) chunk[j] = aux_chunk(a);
}

// ---- resolve plumbing --------------------------------------------------------------------------------------
// base[col] = lifetime population + populations of the lower ranks' shards (u32 engine counters)
__global__ void add_life_k(const unsigned long long *__restrict__ life, const uint32_t *__restrict__ before, int nb1, uint32_t *__restrict__ base) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nb1) base[i] = (uint32_t)life[i] + (before ? before[i] : 0u);
}
// tot[col] = base_after[col] - life[col] (the shard's histogram after a local run), tot[nb1] = 0
__global__ void sub_life_k(const uint32_t *__restrict__ base, const unsigned long long *__restrict__ life, int nb1, uint32_t *__restrict__ tot) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nb1) tot[i] = base[i] - (uint32_t)life[i];
    if (i == nb1) tot[i] = 0u;
}
// ---- helpers of the C++ orchestrator (scb_shard_flush): column sums over the ranks' histogram rows ------------------------
// out[c] = sum over g in [g0, g1) of rows[g * RW + c]   (u32 wrap-around = the u32 sum)
__global__ void sh_sum_rows_k(const uint32_t *__restrict__ rows, int RW, int g0, int g1, int ncols, uint32_t *__restrict__ out) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncols) return;
    uint32_t v = 0;
    for (int g = g0; g < g1; g++) v += rows[(size_t)g * RW + c];
    out[c] = v;
}
// first joint round: rank 0's histogram scaled to the reads before this shard (exact for rank 1)
__global__ void sh_guess_k(const uint32_t *__restrict__ row0, unsigned long long before, unsigned long long n0, int ncols, uint32_t *__restrict__ out) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < ncols) out[c] = n0 ? (uint32_t)(((unsigned long long)row0[c] * before) / n0) : 0u;
}
// decisions that changed on the ranks >= 1 in this round
__global__ void sh_changed_k(const uint32_t *__restrict__ rows, int RW, int G, int ncols, unsigned long long *__restrict__ out) {
    unsigned long long v = 0;
    for (int g = 1; g < G; g++) v += rows[(size_t)g * RW + ncols];
    *out = v;
}

// life[col] += global histogram of this distributed flush (root, index nb, stays local: resolve_finalize_k)
__global__ void life_add_k(unsigned long long *__restrict__ life, const uint32_t *__restrict__ tot, int nb) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nb) life[i] += tot[i];
}

}  // namespace scb
