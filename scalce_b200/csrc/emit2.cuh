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
    return __funnelshift_l(row[k + 1], row[k], sh);
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

// ---- segments ----------------------------------------------------------------------------------------
struct KeyHead {   // 1 where the segment id (top `bits` bits of the sorted key, 0 bits = one segment) changes
    const uint64_t *keys; int shift, bits;
    __device__ __forceinline__ uint32_t operator()(int64_t p) const {
        if (p == 0) return 1u;
        if (bits == 0) return 0u;
        return ((keys[p] >> shift) != (keys[p - 1] >> shift)) ? 1u : 0u;
    }
};
struct NameRec2 {  // bytes of stream 0 for the p-th emitted read
    const uint32_t *perm; const int64_t *name_off;
    __device__ __forceinline__ uint64_t operator()(int64_t p) const {
        uint32_t i = perm[p];
        return (uint64_t)(name_off[i + 1] - name_off[i]) + 1;
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
struct SegBytes {   // stream-1 bytes of a segment
    const uint32_t *pos, *recsz;
    __device__ __forceinline__ uint64_t operator()(int64_t m) const { return (uint64_t)(pos[m + 1] - pos[m]) * recsz[m]; }
};

// ---- emit ---------------------------------------------------------------------------------------------
struct Emit2Params {
    const uint8_t *qual1, *names, *seq2, *qual2;
    const int64_t *name_off;
    const uint32_t *packed; int PW;
    const uint16_t *endv;
    const uint32_t *perm;
    const uint32_t *hsum;              // [n] segment index of position p = hsum[p] + head(p) - 1 -> stored as sidx
    const uint32_t *seg_pos, *seg_recsz, *seg_rank;
    const uint64_t *seg_off;           // [n_seg] stream-1 byte offset of the segment
    const uint8_t *rank_level;
    const uint64_t *offN;              // [n+1]
    int64_t n;
    int L1, L2, use_names, use_quals, paired, sz_meta, nb;
    uint8_t *oN, *oR, *oQ, *oR2, *oQ2;
};

// copy `len` bytes src -> dst (arbitrary alignments) with 4-byte stores; lanes h, h+16, ... of a half warp
__device__ __forceinline__ void copy_bytes16(uint8_t *__restrict__ dst, const uint8_t *__restrict__ src, int len, int h) {
    const uintptr_t d0 = (uintptr_t)dst;
    const int head = (int)((4 - (d0 & 3)) & 3);              // bytes before the first aligned dst word
    if (len <= 8 + head) { for (int k = h; k < len; k += 16) dst[k] = src[k]; return; }
    if (h < head) dst[h] = src[h];
    const int nw = (len - head) >> 2;                         // full aligned dst words
    const uint8_t *s0 = src + head;
    const int sh = (int)(((uintptr_t)s0 & 3) * 8);
    const uint32_t *sa = (const uint32_t *)((uintptr_t)s0 & ~(uintptr_t)3);
    uint32_t *da = (uint32_t *)(dst + head);
    for (int k = h; k < nw; k += 16) {
        uint32_t lo = sa[k];
        uint32_t v = lo;
        if (sh) { uint32_t hi = sa[k + 1]; v = __funnelshift_r(lo, hi, sh); }
        da[k] = v;
    }
    const int done = head + (nw << 2);
    if (h < len - done) dst[done + h] = src[done + h];
}

// ---- stream 2 / 5: qualities are a pure row gather: out row p <- in row perm[p] -----------------------
// One thread per 16 output bytes (the output is dense, so stores are aligned STG.128 and fully
// coalesced); a chunk may straddle two rows. Sources are unaligned: two aligned 16-byte loads and a
// byte shift. Requires L >= 16.
__device__ __forceinline__ uint4 load16_unaligned(const uint8_t *__restrict__ a, int nbytes) {
    const uintptr_t A = (uintptr_t)a;
    const uint4 *a0 = (const uint4 *)(A & ~(uintptr_t)15);
    const int sh = (int)(A & 15);
    uint4 v0 = a0[0], v1 = make_uint4(0, 0, 0, 0);
    if (sh + nbytes > 16) v1 = a0[1];               // only when it holds a needed byte (never past the buffer's last granule)
    const int bs = (sh & 3) * 8;
    uint32_t w0, w1, w2, w3, w4;
    switch (sh >> 2) {
        case 0: w0 = v0.x; w1 = v0.y; w2 = v0.z; w3 = v0.w; w4 = v1.x; break;
        case 1: w0 = v0.y; w1 = v0.z; w2 = v0.w; w3 = v1.x; w4 = v1.y; break;
        case 2: w0 = v0.z; w1 = v0.w; w2 = v1.x; w3 = v1.y; w4 = v1.z; break;
        default: w0 = v0.w; w1 = v1.x; w2 = v1.y; w3 = v1.z; w4 = v1.w; break;
    }
    return make_uint4(__funnelshift_r(w0, w1, bs), __funnelshift_r(w1, w2, bs), __funnelshift_r(w2, w3, bs), __funnelshift_r(w3, w4, bs));
}
// dst bytes [0, k) from x, bytes [k, 16) from y shifted up by k (k in 1..15)
__device__ __forceinline__ uint4 splice16(uint4 x, uint4 y, int k) {
    uint32_t xs[4] = {x.x, x.y, x.z, x.w}, ys[4] = {y.x, y.y, y.z, y.w}, o[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const int lo = 4 * j;                                   // first byte of output word j
        uint32_t v;
        if (lo + 4 <= k) v = xs[j];
        else if (lo >= k) {                                     // bytes lo-k .. lo-k+3 of y
            const int s = lo - k, ws = s >> 2, bs = (s & 3) * 8;
            uint32_t a = 0, b2 = 0;
#pragma unroll
            for (int q = 0; q < 4; q++) { a = (ws == q) ? ys[q] : a; b2 = (ws + 1 == q) ? ys[q] : b2; }
            v = __funnelshift_r(a, b2, bs);
        } else {                                                // mixed word: k - lo bytes of x, rest from y[0..]
            const int nx = k - lo;
            v = (xs[j] & (0xffffffffu >> (8 * (4 - nx)))) | (ys[0] << (8 * nx));
        }
        o[j] = v;
    }
    return make_uint4(o[0], o[1], o[2], o[3]);
}
__global__ void __launch_bounds__(256) gather_rows16_k(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst,
                                                       const uint32_t *__restrict__ perm, int64_t n, int L) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t o = c << 4, total = n * (int64_t)L;
    if (o >= total) return;
    const int64_t p0 = o / L;
    const int r0 = (int)(o - p0 * L);
    const int n0 = min(16, L - r0);
    uint4 v = load16_unaligned(src + (int64_t)perm[p0] * L + r0, n0);
    if (n0 < 16 && p0 + 1 < n) {
        const uint4 y = load16_unaligned(src + (int64_t)perm[p0 + 1] * L, 16 - n0);
        v = splice16(v, y, n0);
    }
    if (o + 16 <= total) *(uint4 *)(dst + o) = v;               // dst is 16-byte aligned (allocation base)
    else {
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
        for (int k = 0; k < (int)(total - o); k++) dst[o + k] = (uint8_t)(w[k >> 2] >> (8 * (k & 3)));
    }
}
__global__ void gather_rows_small_k(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst, const uint32_t *__restrict__ perm,
                                    int64_t n, int L) {
    const int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= n * (int64_t)L) return;
    const int64_t p = o / L;
    dst[o] = src[(int64_t)perm[p] * L + (o - p * L)];
}

// ---- stream 0: names, one thread per read --------------------------------------------------------------
__global__ void emit_names_k(Emit2Params e) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= e.n) return;
    const uint32_t i = e.perm[p];
    const int64_t a = e.name_off[i];
    const int nl = (int)(e.offN[p + 1] - e.offN[p]) - 1;
    uint8_t *d = e.oN + e.offN[p];
    d[0] = (uint8_t)nl;                                                           // names.cpp:58
    for (int k = 0; k < nl; k++) d[1 + k] = e.names[a + k];
}

// ---- stream 1: rotated 2-bit reads + end marker, 16 lanes per read -------------------------------------
__global__ void __launch_bounds__(256) emit_reads_k(Emit2Params e) {
    const int64_t p = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 4;
    if (p >= e.n) return;
    const int h = threadIdx.x & 15;
    const uint32_t i = e.perm[p];
    // output_read(read, dest, end-level, level): bases [end, L) then [0, end-level); reads.cpp:432-461
    uint32_t m = e.hsum[p];                              // exclusive head count: a head's own index, else one past
    if (e.seg_pos[m] != (uint32_t)p) m -= 1;             // seg_pos has n_seg+1 entries (last = n)
    const uint32_t r = e.seg_rank[m];
    const int lv = r == (uint32_t)e.nb ? 0 : e.rank_level[r];
    const int end = e.endv[i];
    const int tail = e.L1 - end, total = e.L1 - lv;
    const int nbytes = sz_read(total);
    uint8_t *d = e.oR + e.seg_off[m] + (uint64_t)((uint32_t)p - e.seg_pos[m]) * e.seg_recsz[m];
    const uint32_t *row = e.packed + (int64_t)i * e.PW;   // rows have 2 words of slack behind the last one
    // one lane builds 16 rotated bases (4 output bytes): `a` of them come from behind the core, the
    // rest from the front of the read; whatever lies past `total` is zero fill
    for (int w = h; 4 * w < nbytes; w += 16) {
        const int j0 = 16 * w;
        int a = tail - j0; a = a < 0 ? 0 : (a > 16 ? 16 : a);
        int nv = total - j0; nv = nv > 16 ? 16 : nv;
        uint32_t v = 0;
        if (a > 0) v = pk_bits32(row, end + j0) & (a == 16 ? 0xffffffffu : ~(0xffffffffu >> (2 * a)));
        if (a < 16) v |= pk_bits32(row, j0 + a - tail) >> (2 * a);
        if (nv < 16) v &= ~(0xffffffffu >> (2 * nv));
        const int nbw = nbytes - 4 * w;                   // bytes of this word that belong to the record
        uint8_t *dw = d + 4 * w;
        dw[0] = (uint8_t)(v >> 24);
        if (nbw > 1) dw[1] = (uint8_t)(v >> 16);
        if (nbw > 2) dw[2] = (uint8_t)(v >> 8);
        if (nbw > 3) dw[3] = (uint8_t)v;
    }
    if (h < e.sz_meta) d[nbytes + h] = (uint8_t)((uint32_t)end >> (8 * h));   // low bytes of int16 end, reads.cpp:130
}

// ---- stream 4: mate 2 is packed without rotation (output_read(read2, dest, 0, 0), compress.cpp:696) ----
__global__ void emit_reads2_k(const uint8_t *__restrict__ seq2, const uint32_t *__restrict__ perm, int64_t n, int L2,
                              uint8_t *__restrict__ oR2) {
    const int nb2 = sz_read(L2), nw = (nb2 + 3) >> 2;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * nw) return;
    const int64_t p = t / nw;
    const int w = (int)(t - p * nw);
    const uint8_t *s = seq2 + (int64_t)perm[p] * L2;
    uint32_t v = 0;
#pragma unroll
    for (int j = 0; j < 16; j++) {
        const int q = 16 * w + j;
        v = (v << 2) | (q < L2 ? base_code(s[q]) : 0u);
    }
    uint8_t *d = oR2 + p * (int64_t)nb2 + 4 * w;
    const int nbw = nb2 - 4 * w;
    d[0] = (uint8_t)(v >> 24);
    if (nbw > 1) d[1] = (uint8_t)(v >> 16);
    if (nbw > 2) d[2] = (uint8_t)(v >> 8);
    if (nbw > 3) d[3] = (uint8_t)v;
}

// one meta record per segment: int32 id, int32 core, int64 tN, tR, tQ [, tR2, tQ2]   reads.cpp:160-176
__global__ void meta2_k(SegTab t, int64_t n_seg, const uint64_t *__restrict__ offN, const int32_t *__restrict__ rank_node_id,
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
    v[1] = cnt * (int64_t)t.recsz[m];
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
