// pipeline.cuh - kernels of the boosting transform (v1: straightforward, every stage on the GPU).
//   scan     : DFA walk per read -> max core level + ordered distinct candidates   (aho_search minus the counts)
//   resolve  : the stateful tie-break on running bucket populations               (reads.cpp:420-421,246)
//   sizes    : rd.sz + sizeof(bin_node) accounting and flush-chunk ids            (compress.cpp:675-715)
//   keys     : (segment, suffix-key prefix) sort keys                              (reads.cpp:547-634 order)
//   ties     : refinement of equal-prefix runs with further key bases
//   emit     : rotate + 2-bit pack, gather of names / qualities / mate 2, meta     (reads.cpp:432-461, 91-180)
#pragma once
#include "common.cuh"
#include "prims.cuh"

namespace scb {

struct DfaDev {
    const uint32_t *next;      // [n_states*4]
    const int32_t *nto_rank;   // [n_states]
    const uint8_t *rank_level; // [n_buckets]
    int32_t n_states, n_buckets;
};

// -------------------------------------------------------------------------------------------------
// scan (v1: one thread per read, table in global memory / L1)
// -------------------------------------------------------------------------------------------------
constexpr int kScanListCap = 32;

__device__ __forceinline__ bool seen_before(const uint8_t *s, int p, int32_t r, const DfaDev &dfa) {
    uint32_t st = 0;
    for (int q = 0; q < p; q++) {
        st = dfa.next[st * 4 + base_code(s[q])];
        if (dfa.nto_rank[st] == r) return true;
    }
    return false;
}

template <bool EMIT>
__global__ void __launch_bounds__(128) scan_k(const uint8_t *__restrict__ seq, int64_t n, int L, DfaDev dfa,
                                              uint8_t *__restrict__ lvl, uint16_t *__restrict__ ncand,
                                              const uint64_t *__restrict__ cand_off, uint32_t *__restrict__ cand_rank,
                                              uint16_t *__restrict__ cand_pos) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint8_t *s = seq + i * (int64_t)L;
    uint32_t list[kScanListCap];
    uint32_t st = 0;
    int best = 0, cnt = 0;
    int final_lvl = EMIT ? (int)lvl[i] : 0;
    uint64_t off = EMIT ? cand_off[i] : 0;
    for (int p = 0; p < L; p++) {
        st = dfa.next[st * 4 + base_code(s[p])];
        int32_t r = dfa.nto_rank[st];
        if (r < 0) continue;
        int lv = dfa.rank_level[r];
        if (lv > best) { best = lv; cnt = 0; }
        if (lv == best) {
            bool dup = false;
            int lim = cnt < kScanListCap ? cnt : kScanListCap;
            for (int k = 0; k < lim; k++) dup |= (list[k] == (uint32_t)r);
            if (!dup && cnt >= kScanListCap) dup = seen_before(s, p, r, dfa);
            if (!dup) {
                if (cnt < kScanListCap) list[cnt] = (uint32_t)r;
                if (EMIT && lv == final_lvl) { cand_rank[off + cnt] = (uint32_t)r; cand_pos[off + cnt] = (uint16_t)p; }
                cnt++;
            }
        }
    }
    if (!EMIT) { lvl[i] = (uint8_t)best; ncand[i] = (uint16_t)cnt; }
}

// -------------------------------------------------------------------------------------------------
// resolve (v1: one warp walks the input in order, 32 reads per step; a read commits as soon as no
// earlier uncommitted read of the step shares a candidate bucket with it)
// life[r] = lifetime population of bucket rank r (aho_trie::bin_size), life[nb] = root.
// -------------------------------------------------------------------------------------------------
template <bool SMEM>
__global__ void __launch_bounds__(32) resolve_seq_k(int64_t n, const uint16_t *__restrict__ ncand,
                                                    const uint64_t *__restrict__ cand_off,
                                                    const uint32_t *__restrict__ cand_rank,
                                                    const uint16_t *__restrict__ cand_pos, unsigned long long *life_g,
                                                    uint32_t *claim_g, uint32_t *__restrict__ asg,
                                                    uint16_t *__restrict__ endv, int nb) {
    extern __shared__ unsigned long long sm_raw[];
    unsigned long long *life = SMEM ? sm_raw : life_g;
    uint32_t *claim = SMEM ? (uint32_t *)(sm_raw + nb + 1) : claim_g;
    const uint32_t l = threadIdx.x;
    if (SMEM) {
        for (int b = l; b <= nb; b += 32) { life[b] = life_g[b]; claim[b] = 0xffffffffu; }
        __syncwarp();
    }
    unsigned long long root_add = 0;
    for (int64_t g = 0; g < n; g += 32) {
        int64_t i = g + l;
        int nc = 0;
        uint64_t off = 0;
        if (i < n) { nc = ncand[i]; off = cand_off[i]; }
        bool pending = nc > 0;
        if (i < n && nc == 0) { asg[i] = (uint32_t)nb; endv[i] = 0; root_add++; }
        while (__any_sync(0xffffffffu, pending)) {
            if (pending)
                for (int k = 0; k < nc; k++) atomicMin(&claim[cand_rank[off + k]], l);
            __syncwarp();
            bool safe = pending;
            if (pending)
                for (int k = 0; k < nc; k++) safe &= (claim[cand_rank[off + k]] == l);
            __syncwarp();
            if (safe) {
                // first arg-max over the ordered candidates: replace only on strictly larger (reads.cpp:420-421)
                uint32_t br = cand_rank[off];
                unsigned long long bc = life[br];
                int bk = 0;
                for (int k = 1; k < nc; k++) {
                    uint32_t r = cand_rank[off + k];
                    unsigned long long c = life[r];
                    if (c > bc) { bc = c; br = r; bk = k; }
                }
                life[br] = bc + 1;                           // reads.cpp:246
                asg[i] = br;
                endv[i] = (uint16_t)(cand_pos[off + bk] + 1); // rd.end = n + 1, compress.cpp:682
            }
            if (pending)
                for (int k = 0; k < nc; k++) claim[cand_rank[off + k]] = 0xffffffffu;
            pending = pending && !safe;
            __syncwarp();
        }
    }
    __syncwarp();
    if (SMEM) {
        for (int b = l; b < nb; b += 32) life_g[b] = life[b];
    }
    atomicAdd(&life_g[nb], root_add);
}

// -------------------------------------------------------------------------------------------------
// size accounting + flush chunks
// -------------------------------------------------------------------------------------------------
struct RdSize {  // rd.sz + sizeof(bin_node), compress.cpp:675-702
    const int64_t *name_off; const uint8_t *lvl;
    int L1, fixed;           // fixed = quals + mate-2 bytes + 40
    int use_names;
    __device__ __forceinline__ uint64_t operator()(int64_t i) const {
        int nm = use_names ? (int)(name_off[i + 1] - name_off[i]) + 1 : 1;
        return (uint64_t)(nm + sz_read(L1 - (int)lvl[i]) + fixed);
    }
};

// S[0..n] exclusive prefix of RdSize (S[n] = total). A chunk ends with the first read that brings
// the running sum to >= B (compress.cpp:708-713). One thread; a binary search per chunk.
__global__ void chunk_bounds_k(const uint64_t *__restrict__ S, int64_t n, uint64_t B, uint32_t *chunk_start, int cap,
                               int *n_chunks) {
    int c = 0;
    int64_t start = 0;
    while (start < n) {
        if (c < cap) chunk_start[c] = (uint32_t)start;
        c++;
        uint64_t base = S[start];
        if (S[n] - base < B) break;
        int64_t lo = start, hi = n - 1;  // smallest i with S[i+1]-base >= B
        while (lo < hi) {
            int64_t mid = (lo + hi) >> 1;
            if (S[mid + 1] - base >= B) hi = mid; else lo = mid + 1;
        }
        start = lo + 1;
    }
    *n_chunks = c;
}

__global__ void chunk_ids_k(const uint32_t *__restrict__ chunk_start, int n_chunks, int64_t n, uint32_t *__restrict__ chunk) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int lo = 0, hi = n_chunks - 1;  // last c with chunk_start[c] <= i
    while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        if ((int64_t)chunk_start[mid] <= i) lo = mid; else hi = mid - 1;
    }
    chunk[i] = (uint32_t)lo;
}

// -------------------------------------------------------------------------------------------------
// sort keys
// -------------------------------------------------------------------------------------------------
// `nbases` bases of the in-bucket sort key of read i starting at key offset `from`:
// key = s[end..L) right-padded with A (reads.cpp:547-559); MSB-first 2 bits per base.
__device__ __forceinline__ uint64_t key_bits(const uint8_t *__restrict__ s, int L, int end, int from, int nbases) {
    uint64_t v = 0;
    for (int j = 0; j < nbases; j++) {
        int q = end + from + j;
        v = (v << 2) | (q < L ? base_code(s[q]) : 0u);
    }
    return v;
}

__device__ __forceinline__ uint32_t bucket_ord(uint32_t rank, int nb, int root_pos) {
    return rank == (uint32_t)nb ? (uint32_t)root_pos : rank + (rank >= (uint32_t)root_pos ? 1u : 0u);
}

// key = [seg : seg_bits][first pb key bases : 2*pb][zero fill]
__global__ void build_keys_k(const uint8_t *__restrict__ seq, int64_t n, int L, const uint32_t *__restrict__ asg,
                             const uint16_t *__restrict__ endv, const uint32_t *__restrict__ chunk, int nb, int root_pos,
                             int seg_bits, int pb, uint64_t *__restrict__ keys, uint32_t *__restrict__ vals) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t seg = (uint64_t)(chunk ? chunk[i] : 0u) * (uint64_t)(nb + 1) + bucket_ord(asg[i], nb, root_pos);
    uint64_t kb = key_bits(seq + i * (int64_t)L, L, endv[i], 0, pb);
    keys[i] = (seg_bits ? (seg << (64 - seg_bits)) : 0ull) | (pb ? (kb << (64 - seg_bits - 2 * pb)) : 0ull);
    vals[i] = (uint32_t)i;
}

// -------------------------------------------------------------------------------------------------
// tie refinement: elements whose key equals a neighbour's are re-keyed with further key bases
// -------------------------------------------------------------------------------------------------
__global__ void tie_flags_k(const uint64_t *__restrict__ k, int64_t n, uint8_t *__restrict__ flag) {
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    uint64_t v = k[p];
    bool t = (p > 0 && k[p - 1] == v) || (p + 1 < n && k[p + 1] == v);
    flag[p] = t ? 1 : 0;
}
struct HeadFlag {  // 1 at the first element of each run of tied elements (in compacted order)
    const uint64_t *k; const uint8_t *flag;
    __device__ __forceinline__ uint32_t operator()(int64_t p) const {
        return (flag[p] && !(p > 0 && k[p - 1] == k[p])) ? 1u : 0u;
    }
};
// compact tied elements: c_pos = global output position, c_idx = read index, c_grp = run number (0-based)
__global__ void tie_compact_k(const uint64_t *__restrict__ k, const uint8_t *__restrict__ flag,
                              const uint32_t *__restrict__ cpos, const uint32_t *__restrict__ hsum,
                              const uint32_t *__restrict__ pos_in, const uint32_t *__restrict__ idx_in, int64_t n,
                              uint32_t *__restrict__ c_pos, uint32_t *__restrict__ c_idx, uint32_t *__restrict__ c_grp) {
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n || !flag[p]) return;
    uint32_t c = cpos[p];
    bool head = !(p > 0 && k[p - 1] == k[p]);
    c_pos[c] = pos_in ? pos_in[p] : (uint32_t)p;
    c_idx[c] = idx_in[p];
    c_grp[c] = hsum[p] + (head ? 1u : 0u) - 1u;  // hsum exclusive
}
// sparse variant of the first round: positions of tied elements appended in no particular order
// (warp-aggregated), keys[] = position for the short sort that orders them
__global__ void tie_find_k(const uint64_t *__restrict__ k, int64_t n, uint64_t *__restrict__ out_key, uint32_t *__restrict__ out_val,
                           uint32_t cap, uint32_t *__restrict__ counter) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool t = false;
    if (p < n) {
        const uint64_t v = k[p];
        t = (p > 0 && k[p - 1] == v) || (p + 1 < n && k[p + 1] == v);
    }
    const uint32_t m = __ballot_sync(0xffffffffu, t);
    if (!m) return;
    uint32_t base = 0;
    if (lane_id() == (uint32_t)(__ffs(m) - 1)) base = atomicAdd(counter, (uint32_t)__popc(m));
    base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
    if (t) {
        const uint32_t o = base + __popc(m & lanemask_lt());
        if (o < cap) { out_key[o] = (uint64_t)p; out_val[o] = (uint32_t)p; }
    }
}
struct HeadAtPos {  // 1 where the c-th tied position (ascending) starts a run of equal keys
    const uint64_t *k; const uint32_t *pos;
    __device__ __forceinline__ uint32_t operator()(int64_t c) const {
        const uint32_t p = pos[c];
        return (p == 0 || k[p - 1] != k[p]) ? 1u : 0u;
    }
};
__global__ void tie_gather_sparse_k(HeadAtPos hp, const uint32_t *__restrict__ hsum, const uint32_t *__restrict__ idx_in, int64_t t,
                                    uint32_t *__restrict__ c_pos, uint32_t *__restrict__ c_idx, uint32_t *__restrict__ c_grp) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= t) return;
    const uint32_t p = hp.pos[c];
    c_pos[c] = p;
    c_idx[c] = idx_in[p];
    c_grp[c] = hsum[c] + hp(c) - 1u;   // hsum exclusive
}
__global__ void tie_rekey_k(const uint8_t *__restrict__ seq, int L, const uint16_t *__restrict__ endv,
                            const uint32_t *__restrict__ c_idx, const uint32_t *__restrict__ c_grp, int64_t m, int grp_bits,
                            int from, int nbases, uint64_t *__restrict__ keys, uint32_t *__restrict__ vals) {
    int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= m) return;
    uint32_t i = c_idx[c];
    uint64_t kb = key_bits(seq + (int64_t)i * L, L, endv[i], from, nbases);
    keys[c] = (grp_bits ? ((uint64_t)c_grp[c] << (64 - grp_bits)) : 0ull) | (kb << (64 - grp_bits - 2 * nbases));
    vals[c] = i;
}
__global__ void tie_writeback_k(const uint32_t *__restrict__ c_pos, const uint32_t *__restrict__ sorted_idx, int64_t m,
                                uint32_t *__restrict__ perm) {
    int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c < m) perm[c_pos[c]] = sorted_idx[c];
}

// -------------------------------------------------------------------------------------------------
// emit
// -------------------------------------------------------------------------------------------------
struct EmitParams {
    const uint8_t *seq1, *qual1, *names, *seq2, *qual2;
    const int64_t *name_off;
    const uint32_t *asg; const uint16_t *endv; const uint8_t *lvl; const uint32_t *chunk;
    const uint32_t *perm;
    int64_t n;
    int L1, L2, use_names, use_quals, paired, sz_meta, nb, root_pos;
};
struct NameRec {  // bytes of stream 0 for the p-th emitted read
    EmitParams e;
    __device__ __forceinline__ uint64_t operator()(int64_t p) const {
        if (!e.use_names) return 0;
        uint32_t i = e.perm[p];
        return (uint64_t)(e.name_off[i + 1] - e.name_off[i]) + 1;
    }
};
struct ReadRec {  // bytes of stream 1: packed rotated read + end marker (reads.cpp:128-130)
    EmitParams e;
    __device__ __forceinline__ uint64_t operator()(int64_t p) const {
        uint32_t i = e.perm[p];
        return (uint64_t)(sz_read(e.L1 - (int)e.lvl[i]) + e.sz_meta);
    }
};

// one warp per emitted read; byte-granular copies (v1)
__global__ void __launch_bounds__(256) emit_k(EmitParams e, const uint64_t *__restrict__ offN, const uint64_t *__restrict__ offR,
                                              uint8_t *__restrict__ oN, uint8_t *__restrict__ oR, uint8_t *__restrict__ oQ,
                                              uint8_t *__restrict__ oR2, uint8_t *__restrict__ oQ2) {
    int64_t p = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (p >= e.n) return;
    const int l = lane_id();
    const uint32_t i = e.perm[p];
    if (e.use_names) {
        int64_t a = e.name_off[i];
        int nl = (int)(e.name_off[i + 1] - a);
        uint8_t *d = oN + offN[p];
        if (l == 0) d[0] = (uint8_t)nl;                    // names.cpp:58
        for (int k = l; k < nl; k += 32) d[1 + k] = e.names[a + k];
    }
    {   // output_read(read, dest, end-level, level): bases [end, L) then [0, end-level)   reads.cpp:432-461
        const uint8_t *s = e.seq1 + (int64_t)i * e.L1;
        int lv = e.lvl[i], end = e.endv[i];
        int tail = e.L1 - end;         // bases after the core
        int total = e.L1 - lv;
        int nbytes = sz_read(total);
        uint8_t *d = oR + offR[p];
        for (int b = l; b < nbytes; b += 32) {
            uint32_t v = 0;
#pragma unroll
            for (int t = 0; t < 4; t++) {
                int j = 4 * b + t;
                uint32_t c = 0;
                if (j < total) c = base_code(s[j < tail ? end + j : j - tail]);
                v = (v << 2) | c;
            }
            d[b] = (uint8_t)v;
        }
        if (l < e.sz_meta) d[nbytes + l] = (uint8_t)((uint32_t)end >> (8 * l));   // low bytes of int16 end
    }
    if (e.use_quals) {
        const uint8_t *q = e.qual1 + (int64_t)i * e.L1;
        uint8_t *d = oQ + p * (int64_t)e.L1;
        for (int k = l; k < e.L1; k += 32) d[k] = q[k];
    }
    if (e.paired) {
        const uint8_t *s = e.seq2 + (int64_t)i * e.L2;
        int nbytes = sz_read(e.L2);
        uint8_t *d = oR2 + p * (int64_t)nbytes;
        for (int b = l; b < nbytes; b += 32) {
            uint32_t v = 0;
#pragma unroll
            for (int t = 0; t < 4; t++) {
                int j = 4 * b + t;
                v = (v << 2) | (j < e.L2 ? base_code(s[j]) : 0u);
            }
            d[b] = (uint8_t)v;
        }
        if (e.use_quals) {
            const uint8_t *q = e.qual2 + (int64_t)i * e.L2;
            uint8_t *dq = oQ2 + p * (int64_t)e.L2;
            for (int k = l; k < e.L2; k += 32) dq[k] = q[k];
        }
    }
}

// segment of the p-th emitted read: chunk-major (per-flush files) or bucket-major (merged)
struct SegOf {
    EmitParams e; int n_chunks; int merged;
    __device__ __forceinline__ uint64_t operator()(int64_t p) const {
        uint32_t i = e.perm[p];
        uint64_t o = bucket_ord(e.asg[i], e.nb, e.root_pos), c = e.chunk ? e.chunk[i] : 0u;
        return merged ? o : c * (uint64_t)(e.nb + 1) + o;
    }
};
struct SegHead {
    SegOf s;
    __device__ __forceinline__ uint32_t operator()(int64_t p) const { return (p == 0 || s(p) != s(p - 1)) ? 1u : 0u; }
};
__global__ void seg_heads_k(SegOf s, const uint32_t *__restrict__ hsum, int64_t n, uint32_t *__restrict__ hpos) {
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    if (p == 0 || s(p) != s(p - 1)) hpos[hsum[p]] = (uint32_t)p;
}
// one meta record per segment: int32 id, int32 core, int64 tN, tR, tQ [, tR2, tQ2]   reads.cpp:160-176
__global__ void meta_k(SegOf s, const uint32_t *__restrict__ hpos, int64_t n_seg, const uint64_t *__restrict__ offN,
                       const uint64_t *__restrict__ offR, const int32_t *__restrict__ rank_node_id,
                       const int32_t *__restrict__ rank_core, uint8_t *__restrict__ meta, int64_t *__restrict__ chunk_first /*[2][n_chunks]: pos, meta idx*/) {
    int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= n_seg) return;
    const EmitParams &e = s.e;
    int64_t p0 = hpos[m], p1 = (m + 1 < n_seg) ? (int64_t)hpos[m + 1] : e.n;
    uint32_t i = e.perm[p0];
    uint32_t r = e.asg[i];
    int32_t id = r == (uint32_t)e.nb ? SCB_ROOT_ID_DEV : rank_node_id[r];
    int32_t core = r == (uint32_t)e.nb ? SCB_ROOT_ID_DEV : rank_core[r];
    int64_t cnt = p1 - p0;
    int nlen = 3 + 2 * e.paired;
    uint8_t *d = meta + m * (int64_t)(8 + 8 * nlen);
    int64_t v[5];
    v[0] = e.use_names ? (int64_t)(offN[p1] - offN[p0]) : 0;
    v[1] = (int64_t)(offR[p1] - offR[p0]);
    v[2] = e.use_quals ? cnt * e.L1 : 0;
    v[3] = cnt * sz_read(e.L2);
    v[4] = e.use_quals ? cnt * e.L2 : 0;
    memcpy(d, &id, 4); memcpy(d + 4, &core, 4);
    for (int k = 0; k < nlen; k++) memcpy(d + 8 + 8 * k, &v[k], 8);
    if (chunk_first) {
        uint32_t c = e.chunk ? e.chunk[i] : 0u;
        bool first = (m == 0);
        if (!first) { uint32_t ip = e.perm[hpos[m - 1]]; first = (e.chunk ? e.chunk[ip] : 0u) != c; }
        if (first) { chunk_first[c] = p0; chunk_first[s.n_chunks + c] = m; }
    }
}

// per-read arrays the reference exposes through meta / debugging
__global__ void debug_arrays_k(int64_t n, const uint32_t *__restrict__ asg, int nb, const int32_t *__restrict__ rank_node_id,
                               const int32_t *__restrict__ rank_core, int32_t *__restrict__ bucket_id, int32_t *__restrict__ core_idx) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t r = asg[i];
    bucket_id[i] = r == (uint32_t)nb ? SCB_ROOT_ID_DEV : rank_node_id[r];
    core_idx[i] = r == (uint32_t)nb ? -1 : rank_core[r];
}
// keys for the merged order: bucket order of the p-th read of the chunk-major order
__global__ void merged_keys_k(const uint32_t *__restrict__ asg, const uint32_t *__restrict__ perm, int64_t n, int nb, int root_pos,
                              uint64_t *__restrict__ keys, uint32_t *__restrict__ vals) {
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    uint32_t i = perm[p];
    keys[p] = bucket_ord(asg[i], nb, root_pos);
    vals[p] = i;
}
__global__ void widen_u16_k(const uint16_t *__restrict__ a, int64_t n, int32_t *__restrict__ o) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) o[i] = a[i];
}
__global__ void fill_u32_k(uint32_t *a, int64_t n, uint32_t v) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] = v;
}

}  // namespace scb
