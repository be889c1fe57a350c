// pipeline.cuh - the general-case kernels of the boosting transform and small helpers shared by all stages.
//   scan_k        : DFA walk per read with the table in global memory (automata too large for shared memory)
//   resolve_seq_k : the stateful tie-break, one warp in input order (bucket sets too large for the dense engine)
//   sizes         : rd.sz + sizeof(bin_node) accounting and flush-chunk ids            (compress.cpp:675-715)
//   ties          : detection / compaction of equal-prefix runs for the iterative refinement rounds
//   debug arrays  : per-read bucket id / core index as the reference writes them to meta
// The shared-memory scan is scan_smem.cuh, the dense tie-break resolve_dense.cuh, the output side emit2.cuh.
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

// name lengths of a device-resident batch: out[0] = max length, out[1] = max of -length (i.e. -min); out pre-set to 0
__global__ void __launch_bounds__(256) name_len_range_k(const int64_t *__restrict__ name_off, int64_t n, long long *__restrict__ out) {
    // grid-stride: a few thousand warps in all, so that the one shared word each of them reads (and rarely updates) at the end is
    // not hit by 1.5 M warps (that same-address traffic was 0.8 ms of a 0.07 ms pass over 50 M offsets)
    long long mx = 0, mn = -(1ll << 62);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const long long len = (long long)(name_off[i + 1] - name_off[i]);
        mx = len > mx ? len : mx; mn = -len > mn ? -len : mn;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        const long long a = __shfl_xor_sync(0xffffffffu, mx, d), b = __shfl_xor_sync(0xffffffffu, mn, d);
        mx = a > mx ? a : mx; mn = b > mn ? b : mn;
    }
    // a racy read first: the maxima only grow, so a stale value can only cause a redundant atomic, never a missed one
    if (lane_id() == 0) { if (mx > ((volatile long long *)out)[0]) atomicMax(out, mx); if (mn > ((volatile long long *)out)[1]) atomicMax(out + 1, mn); }
}

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

// streaming flush: first read of the flush's OPEN chunk (n if the last chunk closed with the last read) - compress.cpp:708-713
__global__ void chunk_open_from_k(const uint64_t *__restrict__ S, int64_t n, uint64_t B, const uint32_t *__restrict__ chunk_start, int n_chunks, long long *out) {
    if (n_chunks <= 0 || n == 0) { *out = 0; return; }
    const int64_t s = chunk_start[n_chunks - 1];
    *out = (S[n] - S[s] >= B) ? (long long)n : (long long)s;
}
// takes the reads [from, n) out of the lifetime populations again (they stay pending); roots = how many of them had no core
__global__ void uncommit_tail_k(const uint32_t *__restrict__ asg, int64_t from, int64_t n, unsigned long long *__restrict__ life, int nb,
                                unsigned long long *__restrict__ roots) {
    const int64_t i = from + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t a = asg[i];
    atomicAdd(&life[a], ~0ull);              // - 1
    if (a == (uint32_t)nb) atomicAdd(roots, 1ull);
}
__global__ void rebase_off_k(const int64_t *__restrict__ off, int64_t m1, int64_t *__restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < m1) out[i] = off[i] - off[0];
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
__device__ __forceinline__ uint32_t bucket_ord(uint32_t rank, int nb, int root_pos) {
    return rank == (uint32_t)nb ? (uint32_t)root_pos : rank + (rank >= (uint32_t)root_pos ? 1u : 0u);
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
__global__ void tie_writeback_k(const uint32_t *__restrict__ c_pos, const uint32_t *__restrict__ sorted_idx, int64_t m,
                                uint32_t *__restrict__ perm) {
    int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c < m) perm[c_pos[c]] = sorted_idx[c];
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
}  // namespace scb
