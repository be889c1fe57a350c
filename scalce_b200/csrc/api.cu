// api.cu - handle, memory management and the C ABI (include/scalce_b200.h).
#include "../../include/scalce_b200.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <vector>

#include "common.cuh"
#include "core_table.h"
#include "pipeline.cuh"
#include "prims.cuh"
#include "resolve_dense.cuh"
#include "scan_smem2.cuh"
#include "scan_big.cuh"
#include "resolve_sparse.cuh"
#include "emit2.cuh"
#include "emit_offsets.cuh"
#include "emit_reads_fast.cuh"
#include "shard.cuh"
#include "container.cuh"
#include "parse.cuh"

namespace scb {
std::atomic<long long> g_launches{0};
static thread_local std::string g_last_error;

// Flush-scoped device memory: one slab (grown on demand, kept across flushes) with bump allocation.
// The transform allocates ~40 temporaries and 6 output streams per flush; going through the driver's
// pool for each of them costs milliseconds and occasionally stalls for hundreds, a slab costs nothing.
struct Arena {
    struct Slab { char *p; size_t cap; };
    std::vector<Slab> slabs;
    size_t cur = 0, off = 0;
    size_t peak = 0;            // high-water mark of the bytes handed out in one flush (scb_device_bytes)
    size_t used() const { size_t u = off; for (size_t k = 0; k < cur && k < slabs.size(); k++) u += slabs[k].cap; return u; }
    void reset() { cur = 0; off = 0; }
    void reserve(size_t bytes) {   // make the first slab at least this large (only ever called between flushes)
        if (!slabs.empty() && slabs[0].cap >= bytes) return;
        for (auto &s : slabs) cudaFree(s.p);
        slabs.clear();
        Slab s{nullptr, bytes};
        SCB_CUDA(cudaMalloc((void **)&s.p, bytes));
        slabs.push_back(s);
        reset();
    }
    void *alloc(size_t bytes) {
        bytes = (bytes + 511) & ~(size_t)511;
        if (bytes == 0) bytes = 512;
        while (true) {
            if (cur < slabs.size() && off + bytes <= slabs[cur].cap) { void *r = slabs[cur].p + off; off += bytes; if (used() > peak) peak = used(); return r; }
            if (cur + 1 < slabs.size()) { cur++; off = 0; continue; }
            Slab s{nullptr, std::max(bytes, (size_t)1 << 30)};
            SCB_CUDA(cudaMalloc((void **)&s.p, s.cap));
            slabs.push_back(s);
            cur = slabs.size() - 1; off = 0;
        }
    }
    struct Mark { size_t cur, off; };
    Mark mark() const { return Mark{cur, off}; }
    void rewind(Mark m) { cur = m.cur; off = m.off; }
    void destroy() { for (auto &s : slabs) cudaFree(s.p); slabs.clear(); reset(); }
};
static thread_local Arena *g_arena = nullptr;   // set while a flush runs

// device buffer: from the flush arena while a flush runs (never freed individually), otherwise from the
// stream-ordered pool
struct DevBuf {
    void *p = nullptr;
    size_t bytes = 0;
    cudaStream_t st = nullptr;
    bool pooled = false;
    DevBuf() {}
    DevBuf(size_t b, cudaStream_t s) { alloc(b, s); }
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    DevBuf(DevBuf &&o) noexcept { *this = std::move(o); }
    DevBuf &operator=(DevBuf &&o) noexcept {
        if (this != &o) { release(); p = o.p; bytes = o.bytes; st = o.st; pooled = o.pooled; o.p = nullptr; o.bytes = 0; }
        return *this;
    }
    void alloc(size_t b, cudaStream_t s) {
        release();
        st = s; bytes = b;
        if (g_arena) { p = g_arena->alloc(b); pooled = false; return; }
        // the row / name gathers fetch whole 16-byte granules (load16_unaligned): the granule holding an array's last byte must
        // belong to the allocation too - exact-size requests left up to 15 bytes of it outside (compute-sanitizer, emit_names_st_k)
        b = (b + 16 + 255) & ~(size_t)255;
        SCB_CUDA(cudaMallocAsync(&p, b, s));
        pooled = true;
    }
    void borrow(void *ptr, size_t b) { release(); p = ptr; bytes = b; pooled = false; }   // caller-owned memory
    void release() {
        if (p && pooled) cudaFreeAsync(p, st);
        p = nullptr; bytes = 0; pooled = false;
    }
    ~DevBuf() { release(); }
    template <typename T> T *as() const { return (T *)p; }
};

static int ceil_log2(uint64_t x) {  // bits needed to represent values in [0, x)
    int b = 0;
    while (b < 64 && (1ull << b) < x) b++;
    return b;
}

struct Pending {
    int64_t n = 0;
    bool borrowed = false;  // device pointers owned by the caller (location = 1)
    const uint8_t *seq1 = nullptr, *qual1 = nullptr, *names = nullptr, *seq2 = nullptr, *qual2 = nullptr;
    const int64_t *name_off = nullptr;
    int64_t name_bytes = 0;
    // sharded run, chunk ownership: rows [own_lo, own_hi) of qual1 / seq2 / qual2 are not in those arrays but in the rank's own
    // input (own_* pre-offset: row r at own_x + r * L); see RowSrc
    int64_t own_lo = 0, own_hi = 0;
    const uint8_t *own_qual1 = nullptr, *own_seq2 = nullptr, *own_qual2 = nullptr;
    DevBuf b_seq1, b_qual1, b_names, b_off, b_seq2, b_qual2;
};

struct EmitOut {
    DevBuf data[SCB_N_STREAMS];
    int64_t size[SCB_N_STREAMS] = {0, 0, 0, 0, 0, 0};
    std::vector<int64_t> chunk_off[SCB_N_STREAMS];
    int64_t n_seg = 0;
};

}  // namespace scb

using namespace scb;

struct scb_handle {
    scb_config cfg;
    CoreTable tab;
    cudaStream_t st = nullptr, st_own = nullptr;
    cudaStream_t st_aux[2] = {nullptr, nullptr};   // side streams: independent output kernels run next to the quality gather
    cudaEvent_t ev_fork = nullptr, ev_join[2] = {nullptr, nullptr};
    cudaEvent_t ev_s0 = nullptr, ev_s1 = nullptr;   // around a sharded send on the main stream
    cudaEvent_t ev_a0 = nullptr, ev_a1 = nullptr;   // around the row sends on the side stream (they may bracket a main-stream send)
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaEvent_t stage_ev[SCB_N_STAGES + 1] = {};
    float stage_ms[SCB_N_STAGES] = {};
    // device tables
    DevBuf d_next, d_nto, d_rank_level, d_rank_node_id, d_rank_core;
    DevBuf d_life, d_claim;
    DevBuf d_trans16, d_hit_rank;   // shared-memory form of the automaton (scan_smem.cuh)
    DevBuf d_trans32, d_hit_info;   // global-memory form for automata beyond shared memory (scan_big.cuh)
    DevBuf d_core_len, d_core_off, d_core_chars;   // the core set by core index (container.cuh: bucket records, inverse)
    DevBuf as_body, as_seg_core, as_seg_reads, as_seg_bytes, as_in_start;   // last scb_assemble_reads
    int64_t as_bytes = 0, as_nseg = 0;
    bool big_table = false;
    int H0 = 0, n_hit = 0;
    size_t smem_table_bytes = 0;
    std::vector<Pending> pending;
    Arena arena;
    // last flush
    Pending cur;
    DevBuf meta_in;               // per-read metadata word (emit2.cuh), input order
    DevBuf packed; int PW = 0;   // 2-bit packed mate-1 reads, PW words per read
    DevBuf lvl, ncand, cand_off, cand_rank, cand_pos, asg, endv, chunk, perm_keys, perm, perm_m;
    DevBuf dbg_bucket, dbg_core, dbg_end, dbg_chunk;
    EmitOut chunked, merged;
    int32_t n_chunks = 0;
    int64_t open_from = 0;        // first read of the last flush's open chunk (streaming flush)
    int64_t n_last = 0;
    int64_t unbucketed = 0;
    bool smem_resident = false;
    uint64_t life_total = 0;   // reads ever submitted (bound on any lifetime count)
    int last_rounds = 0;       // fixed-point rounds of the last dense resolve
    int64_t n_perm = 0;        // entries of perm (== n_last except after a sharded run)
    DevBuf srt_k0, srt_k1, srt_v1, srt_hist, srt_histws;   // sort buffers (shared by stage_sort and stage_emit's merged order)
    const uint64_t *srt_keys = nullptr;                    // sorted keys of the chunk-major order
    int srt_seg_bits = 0;
    bool srt_keys_valid = false;   // scb_shard_finish_sort ran: scb_shard_finish only emits
    Pending sh_local;          // the rank's own input after scb_shard_import replaced `cur` (phase-2 sends still read it)
    // ---- sharded run (scb_shard_*): state between the phases of one distributed flush -------------------
    int sh_phase = 0;          // 0 idle, 1 scanned, 2 resolved (finalized), 3 sized, 4 packed, 5 imported
    Arena::Mark sh_mark{0, 0}; // arena position after the arrays that survive the exchange
    int64_t sh_n_local = 0;    // reads of this rank's input shard
    int sh_W = 0, sh_grid = 0; // dense-resolve geometry
    DevBuf sh_sel, sh_base, sh_H, sh_S, sh_Csum, sh_Cpre, sh_changed, sh_blk, sh_stat, sh_tot;
    DevBuf sh_stale, sh_nstale;   // deferred re-sweeps (resolve_dense_k<*, *, true>)
    // (f2) input-order quality statistics of scb_submit_fastq: counts per mate, the two symbols before the next one (>= 256: none)
    DevBuf q_freq3[2], q_freq4[2];
    uint32_t q_prev[2][2] = {{500, 500}, {500, 500}};
    uint64_t q_seen[2] = {0, 0};
    // C++ orchestrator (scb_shard_flush): peers' receive arrays as mapped here, phase times of the last call
    std::map<std::pair<int, int>, std::pair<std::vector<uint8_t>, void *>> fl_ipc;   // (rank, array) -> (IPC handle bytes, mapped pointer)
    std::vector<std::vector<void *>> fl_table;                                       // [rank][array]
    float fl_ms[SCB_N_SHARD_PHASES] = {};
    float fl_wall[SCB_N_SHARD_PHASES] = {};   // host wall time per phase of the last scb_shard_flush (includes collectives and waiting)
    int fl_rounds = 0;
    // sparse resolve engine (resolve_sparse.cuh): bucket-major view of the candidate pairs, built once per flush
    int engine = 0;            // engine of the current flush: 0 dense, 1 sparse, 2 sequential
    bool sp_ready = false;
    int64_t sp_M = 0;
    DevBuf sp_doff, sp_sb, sp_sread, sp_sk, sp_sval, sp_cnt, sp_fbyte, sp_tail, sp_treset, sp_X, sp_hist, sp_changed, sp_base2, sp_dirty, sp_base_prev, sp_tile_clean, sp_ractive;
    uint32_t sp_dirty_tiles = 0;
    int sp_rank_bits = 24; int sp_nblk = 1;   // (block, bucket) keys of the sorted view
    std::vector<int64_t> sp_blk_first, sp_blk_pair, sp_blk_tile;           // [nblk + 1] first read / first pair / first tile slot of every block
    DevBuf sp_dblk_first;
    uint32_t sp_round = 0;      // rounds of the sparse engine since its set-up (the stamps in sp_dirty refer to it)
    DevBuf sh_S0, sh_H0, sh_frused, sh_frbuf, sh_fridx, sh_incr_stat;   // incremental resolve rounds (resolve_dense.cuh "fragile reads")
    DevBuf sh_sizes;           // u64 [n+1] exclusive prefix of rd.sz + 40 over the local shard
    DevBuf sh_perm, sh_aux, sh_packed, sh_qual1, sh_names, sh_seq2, sh_qual2, sh_noff;   // send side, destination-major
    std::vector<int64_t> sh_cnt_reads, sh_cnt_name_bytes, sh_first, sh_nbytes;
    int sh_G = 0;
    int64_t sh_layout[5] = {0, 0, 0, 0, 0};   // flush chunks of the local shard: first chunk id, new chunks, reads, reads of the first chunk, reads of the last
    int sh_split_mode = 0;                   // ownership of the last sharded flush: 0 bucket ranges, 1 flush chunks
    bool sh_resolved = false;                // scb_shard_finalize has run for the current shard
    bool sh_sized = false;                   // scb_shard_sizes has run for the current shard (global chunk ids exist)
    bool sh_aux_pending = false;             // chunk ownership, partitioned before the tie-break: the aux words are packed by the first send that carries them
    bool sh_rows_in_place = false;           // chunk ownership: this rank's own quality / mate-2 rows were not copied into its receive arrays
    int64_t sh_self_lo = 0; int sh_rank = 0;  // where this rank's reads start in its own receive numbering; its rank
    const uint8_t *sh_names_src = nullptr;   // names in send order: the staged copy (bucket ranges) or the input itself (flush chunks: send order = input order)
    void *rx[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // receive buffers: aux, packed, qual1, names, seq2, qual2
    size_t rx_cap[6] = {0, 0, 0, 0, 0, 0};
    std::vector<void *> rx_retired;
    DevBuf sh_name_off;        // import side: name offsets rebuilt from the lengths
    float sh_ms = 0;           // device time of the last scb_shard_* call
};

namespace scb {

static DfaDev dfa_of(scb_handle *h) {
    return DfaDev{h->d_next.as<uint32_t>(), h->d_nto.as<int32_t>(), h->d_rank_level.as<uint8_t>(), h->tab.n_states, h->tab.n_buckets};
}

template <typename T>
static void upload(DevBuf &b, const std::vector<T> &v, cudaStream_t st) {
    b.alloc(v.size() * sizeof(T), st);
    if (!v.empty()) SCB_CUDA(cudaMemcpyAsync(b.p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, st));
}

static int create_common(const std::vector<std::string> &cores, const scb_config *cfg, scb_handle **out) {
    if (!cfg || !out) { g_last_error = "null argument"; return SCB_EINVAL; }
    int L1 = cfg->read_length[0], L2 = cfg->read_length[1];
    // 2498 = the longest read the reference's line buffer holds (fgets into MAXLINE = 2500 bytes: text + newline + NUL, const.h:87)
    if (L1 <= 0 || L1 > SCB_MAX_READ_LENGTH || (cfg->paired && (L2 <= 0 || L2 > SCB_MAX_READ_LENGTH))) { g_last_error = "read length out of range (1..2498)"; return SCB_EINVAL; }
    if (cfg->bucket_set_bytes == 0) { g_last_error = "bucket_set_bytes must be > 0"; return SCB_EINVAL; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || cfg->device < 0 || cfg->device >= ndev) {
        cudaGetLastError();
        g_last_error = "no usable CUDA device (this library has no CPU path)";
        return SCB_ENODEVICE;
    }
    std::unique_ptr<scb_handle> h(new scb_handle());
    h->cfg = *cfg;
    if (!cfg->paired) h->cfg.read_length[1] = 0;
    std::string err = build_core_table(cores, h->tab);
    if (!err.empty()) { g_last_error = err; return SCB_EINVAL; }
    try {
        SCB_CUDA(cudaSetDevice(cfg->device));
        cudaDeviceProp prop;
        SCB_CUDA(cudaGetDeviceProperties(&prop, cfg->device));
        if (prop.major < 10) { g_last_error = "device is not sm_100 class"; return SCB_ENODEVICE; }
        SCB_CUDA(cudaStreamCreateWithFlags(&h->st_own, cudaStreamNonBlocking));
        h->st = h->st_own;
        for (auto &a : h->st_aux) SCB_CUDA(cudaStreamCreateWithFlags(&a, cudaStreamNonBlocking));
        SCB_CUDA(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
        for (auto &e : h->ev_join) SCB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        SCB_CUDA(cudaEventCreate(&h->ev_s0));
        SCB_CUDA(cudaEventCreate(&h->ev_s1));
        SCB_CUDA(cudaEventCreate(&h->ev_a0));
        SCB_CUDA(cudaEventCreate(&h->ev_a1));
        SCB_CUDA(cudaEventCreate(&h->ev0));
        SCB_CUDA(cudaEventCreate(&h->ev1));
        for (auto &e : h->stage_ev) SCB_CUDA(cudaEventCreate(&e));
        cudaMemPool_t pool;
        SCB_CUDA(cudaDeviceGetDefaultMemPool(&pool, cfg->device));
        uint64_t thr = ~0ull;
        SCB_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr));
        upload(h->d_next, h->tab.next, h->st);
        upload(h->d_nto, h->tab.nto_rank, h->st);
        upload(h->d_rank_level, h->tab.rank_level, h->st);
        upload(h->d_rank_node_id, h->tab.rank_node_id, h->st);
        upload(h->d_rank_core, h->tab.rank_core, h->st);
        {
            std::vector<int32_t> cl; std::vector<uint64_t> co; std::vector<uint8_t> cc;
            cl.reserve(h->tab.cores.size()); co.reserve(h->tab.cores.size() + 1);
            for (auto &c : h->tab.cores) { cl.push_back((int32_t)c.size()); co.push_back((uint64_t)cc.size()); cc.insert(cc.end(), c.begin(), c.end()); }
            co.push_back((uint64_t)cc.size());
            if (cl.empty()) cl.push_back(0);
            if (cc.empty()) cc.push_back(0);
            upload(h->d_core_len, cl, h->st); upload(h->d_core_off, co, h->st); upload(h->d_core_chars, cc, h->st);
            SCB_CUDA(cudaStreamSynchronize(h->st));
        }
        {   // scan tables: states renumbered so that states where some core ends come last (one compare per base finds a hit)
            const CoreTable &t = h->tab;
            const int ns = t.n_states;
            size_t nhit = 0;
            for (int u = 0; u < ns; u++) nhit += t.nto_rank[u] >= 0;
            std::vector<uint32_t> newid(ns);
            {
                uint32_t a = 0, b = (uint32_t)(ns - nhit);
                for (int u = 0; u < ns; u++) newid[u] = t.nto_rank[u] >= 0 ? b++ : a++;
            }
            h->H0 = (int)(ns - nhit); h->n_hit = (int)nhit;
            const size_t bytes = (size_t)ns * 8 + nhit * 4 + (size_t)t.n_buckets;
            const char *tf = getenv("SCB_TABLE");         // "global": keep even a small automaton in global memory (tests of scan_big_k)
            const bool force_global = tf && !strcmp(tf, "global");
            if (ns <= 16383 && bytes <= 160 * 1024 && !force_global) {   // shared-memory form: u16 entries hold next-state * 4
                std::vector<uint16_t> tr((size_t)ns * 4);
                std::vector<uint32_t> hr(nhit);
                for (int u = 0; u < ns; u++) {
                    for (int c = 0; c < 4; c++) tr[(size_t)newid[u] * 4 + c] = (uint16_t)(newid[t.next[(size_t)u * 4 + c]] * 4u);
                    if (t.nto_rank[u] >= 0) hr[newid[u] - (ns - nhit)] = (uint32_t)t.nto_rank[u];
                }
                upload(h->d_trans16, tr, h->st);
                upload(h->d_hit_rank, hr, h->st);
                h->smem_table_bytes = bytes;
                h->smem_resident = true;
            } else if (t.n_buckets < (1 << 24) && (uint64_t)ns < (1ull << kBigStateBits) && t.max_level < 64) {
                // global-memory form (scan_big.cuh): entry = next state | level of the longest core ending there << 26; hit_info = rank | level << 24
                std::vector<uint32_t> tr((size_t)ns * 4);
                std::vector<uint32_t> hi(nhit);
                for (int u = 0; u < ns; u++) {
                    for (int c = 0; c < 4; c++) {
                        const uint32_t v = t.next[(size_t)u * 4 + c];
                        const uint32_t lv = t.nto_rank[v] >= 0 ? (uint32_t)t.rank_level[(size_t)t.nto_rank[v]] : 0u;
                        tr[(size_t)newid[u] * 4 + c] = newid[v] | (lv << kBigStateBits);
                    }
                    if (t.nto_rank[u] >= 0) hi[newid[u] - (ns - nhit)] = (uint32_t)t.nto_rank[u] | ((uint32_t)t.rank_level[(size_t)t.nto_rank[u]] << 24);
                }
                upload(h->d_trans32, tr, h->st);
                upload(h->d_hit_info, hi, h->st);
                SCB_CUDA(cudaStreamSynchronize(h->st));   // the host vectors go out of scope
                h->big_table = true;
            }
        }
        size_t nb1 = (size_t)h->tab.n_buckets + 1;
        h->d_life.alloc(nb1 * 8, h->st);
        h->d_claim.alloc(nb1 * 4, h->st);
        SCB_CUDA(cudaMemsetAsync(h->d_life.p, 0, nb1 * 8, h->st));
        SCB_CUDA(cudaMemsetAsync(h->d_claim.p, 0xff, nb1 * 4, h->st));
        SCB_CUDA(cudaStreamSynchronize(h->st));
    } catch (CudaError &e) {
        g_last_error = e.msg;
        return SCB_ECUDA;
    }
    *out = h.release();
    return SCB_OK;
}

// ---- concatenate pending batches into h->cur -------------------------------------------------------
static void gather_pending(scb_handle *h) {
    auto &pv = h->pending;
    cudaStream_t st = h->st;
    const int L1 = h->cfg.read_length[0], L2 = h->cfg.read_length[1];
    if (pv.size() == 1) { h->cur = std::move(pv[0]); pv.clear(); return; }
    Pending c;
    for (auto &p : pv) { c.n += p.n; c.name_bytes += p.name_bytes; }
    c.b_seq1.alloc((size_t)c.n * L1, st);
    if (h->cfg.use_quals) c.b_qual1.alloc((size_t)c.n * L1, st);
    if (h->cfg.paired) { c.b_seq2.alloc((size_t)c.n * L2, st); if (h->cfg.use_quals) c.b_qual2.alloc((size_t)c.n * L2, st); }
    if (h->cfg.use_names) { c.b_names.alloc((size_t)c.name_bytes, st); c.b_off.alloc((size_t)(c.n + 1) * 8, st); }
    int64_t r0 = 0, nb0 = 0;
    std::vector<int64_t> host_off;
    for (auto &p : pv) {
        SCB_CUDA(cudaMemcpyAsync(c.b_seq1.as<uint8_t>() + r0 * L1, p.seq1, (size_t)p.n * L1, cudaMemcpyDeviceToDevice, st));
        if (h->cfg.use_quals) SCB_CUDA(cudaMemcpyAsync(c.b_qual1.as<uint8_t>() + r0 * L1, p.qual1, (size_t)p.n * L1, cudaMemcpyDeviceToDevice, st));
        if (h->cfg.paired) {
            SCB_CUDA(cudaMemcpyAsync(c.b_seq2.as<uint8_t>() + r0 * L2, p.seq2, (size_t)p.n * L2, cudaMemcpyDeviceToDevice, st));
            if (h->cfg.use_quals) SCB_CUDA(cudaMemcpyAsync(c.b_qual2.as<uint8_t>() + r0 * L2, p.qual2, (size_t)p.n * L2, cudaMemcpyDeviceToDevice, st));
        }
        if (h->cfg.use_names) {
            SCB_CUDA(cudaMemcpyAsync(c.b_names.as<uint8_t>() + nb0, p.names, (size_t)p.name_bytes, cudaMemcpyDeviceToDevice, st));
            // offsets were rebased to 0 per batch at submit; shift by nb0 on the host copy
            host_off.resize((size_t)p.n + 1);
            SCB_CUDA(cudaMemcpyAsync(host_off.data(), p.name_off, (size_t)(p.n + 1) * 8, cudaMemcpyDeviceToHost, st));
            SCB_CUDA(cudaStreamSynchronize(st));
            for (auto &o : host_off) o += nb0;
            SCB_CUDA(cudaMemcpyAsync(c.b_off.as<int64_t>() + r0, host_off.data(), (size_t)(p.n + 1) * 8, cudaMemcpyHostToDevice, st));
            SCB_CUDA(cudaStreamSynchronize(st));
        }
        r0 += p.n; nb0 += p.name_bytes;
    }
    c.seq1 = c.b_seq1.as<uint8_t>(); c.qual1 = c.b_qual1.as<uint8_t>(); c.seq2 = c.b_seq2.as<uint8_t>(); c.qual2 = c.b_qual2.as<uint8_t>();
    c.names = c.b_names.as<uint8_t>(); c.name_off = c.b_off.as<int64_t>();
    pv.clear();
    h->cur = std::move(c);
}

// dst row p <- src row perm[p], rows of L bytes (dst dense and 16-byte aligned)
static void gather_rows_any(cudaStream_t st, const RowSrc src, uint8_t *dst, const uint32_t *perm, int64_t n, int L) {
    if (n <= 0 || L <= 0) return;
    if (L >= 16 && src.split()) SCB_LAUNCH(gather_rows16_k<true>, (unsigned)cdiv(cdiv(n * L, 16), 256 * kGatherChunks), 256, 0, st, src, dst, perm, n, L);
    else if (L >= 16) SCB_LAUNCH(gather_rows16_k<false>, (unsigned)cdiv(cdiv(n * L, 16), 256 * kGatherChunks), 256, 0, st, src, dst, perm, n, L);
    else SCB_LAUNCH(gather_rows_small_k, (unsigned)cdiv(n * L, 256), 256, 0, st, src, dst, perm, n, L);
}

// ---- emit one ordering ----------------------------------------------------------------------------------
// keys: the sorted keys of this ordering; the segment id (chunk, bucket order) is their top seg_bits bits
// part: 1 = everything that does not read the quality / mate-2 rows (offsets, names, packed reads, meta records, per-chunk
// offsets; allocates all six streams), 2 = only the kernels that read those rows (streams 2, 4, 5), 3 = both, in the order
// the one-GPU flush has always used. The sharded run calls 1 while the rows are still crossing NVLink and 2 once they landed.
static void emit_order(scb_handle *h, const uint32_t *perm, const uint64_t *keys, int seg_shift, int seg_bits, bool merged, EmitOut &o, int part = 3) {
    cudaStream_t st = h->st;
    const Pending &c = h->cur;
    const int64_t n = c.n;
    const scb_config &cfg = h->cfg;
    const int L1 = cfg.read_length[0], L2 = cfg.read_length[1];
    const int nb = h->tab.n_buckets;
    const int sz_meta = L1 > 255 ? 2 : 1;
    const int nlen = 3 + 2 * cfg.paired;
    const int64_t rsz = 8 + 8 * nlen;
    const int nch = merged ? 1 : h->n_chunks;
    if (part == 2) {   // the row-dependent kernels alone, into the streams part 1 allocated
        if (n == 0) return;
        if (cfg.use_quals) gather_rows_any(st, RowSrc{c.qual1, c.own_qual1, c.own_lo, c.own_hi}, o.data[2].as<uint8_t>(), perm, n, L1);
        if (cfg.paired && cfg.use_quals) gather_rows_any(st, RowSrc{c.qual2, c.own_qual2, c.own_lo, c.own_hi}, o.data[5].as<uint8_t>(), perm, n, L2);
        if (cfg.paired) { const RowSrc rs{c.seq2, c.own_seq2, c.own_lo, c.own_hi}; if (rs.split()) SCB_LAUNCH(emit_reads2_k<true>, (unsigned)cdiv(n * ((sz_read(L2) + 3) / 4), 256), 256, 0, st, rs, perm, n, L2, o.data[4].as<uint8_t>()); else SCB_LAUNCH(emit_reads2_k<false>, (unsigned)cdiv(n * ((sz_read(L2) + 3) / 4), 256), 256, 0, st, rs, perm, n, L2, o.data[4].as<uint8_t>()); }
        return;
    }
    const bool rows_now = (part & 2) != 0;
    for (int k = 0; k < SCB_N_STREAMS; k++) { o.size[k] = 0; o.chunk_off[k].assign((size_t)std::max(nch, 0) + 1, 0); }
    o.n_seg = 0;
    if (n == 0) { for (int k = 0; k < SCB_N_STREAMS; k++) o.data[k].alloc(0, st); return; }

    DevBuf hsum((size_t)(n + 1) * 4, st);
    DevBuf ms((size_t)n * 8, st), offN((size_t)(n + 1) * 8, st), offR((size_t)(n + 1) * 8, st);
    KeyHead kh{keys, seg_shift, seg_bits};
    uint64_t totN = 0, totR = 0; uint32_t nseg = 0;
    {   // metadata gather + the three prefix sums in 3 launches (emit_offsets.cuh)
        const int64_t nt = scan_tiles(n);
        DevBuf ts((size_t)3 * nt * 8, st), tot3(32, st);
        SCB_LAUNCH(emit_off_reduce_k, (unsigned)nt, kScanThreads, 0, st, h->meta_in.as<uint64_t>(), perm, kh, n, L1, sz_meta, (int)cfg.use_names,
                   ms.as<uint64_t>(), ts.as<uint64_t>(), nt);
        SCB_LAUNCH(emit_off_sums_k, 3, 1024, 0, st, ts.as<uint64_t>(), nt, tot3.as<uint64_t>());
        SCB_LAUNCH(emit_off_apply_k, (unsigned)nt, kScanThreads, 0, st, ms.as<uint64_t>(), kh, n, L1, sz_meta, (int)cfg.use_names, ts.as<uint64_t>(), nt,
                   tot3.as<uint64_t>(), hsum.as<uint32_t>(), offN.as<uint64_t>(), offR.as<uint64_t>());
        if (cfg.use_names) SCB_CUDA(cudaMemcpyAsync(&totN, offN.as<uint64_t>() + n, 8, cudaMemcpyDeviceToHost, st));
    }
    SCB_CUDA(cudaMemcpyAsync(&totR, offR.as<uint64_t>() + n, 8, cudaMemcpyDeviceToHost, st));
    SCB_CUDA(cudaMemcpyAsync(&nseg, hsum.as<uint32_t>() + n, 4, cudaMemcpyDeviceToHost, st));
    SCB_CUDA(cudaStreamSynchronize(st));
    o.n_seg = nseg;
    DevBuf spos((size_t)(nseg + 1) * 4, st), srank((size_t)nseg * 4, st), schunk((size_t)nseg * 4, st), srec((size_t)nseg * 4, st);
    SegTab tab{spos.as<uint32_t>(), srank.as<uint32_t>(), schunk.as<uint32_t>(), srec.as<uint32_t>()};
    SCB_LAUNCH(seg_table_k, (unsigned)cdiv(n, 256), 256, 0, st, kh, hsum.as<uint32_t>(), n, perm, h->asg.as<uint32_t>(),
               (!merged && h->n_chunks > 1) ? h->chunk.as<uint32_t>() : (const uint32_t *)nullptr, h->d_rank_level.as<uint8_t>(), nb, L1, sz_meta, tab, nseg);
    o.size[SCB_S_NAMES] = (int64_t)totN;
    o.size[SCB_S_READS] = (int64_t)totR;
    o.size[SCB_S_QUALS] = cfg.use_quals ? n * L1 : 0;
    o.size[SCB_S_META] = (int64_t)nseg * rsz;
    o.size[SCB_S_READS2] = cfg.paired ? n * sz_read(L2) : 0;
    o.size[SCB_S_QUALS2] = (cfg.paired && cfg.use_quals) ? n * L2 : 0;
    for (int k = 0; k < SCB_N_STREAMS; k++) o.data[k].alloc((size_t)o.size[k] + 16, st);

    EmitMParams e;
    e.names = c.names; e.packed = h->packed.as<uint32_t>(); e.PW = h->PW; e.perm = perm; e.ms = ms.as<uint64_t>();
    e.offN = offN.as<uint64_t>(); e.offR = offR.as<uint64_t>(); e.n = n; e.L1 = L1; e.sz_meta = sz_meta;
    e.oN = o.data[0].as<uint8_t>(); e.oR = o.data[1].as<uint8_t>();
    uint8_t *oQ = o.data[2].as<uint8_t>(), *oR2 = o.data[4].as<uint8_t>(), *oQ2 = o.data[5].as<uint8_t>();
    auto gather_rows = [&](const RowSrc src, uint8_t *dst, int L) { gather_rows_any(st, src, dst, perm, n, L); };
    // The output kernels are independent of each other: names and packed reads (latency / issue bound) run on side
    // streams next to the quality-row gather (HBM bound) instead of one after the other.
    SCB_CUDA(cudaEventRecord(h->ev_fork, st));
    // part 1 alone: side stream 0 is carrying the sharded run's row sends, so the names stay on the main stream
    cudaStream_t sN = rows_now ? h->st_aux[0] : st, sR = h->st_aux[1];
    if (sN != st) SCB_CUDA(cudaStreamWaitEvent(sN, h->ev_fork, 0));
    SCB_CUDA(cudaStreamWaitEvent(sR, h->ev_fork, 0));
    if (cfg.use_names) SCB_LAUNCH(emit_names_st_k, (unsigned)cdiv(n, 256), 256, 0, sN, e);
    {   // packed reads + end markers (emit_reads_fast.cuh): rows staged per CTA, records assembled in shared memory
        const int recmax = sz_read(L1) + sz_meta;
        const int PWs = (h->PW + kEmitRowPad) | 1;
        const int per_read = recmax + PWs * 4;
        const int RPB = std::max(1, std::min(256, (40 * 1024) / per_read));
        const size_t smem = (((size_t)RPB * recmax + 48 + 15) & ~(size_t)15) + (size_t)RPB * PWs * 4 + 16;
        const uint32_t half = (uint32_t)((h->PW & 1) == 0 && (((uintptr_t)h->packed.p) & 7) == 0 ? h->PW / 2 : h->PW);
        const uint32_t inv_half = (uint32_t)(((1ull << 32) + half - 1) / half);     // half == 1: the kernel does not use it
        const int64_t n_blk = cdiv(n, RPB);
        SCB_CUDA(cudaFuncSetAttribute(emit_reads_fast_k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        SCB_LAUNCH(emit_reads_fast_k, (unsigned)n_blk, 256, smem, sR, e, RPB, inv_half, recmax, n_blk);
    }
    if (cfg.paired && rows_now) {
        const RowSrc rs{c.seq2, c.own_seq2, c.own_lo, c.own_hi};
        if (rs.split()) SCB_LAUNCH(emit_reads2_k<true>, (unsigned)cdiv(n * ((sz_read(L2) + 3) / 4), 256), 256, 0, sN, rs, perm, n, L2, oR2);
        else SCB_LAUNCH(emit_reads2_k<false>, (unsigned)cdiv(n * ((sz_read(L2) + 3) / 4), 256), 256, 0, sN, rs, perm, n, L2, oR2);
    }
    if (sN != st) SCB_CUDA(cudaEventRecord(h->ev_join[0], sN));
    SCB_CUDA(cudaEventRecord(h->ev_join[1], sR));
    if (cfg.use_quals && rows_now) gather_rows(RowSrc{c.qual1, c.own_qual1, c.own_lo, c.own_hi}, oQ, L1);
    if (cfg.paired && cfg.use_quals && rows_now) gather_rows(RowSrc{c.qual2, c.own_qual2, c.own_lo, c.own_hi}, oQ2, L2);
    if (sN != st) SCB_CUDA(cudaStreamWaitEvent(st, h->ev_join[0], 0));
    SCB_CUDA(cudaStreamWaitEvent(st, h->ev_join[1], 0));
    DevBuf cfirst((size_t)2 * std::max(nch, 1) * 8, st);
    SCB_CUDA(cudaMemsetAsync(cfirst.p, 0xff, (size_t)2 * std::max(nch, 1) * 8, st));   // -1 = chunk has no read here
    SCB_LAUNCH(meta2_k, (unsigned)cdiv(nseg, 128), 128, 0, st, tab, (int64_t)nseg, offN.as<uint64_t>(), offR.as<uint64_t>(),
               h->d_rank_node_id.as<int32_t>(), h->d_rank_core.as<int32_t>(), nb, L1, L2, cfg.use_names, cfg.use_quals, cfg.paired,
               o.data[3].as<uint8_t>(), merged ? (int64_t *)nullptr : cfirst.as<int64_t>(), nch);
    // per-chunk offsets of every stream
    if (!merged) {
        std::vector<int64_t> cf((size_t)2 * nch);
        SCB_CUDA(cudaMemcpyAsync(cf.data(), cfirst.p, cf.size() * 8, cudaMemcpyDeviceToHost, st));
        SCB_CUDA(cudaStreamSynchronize(st));
        // a chunk without reads (possible for a rank's bucket slice in a sharded run) is an empty slice
        for (int ci = nch - 1; ci >= 0; ci--)
            if (cf[ci] < 0) { cf[ci] = ci + 1 < nch ? cf[ci + 1] : n; cf[nch + ci] = ci + 1 < nch ? cf[nch + ci + 1] : (int64_t)nseg; }
        std::vector<uint64_t> on((size_t)nch, 0), orr((size_t)nch, 0);
        for (int ci = 0; ci < nch; ci++) {
            if (cfg.use_names) SCB_CUDA(cudaMemcpyAsync(&on[ci], offN.as<uint64_t>() + cf[ci], 8, cudaMemcpyDeviceToHost, st));
            SCB_CUDA(cudaMemcpyAsync(&orr[ci], offR.as<uint64_t>() + cf[ci], 8, cudaMemcpyDeviceToHost, st));
        }
        SCB_CUDA(cudaStreamSynchronize(st));
        for (int ci = 0; ci < nch; ci++) {
            int64_t p0 = cf[ci], m0 = cf[nch + ci];
            o.chunk_off[SCB_S_NAMES][ci] = (int64_t)on[ci];
            o.chunk_off[SCB_S_READS][ci] = (int64_t)orr[ci];
            o.chunk_off[SCB_S_QUALS][ci] = cfg.use_quals ? p0 * L1 : 0;
            o.chunk_off[SCB_S_META][ci] = m0 * rsz;
            o.chunk_off[SCB_S_READS2][ci] = cfg.paired ? p0 * sz_read(L2) : 0;
            o.chunk_off[SCB_S_QUALS2][ci] = (cfg.paired && cfg.use_quals) ? p0 * L2 : 0;
        }
    }
    for (int k = 0; k < SCB_N_STREAMS; k++) o.chunk_off[k][(size_t)nch] = o.size[k];
}

// ---- the transform ---------------------------------------------------------------------------------------
struct ArenaScope { ArenaScope(Arena *a) { g_arena = a; } ~ArenaScope() { g_arena = nullptr; } };

// sizes the flush slab and concatenates the pending batches into h->cur (call with the arena scope open)
static void flush_begin(scb_handle *h, double extra_factor) {
    const scb_config &cfg = h->cfg;
    const int L1 = cfg.read_length[0], L2 = cfg.read_length[1];
    {   // size the slab for this flush before anything is carved from it
        int64_t n_est = 0, name_est = 0;
        for (auto &p : h->pending) { n_est += p.n; name_est += p.name_bytes; }
        const size_t per_read = (size_t)((L1 + 15) / 16 * 4) + 1 + 2 + 8 + 8 * 6 + 4 + 2 + 4 + 3 * (8 + 4) + 8 + 4 + 12 /* fragile-read lists */ +
                                (size_t)(sz_read(L1) + 3) + (cfg.use_quals ? L1 : 0) + (cfg.paired ? sz_read(L2) + (cfg.use_quals ? L2 : 0) : 0) + 16 + 64;
        size_t est = (size_t)n_est * per_read + (size_t)name_est + (size_t)n_est + ((size_t)256 << 20);
        if (h->pending.size() > 1) est += (size_t)n_est * ((size_t)L1 * 2 + (size_t)L2 * 2 + 8) + (size_t)name_est;
        if (cfg.emit_merged) est += est / 2;
        est = (size_t)((double)est * extra_factor);
        h->arena.reserve(est);
        h->arena.reset();
        h->arena.peak = 0;
    }
    gather_pending(h);
    const Pending &c = h->cur;
    const int64_t n = c.n;
    h->n_last = n;
    if (c.name_bytes >= (1ll << 36)) throw CudaError{"more than 64 GiB of names in one flush (the per-read metadata word keeps 36 bits of name offset)"};
}

// 1. scan: max level + ordered distinct candidates per read (+ the 2-bit packed copy of the reads)
static void stage_scan(scb_handle *h) {
    cudaStream_t st = h->st;
    const scb_config &cfg = h->cfg;
    const int L1 = cfg.read_length[0], L2 = cfg.read_length[1];
    const Pending &c = h->cur;
    const int64_t n = c.n;
    const int nb = h->tab.n_buckets;
    (void)st; (void)cfg; (void)L1; (void)L2; (void)c; (void)n; (void)nb;
    h->PW = (L1 + 15) / 16;
    h->packed.alloc((size_t)n * h->PW * 4 + 64, st);   // slack: emit reads up to 2 words past a row
    h->lvl.alloc((size_t)n, st);
    h->ncand.alloc((size_t)n * 2, st);
    h->cand_off.alloc((size_t)(n + 1) * 8, st);
    DevBuf ws64((size_t)scan_tiles(n) * 8, st);
    DfaDev dfa = dfa_of(h);
    uint64_t M = 0;
    bool scanned = false;
    {
        const char *force = getenv("SCB_SCAN");
        const size_t budget = 227 * 1024 - 64;
        const int PW = h->PW, pitch = scan_smem_pitch(PW);
        const size_t per_warp = scan_smem_warp_bytes(L1, PW);
        const size_t fixed = h->smem_resident ? scan_smem_table_bytes(h->tab.n_states, h->n_hit, nb) : budget;
        int W = 0;
        if (h->smem_resident && fixed + 2 * per_warp <= budget) W = (int)std::min<size_t>(32, (budget - fixed) / per_warp);
        const int R = W * 32;
        if (W >= 2 && n > 0 && ((uintptr_t)c.seq1 & 15) == 0 && !(force && !strcmp(force, "global"))) {
            const size_t smem = scan_smem_total(h->tab.n_states, h->n_hit, nb, W, L1, PW);
            SCB_CUDA(cudaFuncSetAttribute(scan_smem2_k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            int dev_sms = 0;
            SCB_CUDA(cudaDeviceGetAttribute(&dev_sms, cudaDevAttrMultiProcessorCount, cfg.device));
            DevBuf dtot(8, st);
            // candidate space is handed to warps in chunks, so the arrays have holes: up to one chunk per warp
            const uint64_t holes = (uint64_t)dev_sms * W * kCandChunk;
            uint64_t cap = std::max<uint64_t>((uint64_t)n * 8, 1u << 20) + holes;
            for (int attempt = 0; attempt < 2 && !scanned; attempt++) {
                h->cand_rank.alloc((size_t)cap * 4, st);
                h->cand_pos.alloc((size_t)cap * 2, st);
                SCB_CUDA(cudaMemsetAsync(dtot.p, 0, 8, st));
                ScanSmemParams sp;
                sp.seq = c.seq1; sp.n = n; sp.L = L1; sp.trans = h->d_trans16.as<uint16_t>(); sp.hit_rank = h->d_hit_rank.as<uint32_t>();
                sp.rank_level = h->d_rank_level.as<uint8_t>(); sp.ns = h->tab.n_states; sp.n_hit = h->n_hit; sp.nb = nb; sp.H0 = h->H0; sp.R = R;
                sp.lvl = h->lvl.as<uint8_t>(); sp.ncand = h->ncand.as<uint16_t>(); sp.cand_off = h->cand_off.as<uint64_t>();
                sp.cand_rank = h->cand_rank.as<uint32_t>(); sp.cand_pos = h->cand_pos.as<uint16_t>();
                sp.cand_total = dtot.as<unsigned long long>(); sp.cand_cap = cap; sp.n_tiles = cdiv(n, 32);
                sp.packed = h->packed.as<uint32_t>(); sp.PW = PW;
                sp.inv_pw = (uint32_t)(((1ull << 32) + (uint64_t)PW - 1) / (uint64_t)PW); sp.pitch = pitch;
                int grid = (int)std::min<int64_t>(dev_sms, cdiv(sp.n_tiles, W));
                SCB_LAUNCH(scan_smem2_k, grid, R, smem, st, sp);
                SCB_CUDA(cudaMemcpyAsync(&M, dtot.p, 8, cudaMemcpyDeviceToHost, st));
                SCB_CUDA(cudaStreamSynchronize(st));
                if (M <= cap) scanned = true; else cap = M;   // list space was short: rerun with the exact size
            }
        }
    }
    if (!scanned && h->big_table && n > 0 && ((uintptr_t)c.seq1 & 3) == 0 && !(getenv("SCB_SCAN") && !strcmp(getenv("SCB_SCAN"), "global"))) {
        // automaton in global memory / L2 (scan_big.cuh): pack, then one lane per read walks with 32 warps per SM in flight
        const int PW = h->PW;
        const int Q = 32;            // per-lane queue of max-level hits; 16 (64 warps / SM instead of 32) measured no faster: the kernel is L2-rate bound
        const size_t smem = scan_big_smem_bytes(kBigThreads, Q);
        void (*kbig)(ScanBigParams) = Q == 16 ? scan_big_k<16> : scan_big_k<32>;
        SCB_CUDA(cudaFuncSetAttribute(kbig, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int dev_sms = 0;
        SCB_CUDA(cudaDeviceGetAttribute(&dev_sms, cudaDevAttrMultiProcessorCount, cfg.device));
        SCB_LAUNCH(pack_reads16_k, (unsigned)cdiv(n * PW, 256), 256, 0, st, c.seq1, n, L1, PW, h->packed.as<uint32_t>());
        DevBuf dtot(8, st);
        const int grid = (int)std::min<int64_t>((int64_t)dev_sms * (Q == 16 ? 4 : 2), cdiv(cdiv(n, 32), kBigThreads / 32));
        const uint64_t holes = (uint64_t)grid * (kBigThreads / 32) * kCandChunk;
        uint64_t cap = std::max<uint64_t>((uint64_t)n * 6, 1u << 20) + holes;
        for (int attempt = 0; attempt < 2 && !scanned; attempt++) {
            h->cand_rank.alloc((size_t)cap * 4, st);
            h->cand_pos.alloc((size_t)cap * 2, st);
            SCB_CUDA(cudaMemsetAsync(dtot.p, 0, 8, st));
            ScanBigParams bp;
            bp.packed = h->packed.as<uint32_t>(); bp.n = n; bp.L = L1; bp.PW = PW;
            bp.trans = h->d_trans32.as<uint32_t>(); bp.hit_info = h->d_hit_info.as<uint32_t>(); bp.H0 = (uint32_t)h->H0;
            bp.lvl = h->lvl.as<uint8_t>(); bp.ncand = h->ncand.as<uint16_t>(); bp.cand_off = h->cand_off.as<uint64_t>();
            bp.cand_rank = h->cand_rank.as<uint32_t>(); bp.cand_pos = h->cand_pos.as<uint16_t>();
            bp.cand_total = dtot.as<unsigned long long>(); bp.cand_cap = cap;
            SCB_LAUNCH(kbig, grid, kBigThreads, smem, st, bp);
            SCB_CUDA(cudaMemcpyAsync(&M, dtot.p, 8, cudaMemcpyDeviceToHost, st));
            SCB_CUDA(cudaStreamSynchronize(st));
            if (M <= cap) scanned = true; else cap = M;
        }
    }
    if (!scanned) {
        if (n > 0)
            SCB_LAUNCH((scan_k<false>), (unsigned)cdiv(n, 128), 128, 0, st, c.seq1, n, L1, dfa, h->lvl.as<uint8_t>(),
                       h->ncand.as<uint16_t>(), (const uint64_t *)nullptr, (uint32_t *)nullptr, (uint16_t *)nullptr);
        exclusive_scan<uint64_t>(LoadAs<uint16_t, uint64_t>{h->ncand.as<uint16_t>()}, n, h->cand_off.as<uint64_t>(),
                                 h->cand_off.as<uint64_t>() + n, ws64.as<uint64_t>(), st);
        SCB_CUDA(cudaMemcpyAsync(&M, h->cand_off.as<uint64_t>() + n, 8, cudaMemcpyDeviceToHost, st));
        SCB_CUDA(cudaStreamSynchronize(st));
        h->cand_rank.alloc((size_t)M * 4, st);
        h->cand_pos.alloc((size_t)M * 2, st);
        if (n > 0) {
            SCB_LAUNCH((scan_k<true>), (unsigned)cdiv(n, 128), 128, 0, st, c.seq1, n, L1, dfa, h->lvl.as<uint8_t>(),
                       h->ncand.as<uint16_t>(), h->cand_off.as<uint64_t>(), h->cand_rank.as<uint32_t>(), h->cand_pos.as<uint16_t>());
            SCB_LAUNCH(pack_reads_k, (unsigned)cdiv(n * h->PW, 256), 256, 0, st, c.seq1, n, L1, h->PW, h->packed.as<uint32_t>());
        }
    }

}

// ---- which tie-break engine serves this core set -------------------------------------------------------------
// dense (resolve_dense.cuh): per-warp population rows in shared memory, needs >= 4 warps' rows to fit (<= ~6.4k buckets);
// sparse (resolve_sparse.cuh): bucket-major pair lists in global memory, any bucket count < 2^24;
// sequential (resolve_seq_k): exact fallback with 64-bit counters (jobs of >= 2^32 - 1 reads), and the cross-check of the tests.
// SCB_RESOLVE=dense|sparse|seq forces one (tests run every engine on the same inputs).
enum { kEngDense = 0, kEngSparse = 1, kEngSeq = 2 };
static int dense_warps(const scb_handle *h) {
    const int nb1 = h->tab.n_buckets + 1;
    const int P = (nb1 + 3) & ~3;
    return (int)std::min<size_t>(kRdMaxWarps, (200 * 1024) / ((size_t)P * 8));
}
static int pick_engine(const scb_handle *h, uint64_t reads_in_job) {
    const char *force = getenv("SCB_RESOLVE");
    if (force && !strcmp(force, "seq")) return kEngSeq;
    if (reads_in_job >= 0xffffffffull) return kEngSeq;
    const int W = dense_warps(h);
    const bool sparse_ok = h->tab.n_buckets + 1 < (1 << 24);
    if (force && !strcmp(force, "sparse") && sparse_ok) return kEngSparse;
    if (force && !strcmp(force, "dense") && W >= 1) return kEngDense;
    if (W >= 4) return kEngDense;
    return sparse_ok ? kEngSparse : (W >= 1 ? kEngDense : kEngSeq);
}

// ---- sparse resolve engine (resolve_sparse.cuh) ------------------------------------------------------------------
// builds the bucket-major view of the candidate pairs of h->cur (once per flush)
static void sparse_setup(scb_handle *h, bool blocked) {
    cudaStream_t st = h->st;
    const int64_t n = h->cur.n;
    const int nb1 = h->tab.n_buckets + 1;
    h->sp_ready = false;
    h->sh_sel.alloc((size_t)std::max<int64_t>(n, 1) * 2, st);
    h->sh_base.alloc((size_t)nb1 * 4, st);
    h->sp_base2.alloc((size_t)nb1 * 4, st);
    h->sp_hist.alloc((size_t)(nb1 + 1) * 4, st);
    h->sp_changed.alloc(16, st);
    h->sp_dirty.alloc((size_t)nb1 * 4, st);
    h->sp_base_prev.alloc((size_t)nb1 * 4, st);
    SCB_CUDA(cudaMemsetAsync(h->sp_dirty.p, 0, (size_t)nb1 * 4, st));        // every bucket is dirty in round 0
    h->sp_round = 0;
    h->sh_tot.alloc((size_t)(nb1 + 1) * 4, st);
    h->sp_doff.alloc((size_t)(n + 1) * 8, st);
    DevBuf ws64((size_t)scan_tiles(n) * 8, st);
    uint64_t M = 0;
    exclusive_scan<uint64_t>(LoadAs<uint16_t, uint64_t>{h->ncand.as<uint16_t>()}, n, h->sp_doff.as<uint64_t>(), h->sp_doff.as<uint64_t>() + n, ws64.as<uint64_t>(), st);
    SCB_CUDA(cudaMemcpyAsync(&M, h->sp_doff.as<uint64_t>() + n, 8, cudaMemcpyDeviceToHost, st));
    SCB_CUDA(cudaStreamSynchronize(st));
    if (M >= 0xffffffffull) throw CudaError{"more than 2^32-1 candidate pairs in one flush: flush fewer reads at a time"};
    h->sp_M = (int64_t)M;
    const int64_t M1 = std::max<int64_t>((int64_t)M, 1);
    // input-order blocks (one GPU owning the order). The first blocks see (nearly) empty populations, where almost every
    // decision is a tie broken by the reads just before it and a block needs many rounds, so they start small and double up
    // to the steady size (measured at 50M reads x 1M cores: 19 rounds for a first block of 4.5M reads, 6-7 for the later
    // ones). The sharded rounds use one block.
    h->sp_rank_bits = std::max(1, ceil_log2((uint64_t)nb1));
    int64_t blk_max = 4 << 20, blk0 = 128 << 10;
    if (const char *e = getenv("SCB_SPARSE_BLOCK")) blk_max = std::max<int64_t>(256, atoll(e));      // reads per block (tests, experiments)
    if (const char *e = getenv("SCB_SPARSE_BLOCK0")) blk0 = std::max<int64_t>(256, atoll(e));
    blk0 = std::min(blk0, blk_max);
    std::vector<int64_t> first{0};
    if (blocked) {
        const int max_blocks = 1 << std::min(12, 31 - h->sp_rank_bits);      // block and bucket share a 31-bit key
        for (int64_t sz = blk0, at = sz; at < n && (int)first.size() < max_blocks; sz = std::min(sz * 2, blk_max), at += sz) first.push_back(at);
    }
    h->sp_nblk = (int)first.size();
    first.push_back(std::max<int64_t>(n, 0));
    h->sp_blk_first = first;
    h->sp_blk_pair.assign((size_t)h->sp_nblk + 1, 0);
    h->sp_dblk_first.alloc(first.size() * 8, st);
    {
        DevBuf dout(first.size() * 8, st);
        SCB_CUDA(cudaMemcpyAsync(h->sp_dblk_first.p, h->sp_blk_first.data(), first.size() * 8, cudaMemcpyHostToDevice, st));   // sp_blk_first outlives the copy
        SCB_LAUNCH(gather_u64_k, (unsigned)cdiv((int64_t)first.size(), 64), 64, 0, st, h->sp_doff.as<uint64_t>(), h->sp_dblk_first.as<int64_t>(), (int)first.size(), dout.as<int64_t>());
        SCB_CUDA(cudaMemcpyAsync(h->sp_blk_pair.data(), dout.p, first.size() * 8, cudaMemcpyDeviceToHost, st));
        SCB_CUDA(cudaStreamSynchronize(st));
    }
    // per-block tiles: every block's tile-indexed arrays (flag bytes, tails, clean marks) start at their own slot
    int64_t tiles = 0;
    h->sp_blk_tile.assign((size_t)h->sp_nblk + 1, 0);
    for (int k = 0; k < h->sp_nblk; k++) {
        h->sp_blk_tile[(size_t)k] = tiles;
        tiles += cdiv(std::max<int64_t>(h->sp_blk_pair[(size_t)k + 1] - h->sp_blk_pair[(size_t)k], 1), kSpTile);
    }
    h->sp_blk_tile[(size_t)h->sp_nblk] = tiles;
    h->sp_sb.alloc((size_t)M1 * 4, st); h->sp_sread.alloc((size_t)M1 * 4, st); h->sp_sk.alloc((size_t)M1 * 2, st); h->sp_sval.alloc((size_t)M1 * 4, st);
    h->sp_cnt.alloc((size_t)M1 * 4, st); h->sp_fbyte.alloc((size_t)tiles * (kSpTile / 8) + 16, st);
    h->sp_tail.alloc((size_t)tiles * 4, st); h->sp_treset.alloc((size_t)tiles * 4, st); h->sp_X.alloc((size_t)tiles * 4, st);
    h->sp_tile_clean.alloc((size_t)tiles, st); h->sp_ractive.alloc((size_t)std::max<int64_t>(n, 1), st);
    h->sp_dirty_tiles = (uint32_t)std::min<int64_t>(tiles, 0xffffffffll);   // "all dirty" until a round reports otherwise (the sharded rounds never do: no read-back per round)
    // temporaries of the sort: carved after the mark and handed back when the sorted view exists
    const Arena::Mark mk = h->arena.mark();
    {
        const int key_bits = h->sp_rank_bits + ceil_log2((uint64_t)h->sp_nblk);
        const int passes = M > 1 ? (key_bits + 7) / 8 : 0;      // radix_sort_pairs leaves 0 or 1 pairs where they are
        DevBuf k1((size_t)M1 * 4, st), v1((size_t)M1 * 4, st), pread((size_t)M1 * 4, st);
        DevBuf hist((size_t)SortWs::hist_elems(M1) * 4, st), histws((size_t)scan_tiles(SortWs::hist_elems(M1)) * 4, st);
        // the ping-pong starts on the side that makes the last pass land in sp_sb / sp_sval
        uint32_t *ka = (passes & 1) ? k1.as<uint32_t>() : h->sp_sb.as<uint32_t>(), *kb = (passes & 1) ? h->sp_sb.as<uint32_t>() : k1.as<uint32_t>();
        uint32_t *va = (passes & 1) ? v1.as<uint32_t>() : h->sp_sval.as<uint32_t>(), *vb = (passes & 1) ? h->sp_sval.as<uint32_t>() : v1.as<uint32_t>();
        if (n > 0)
            SCB_LAUNCH(sp_pairs_k, (unsigned)cdiv(n, 256), 256, 0, st, n, h->ncand.as<uint16_t>(), h->cand_off.as<uint64_t>(), h->cand_rank.as<uint32_t>(),
                       h->sp_doff.as<uint64_t>(), ka, va, pread.as<uint32_t>(), h->sh_sel.as<uint16_t>(), h->sp_dblk_first.as<int64_t>(), h->sp_nblk, h->sp_rank_bits);
        if (M > 0) {
            SortWs ws; ws.hist = hist.as<uint32_t>(); ws.tile_ws = histws.as<uint32_t>();
            radix_sort_pairs(&ka, &va, &kb, &vb, (int64_t)M, 0, key_bits, ws, st);
            if (ka != h->sp_sb.as<uint32_t>() || va != h->sp_sval.as<uint32_t>()) throw CudaError{"sparse set-up: sorted view landed in the wrong buffer"};
            SCB_LAUNCH(sp_post_k, (unsigned)cdiv((int64_t)M, 256), 256, 0, st, (int64_t)M, va, pread.as<uint32_t>(), h->sp_doff.as<uint64_t>(),
                       h->sp_sread.as<uint32_t>(), h->sp_sk.as<uint16_t>());
        }
    }
    if (g_arena) h->arena.rewind(mk);   // stream order keeps later users of this space behind the kernels above
    h->sp_ready = true;
}

// one round over input-order block `blk`: counts of the current assignment from `base` (the populations before the block's
// first read), then every read of the block re-decides. hist_mode: 0 none, 1 rebuild the local bucket histogram (sp_hist),
// 2 update it by the changes (sharded rounds: one block). The changed count lands in sp_changed.
static void sparse_round(scb_handle *h, const uint32_t *base, int hist_mode, int blk, bool first_of_block) {
    cudaStream_t st = h->st;
    const int64_t n = h->cur.n;
    const int nb1 = h->tab.n_buckets + 1;
    SCB_CUDA(cudaMemsetAsync(h->sp_changed.p, 0, 16, st));
    if (h->sp_M == 0 || n == 0) { if (hist_mode == 1) SCB_CUDA(cudaMemsetAsync(h->sp_hist.p, 0, (size_t)(nb1 + 1) * 4, st)); return; }
    const int64_t p0 = h->sp_blk_pair[(size_t)blk], M = h->sp_blk_pair[(size_t)blk + 1] - p0;
    const int64_t i0 = h->sp_blk_first[(size_t)blk], nr = h->sp_blk_first[(size_t)blk + 1] - i0;
    const int64_t t0 = h->sp_blk_tile[(size_t)blk];
    const uint32_t round = h->sp_round++;
    if (hist_mode == 1) SCB_CUDA(cudaMemsetAsync(h->sp_hist.p, 0, (size_t)(nb1 + 1) * 4, st));
    if (M == 0) return;                                        // no read of the block has a candidate: nothing to decide
    const int64_t tiles = cdiv(M, kSpTile);
    // reads are marked active (one more random store per rewritten count) only once few tiles are dirty; before that every read re-decides
    const bool use_active = !first_of_block && (uint64_t)h->sp_dirty_tiles * 4 < (uint64_t)tiles;
    if (hist_mode != 0) {   // sharded rounds: `base` moves between rounds; buckets whose value moved are dirty (first round: all are anyway)
        if (first_of_block) SCB_CUDA(cudaMemcpyAsync(h->sp_base_prev.p, base, (size_t)nb1 * 4, cudaMemcpyDeviceToDevice, st));
        else SCB_LAUNCH(sp_mark_base_k, (unsigned)cdiv(nb1, 256), 256, 0, st, base, h->sp_base_prev.as<uint32_t>(), nb1, h->sp_dirty.as<uint32_t>(), round);
    }
    SpRound r;
    r.dirty = h->sp_dirty.as<uint32_t>(); r.round = round; r.all = first_of_block ? 1u : 0u; r.rank_mask = (1u << h->sp_rank_bits) - 1u;
    r.tile_clean = h->sp_tile_clean.as<uint8_t>() + t0; r.n_dirty_tiles = h->sp_changed.as<uint32_t>() + 1;
    r.M = M; r.sb = h->sp_sb.as<uint32_t>() + p0; r.sread = h->sp_sread.as<uint32_t>() + p0; r.sval = h->sp_sval.as<uint32_t>() + p0; r.sk = h->sp_sk.as<uint16_t>() + p0;
    r.sel = h->sh_sel.as<uint16_t>(); r.fbyte = h->sp_fbyte.as<uint8_t>() + t0 * (kSpTile / 8); r.tail = h->sp_tail.as<uint32_t>() + t0; r.treset = h->sp_treset.as<uint32_t>() + t0;
    SCB_LAUNCH(sp_flags_k, (unsigned)tiles, kSpThreads, 0, st, r);
    SCB_LAUNCH(sp_tilescan_k, 1, 1024, 0, st, r.tail, r.treset, tiles, h->sp_X.as<uint32_t>() + t0);
    SpCounts c;
    c.M = M; c.sb = r.sb; c.sval = r.sval; c.fbyte = r.fbyte; c.X = h->sp_X.as<uint32_t>() + t0; c.base = base; c.cnt = h->sp_cnt.as<uint32_t>(); c.fold = nullptr;
    c.dirty = r.dirty; c.round = round; c.all = r.all; c.rank_mask = r.rank_mask; c.tile_clean = r.tile_clean; c.sread = r.sread;
    c.ractive = use_active ? h->sp_ractive.as<uint8_t>() : (uint8_t *)nullptr; c.stamp = (uint8_t)(round & 0xffu);
    SCB_LAUNCH(sp_counts_k, (unsigned)tiles, kSpThreads, 0, st, c);
    SCB_LAUNCH(sp_decide_k, (unsigned)cdiv(nr, 256), 256, 0, st, nr, h->ncand.as<uint16_t>() + i0, h->sp_doff.as<uint64_t>() + i0, h->sp_cnt.as<uint32_t>(),
               h->cand_off.as<uint64_t>() + i0, h->cand_rank.as<uint32_t>(), h->sh_sel.as<uint16_t>() + i0, h->sp_changed.as<uint32_t>(),
               hist_mode ? h->sp_hist.as<uint32_t>() : (uint32_t *)nullptr, hist_mode == 1 ? 1 : 0, h->sp_dirty.as<uint32_t>(), round + 1,
               h->sp_ractive.as<uint8_t>() + i0, use_active ? (int)(round & 0xffu) : -1);
}

// iterates the local reads to their fixed point from the populations in sh_base, block after block in input order, folding every
// finished block into the populations the next one starts from: sp_base2 = populations after the last read
static void sparse_local(scb_handle *h) {
    cudaStream_t st = h->st;
    const int nb1 = h->tab.n_buckets + 1;
    h->last_rounds = 0;
    SCB_CUDA(cudaMemcpyAsync(h->sp_base2.p, h->sh_base.p, (size_t)nb1 * 4, cudaMemcpyDeviceToDevice, st));
    if (h->sp_M == 0 || h->cur.n == 0) return;
    const bool prof = getenv("SCB_SPARSE_PROF") != nullptr;      // per-block rounds and device time, to stderr
    uint32_t *cur = h->sh_base.as<uint32_t>(), *nxt = h->sp_base2.as<uint32_t>();     // both hold the populations before block 0 here
    for (int blk = 0; blk < h->sp_nblk; blk++) {
        const int64_t p0 = h->sp_blk_pair[(size_t)blk], M = h->sp_blk_pair[(size_t)blk + 1] - p0;
        if (M == 0) continue;
        const int64_t tiles = cdiv(M, kSpTile), t0 = h->sp_blk_tile[(size_t)blk];
        h->sp_dirty_tiles = (uint32_t)std::min<int64_t>(tiles, 0xffffffffll);
        if (prof) SCB_CUDA(cudaEventRecord(h->ev_s0, st));
        int rounds = 0;
        while (true) {
            sparse_round(h, cur, 0, blk, rounds == 0);
            rounds++; h->last_rounds++;
            uint32_t cc[2] = {0, 0};
            SCB_CUDA(cudaMemcpyAsync(cc, h->sp_changed.p, 8, cudaMemcpyDeviceToHost, st));
            SCB_CUDA(cudaStreamSynchronize(st));
            h->sp_dirty_tiles = cc[1];
            if (cc[0] == 0) break;
            if (rounds >= kRdMaxRounds) throw CudaError{"resolve: round cap hit"};
        }
        // the last round changed nothing: its flags and tile scan describe the block's final assignment. Populations after the
        // block = populations before it, with the buckets the block touches overwritten (each once, at its segment's last pair)
        SpCounts c;
        c.M = M; c.sb = h->sp_sb.as<uint32_t>() + p0; c.sval = h->sp_sval.as<uint32_t>() + p0; c.fbyte = h->sp_fbyte.as<uint8_t>() + t0 * (kSpTile / 8);
        c.X = h->sp_X.as<uint32_t>() + t0; c.base = cur; c.cnt = nullptr; c.fold = nxt;
        c.dirty = h->sp_dirty.as<uint32_t>(); c.round = h->sp_round; c.all = 0; c.rank_mask = (1u << h->sp_rank_bits) - 1u;
        c.tile_clean = nullptr; c.sread = nullptr; c.ractive = nullptr; c.stamp = 0;
        SCB_LAUNCH(sp_counts_k, (unsigned)tiles, kSpThreads, 0, st, c);
        if (blk + 1 < h->sp_nblk) {   // ping-pong: the next block reads what this one folded
            SCB_CUDA(cudaMemcpyAsync(cur, nxt, (size_t)nb1 * 4, cudaMemcpyDeviceToDevice, st));
        }
        if (prof) {
            float rms = 0;
            SCB_CUDA(cudaEventRecord(h->ev_s1, st));
            SCB_CUDA(cudaEventSynchronize(h->ev_s1));
            SCB_CUDA(cudaEventElapsedTime(&rms, h->ev_s0, h->ev_s1));
            fprintf(stderr, "sparse block %3d of %d: %9lld pairs, %3d rounds, %8.3f ms\n", blk, h->sp_nblk, (long long)M, rounds, rms);
        }
    }
}

// ---- dense resolve engine (resolve_dense.cuh): geometry, buffers, launches --------------------------------
// Buffers live in the flush arena and are kept in the handle so that a sharded run can launch the kernel
// once per global round. Returns false when the engine does not apply (too many buckets for shared
// memory, counts that would overflow u32, or forced off).
static bool dense_setup(scb_handle *h, uint64_t reads_in_job /* upper bound on any population after this flush */) {
    cudaStream_t st = h->st;
    const int64_t n = h->cur.n;
    const int nb1 = h->tab.n_buckets + 1;
    const int P = (nb1 + 3) & ~3;   // row pitch: 128-bit rows
    int W = dense_warps(h);
    if (!(W >= 1 && n > 0 && reads_in_job < 0xffffffffull)) return false;
    int dev_sms = 0;
    SCB_CUDA(cudaDeviceGetAttribute(&dev_sms, cudaDevAttrMultiProcessorCount, h->cfg.device));
    const size_t smem = (size_t)W * P * 8;
    SCB_CUDA(cudaFuncSetAttribute(resolve_dense_k<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    SCB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, resolve_dense_k<true>, W * 32, smem));
    if (occ < 1) return false;
    const int grid = std::min(dev_sms, 160);
    const size_t max_sub = (size_t)grid * W;
    h->sh_W = W; h->sh_grid = grid;
    h->sh_sel.alloc((size_t)n * 2, st); h->sh_base.alloc((size_t)P * 4, st);
    h->sh_H.alloc(max_sub * P * 4, st);
    h->sh_S0.alloc(max_sub * P * 4, st); h->sh_H0.alloc(max_sub * P * 4, st);
    h->sh_frused.alloc(max_sub * 4, st);
    h->sh_frbuf.alloc((size_t)n * 8 + 64, st);
    h->sh_fridx.alloc((size_t)n * 4 + 64, st);
    h->sh_incr_stat.alloc(32, st);
    h->sh_stale.alloc(max_sub * 4, st); h->sh_nstale.alloc((size_t)kRdMaxRounds * 4, st);
    SCB_CUDA(cudaMemsetAsync(h->sh_stale.p, 0, max_sub * 4, st));
    SCB_CUDA(cudaMemsetAsync(h->sh_frused.p, 0xff, max_sub * 4, st));
    SCB_CUDA(cudaMemsetAsync(h->sh_incr_stat.p, 0, 32, st));
    h->sh_Csum.alloc((size_t)grid * P * 4, st); h->sh_Cpre.alloc((size_t)grid * P * 4, st);
    h->sh_changed.alloc((size_t)kRdMaxRounds * 4, st); h->sh_stat.alloc(8, st);
    h->sh_tot.alloc((size_t)(nb1 + 1) * 4, st);
    SCB_CUDA(cudaMemsetAsync(h->sh_base.p, 0, (size_t)P * 4, st));
    SCB_CUDA(cudaMemsetAsync(h->sh_sel.p, 0xff, (size_t)n * 2, st));
    return true;
}

// one launch of the engine over blocks `blk` of the local reads; returns status (0 ok)
static int dense_launch(scb_handle *h, int mode, int64_t g0, const std::vector<int64_t> &blk, uint32_t *tot_out) {
    cudaStream_t st = h->st;
    const int nb1 = h->tab.n_buckets + 1;
    const int W = h->sh_W, grid = h->sh_grid;
    const int P = (nb1 + 3) & ~3;
    const size_t smem = (size_t)W * P * 8;
    if (mode != 2) {   // a warm round reuses the block list of the round before
        h->sh_blk.alloc(blk.size() * 8, st);
        SCB_CUDA(cudaMemcpyAsync(h->sh_blk.p, blk.data(), blk.size() * 8, cudaMemcpyHostToDevice, st));
        SCB_CUDA(cudaMemsetAsync(h->sh_stat.p, 0, 8, st));
    }
    SCB_CUDA(cudaMemsetAsync(h->sh_changed.p, 0, (size_t)((mode == 0) ? kRdMaxRounds : 1) * 4, st));
    RdParams rp;
    rp.n = h->cur.n; rp.ncand = h->ncand.as<uint16_t>(); rp.cand_off = h->cand_off.as<uint64_t>(); rp.cand_rank = h->cand_rank.as<uint32_t>();
    rp.sel = h->sh_sel.as<uint16_t>(); rp.base = h->sh_base.as<uint32_t>(); rp.H = h->sh_H.as<uint32_t>(); rp.S = nullptr;
    rp.Csum = h->sh_Csum.as<uint32_t>(); rp.Cpre = h->sh_Cpre.as<uint32_t>(); rp.changed = h->sh_changed.as<uint32_t>(); rp.blk = h->sh_blk.as<int64_t>();
    rp.nblk = (int)blk.size() - 1; rp.nb1 = nb1; rp.W = W; rp.status = h->sh_stat.as<int>(); rp.rounds_out = h->sh_stat.as<int>() + 1;
    rp.g0 = g0; rp.mode = mode; rp.tot_out = tot_out; rp.pitch = P;
    rp.S0 = h->sh_S0.as<uint32_t>(); rp.H0 = h->sh_H0.as<uint32_t>(); rp.fr_buf = h->sh_frbuf.as<uint32_t>(); rp.fr_idx = h->sh_fridx.as<uint32_t>(); rp.fr_used = h->sh_frused.as<uint32_t>();
    rp.incr_T = nb1 > 0xffff ? 0 : 1024;             // decision-margin threshold of the incremental rounds (records hold bucket ranks in 16 bits)
    rp.incr_stat = (getenv("SCB_RESOLVE_PROF") || getenv("SCB_RESOLVE_STAT")) ? h->sh_incr_stat.as<unsigned long long>() : nullptr;
    rp.stale = h->sh_stale.as<uint32_t>(); rp.n_stale = h->sh_nstale.as<uint32_t>();
    if (mode == 0) SCB_CUDA(cudaMemsetAsync(h->sh_nstale.p, 0, (size_t)kRdMaxRounds * 4, st));
    DevBuf dts;
    const bool prof = mode == 0 && getenv("SCB_RESOLVE_PROF") != nullptr;
    if (prof) { dts.alloc(4096 * 8 * 8, st); SCB_CUDA(cudaMemsetAsync(dts.p, 0, 4096 * 8 * 8, st)); }
    rp.tstamps = prof ? dts.as<unsigned long long>() : nullptr;
    void *args[] = {&rp};
    // deferred re-sweeps (resolve_dense.cuh): 9.03 -> 8.43 ms at 50M x 150, profiles/r02_resolve_ab.txt
    void *kfn = (void *)resolve_dense_k<true>;
    SCB_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    SCB_CUDA(cudaLaunchCooperativeKernel(kfn, dim3(grid), dim3(W * 32), args, smem, st));
    g_launches++;
    if (mode != 0) {
        h->last_rounds++;
        if (getenv("SCB_RESOLVE_STAT")) {   // debugging aid: subtile sweeps of this round (synchronises)
            unsigned long long is[4] = {0, 0, 0, 0};
            SCB_CUDA(cudaStreamSynchronize(st));
            SCB_CUDA(cudaMemcpy(is, h->sh_incr_stat.p, 32, cudaMemcpyDeviceToHost));
            SCB_CUDA(cudaMemset(h->sh_incr_stat.p, 0, 32));
            fprintf(stderr, "round %d (device %d): %llu full, %llu incremental, %llu redone, %llu records\n", h->last_rounds, h->cfg.device, is[0], is[1], is[2], is[3]);
        }
        return 0;   // single round: nothing to read back, the stream orders the rest
    }
    int stat[2] = {0, 0};
    SCB_CUDA(cudaMemcpyAsync(stat, h->sh_stat.p, 8, cudaMemcpyDeviceToHost, st));
    SCB_CUDA(cudaStreamSynchronize(st));
    h->last_rounds = stat[1];
    if (prof) {
        std::vector<unsigned long long> ts(4096 * 8);
        SCB_CUDA(cudaMemcpy(ts.data(), dts.p, ts.size() * 8, cudaMemcpyDeviceToHost));
        double acc[6] = {0, 0, 0, 0, 0, 0};
        for (int r = 0; r < stat[1] && r < 4096; r++) {
            for (int k = 0; k < 6; k++) acc[k] += (double)(ts[r * 8 + k + 1] - ts[r * 8 + k]) * 1e-3;
            fprintf(stderr, "round %3d len %9llu: P %.1f D %.1f E %.1f sync1 %.1f scan %.1f sync2 %.1f us\n", r, ts[r * 8 + 7],
                    (ts[r * 8 + 1] - ts[r * 8]) * 1e-3, (ts[r * 8 + 2] - ts[r * 8 + 1]) * 1e-3, (ts[r * 8 + 3] - ts[r * 8 + 2]) * 1e-3,
                    (ts[r * 8 + 4] - ts[r * 8 + 3]) * 1e-3, (ts[r * 8 + 5] - ts[r * 8 + 4]) * 1e-3, (ts[r * 8 + 6] - ts[r * 8 + 5]) * 1e-3);
        }
        fprintf(stderr, "resolve totals (us): P %.0f D %.0f E %.0f sync1 %.0f scan %.0f sync2 %.0f\n", acc[0], acc[1], acc[2], acc[3], acc[4], acc[5]);
        unsigned long long is[4] = {0, 0, 0, 0};
        SCB_CUDA(cudaMemcpy(is, h->sh_incr_stat.p, 32, cudaMemcpyDeviceToHost));
        fprintf(stderr, "resolve subtile sweeps: %llu full, %llu incremental, %llu incremental redone in full, %llu records replayed (T = %d)\n", is[0], is[1], is[2], is[3], rp.incr_T);
    }
    return stat[0];
}

// geometric block schedule: block k+1 is as long as everything before it, earlier reads of the job included
static std::vector<int64_t> dense_blocks(int64_t n, int64_t g0) {
    std::vector<int64_t> blk;
    blk.push_back(0);
    const int64_t first = 4096;
    // early blocks are bound by the fixed cost of a round, not by their sweeps: they grow faster (x growth_small)
    // until the input before them reaches `small`
    const int64_t growth_small = 4, small = 1 << 20;     // x8 / x16 growth and later switch-over points were measured and lost (profiles/r02_resolve_ab.txt)
    while (blk.back() < n) {
        const int64_t n0 = blk.back(), before = n0 + g0;
        const int64_t len = before < small ? (growth_small - 1) * before : before;
        blk.push_back(std::min<int64_t>(n, n0 + std::max<int64_t>(len, first)));
    }
    return blk;
}

// asg / end marker from the converged slots; counts the reads without a candidate into the root population
static void dense_finalize(scb_handle *h) {
    cudaStream_t st = h->st;
    const int64_t n = h->cur.n;
    const int nb = h->tab.n_buckets;
    SCB_LAUNCH(resolve_finalize_k, (unsigned)cdiv(n, 256 * kFinPer), 256, 0, st, n, h->ncand.as<uint16_t>(), h->cand_off.as<uint64_t>(),
               h->cand_rank.as<uint32_t>(), h->cand_pos.as<uint16_t>(), h->sh_sel.as<uint16_t>(), nb, h->asg.as<uint32_t>(),
               h->endv.as<uint16_t>(), h->d_life.as<unsigned long long>() + nb);
}

// 2. resolve: the stateful tie-break over the whole flush, this GPU owning the whole input order
static void stage_resolve(scb_handle *h) {
    cudaStream_t st = h->st;
    const int64_t n = h->cur.n;
    const int nb = h->tab.n_buckets;
    const int nb1 = nb + 1;
    h->asg.alloc((size_t)n * 4, st);
    h->endv.alloc((size_t)n * 2, st);
    unsigned long long root_before = 0, root_after = 0;
    SCB_CUDA(cudaMemcpyAsync(&root_before, h->d_life.as<unsigned long long>() + nb, 8, cudaMemcpyDeviceToHost, st));
    bool dense_done = false;
    h->last_rounds = 0;
    h->engine = n > 0 ? pick_engine(h, h->life_total + (uint64_t)n) : kEngSeq;
    if (h->engine == kEngSparse) {
        sparse_setup(h, true);
        SCB_LAUNCH(life_to_base_k, (unsigned)cdiv(nb1, 256), 256, 0, st, h->d_life.as<unsigned long long>(), h->sh_base.as<uint32_t>(), nb1);
        sparse_local(h);
        SCB_LAUNCH(base_to_life_k, (unsigned)cdiv(nb1, 256), 256, 0, st, h->sp_base2.as<uint32_t>(), h->d_life.as<unsigned long long>(), nb);
        dense_finalize(h);
        dense_done = true;
    }
    if (h->engine == kEngDense && dense_setup(h, h->life_total + (uint64_t)n)) {
        const int64_t g0 = (int64_t)std::min<uint64_t>(h->life_total, (uint64_t)1 << 40);
        SCB_LAUNCH(life_to_base_k, (unsigned)cdiv(nb1, 256), 256, 0, st, h->d_life.as<unsigned long long>(), h->sh_base.as<uint32_t>(), nb1);
        if (dense_launch(h, 0, g0, dense_blocks(n, g0), nullptr) == 0) {
            SCB_LAUNCH(base_to_life_k, (unsigned)cdiv(nb1, 256), 256, 0, st, h->sh_base.as<uint32_t>(), h->d_life.as<unsigned long long>(), nb);
            dense_finalize(h);
            dense_done = true;
        }
    }
    if (!dense_done) {
        h->last_rounds = 0;
        size_t smem = ((size_t)nb + 1) * 12;
        if (smem <= 200 * 1024) {
            SCB_CUDA(cudaFuncSetAttribute(resolve_seq_k<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            SCB_LAUNCH((resolve_seq_k<true>), 1, 32, smem, st, n, h->ncand.as<uint16_t>(), h->cand_off.as<uint64_t>(),
                       h->cand_rank.as<uint32_t>(), h->cand_pos.as<uint16_t>(), h->d_life.as<unsigned long long>(),
                       h->d_claim.as<uint32_t>(), h->asg.as<uint32_t>(), h->endv.as<uint16_t>(), nb);
        } else {
            SCB_LAUNCH((resolve_seq_k<false>), 1, 32, 0, st, n, h->ncand.as<uint16_t>(), h->cand_off.as<uint64_t>(),
                       h->cand_rank.as<uint32_t>(), h->cand_pos.as<uint16_t>(), h->d_life.as<unsigned long long>(),
                       h->d_claim.as<uint32_t>(), h->asg.as<uint32_t>(), h->endv.as<uint16_t>(), nb);
        }
    }
    h->life_total += (uint64_t)n;
    SCB_CUDA(cudaMemcpyAsync(&root_after, h->d_life.as<unsigned long long>() + nb, 8, cudaMemcpyDeviceToHost, st));
    SCB_CUDA(cudaStreamSynchronize(st));
    if (h->tab.root_counts_unbucketed) h->unbucketed += (int64_t)(root_after - root_before);
}

// per-read metadata word for the output side (needs lvl, endv, name offsets)
static void stage_meta(scb_handle *h) {
    cudaStream_t st = h->st;
    const scb_config &cfg = h->cfg;
    const int L1 = cfg.read_length[0], L2 = cfg.read_length[1];
    const Pending &c = h->cur;
    const int64_t n = c.n;
    const int nb = h->tab.n_buckets;
    (void)st; (void)cfg; (void)L1; (void)L2; (void)c; (void)n; (void)nb;
    h->meta_in.alloc((size_t)n * 8, st);
    if (n > 0)
        SCB_LAUNCH(build_meta_k, (unsigned)cdiv(n, 256), 256, 0, st, n, cfg.use_names ? c.name_off : (const int64_t *)nullptr, h->lvl.as<uint8_t>(),
                   h->endv.as<uint16_t>(), h->meta_in.as<uint64_t>());

}

// 3. sizes -> flush chunks (compress.cpp:702, 708-713)
static void stage_chunks(scb_handle *h) {
    cudaStream_t st = h->st;
    const scb_config &cfg = h->cfg;
    const int L1 = cfg.read_length[0], L2 = cfg.read_length[1];
    const Pending &c = h->cur;
    const int64_t n = c.n;
    const int nb = h->tab.n_buckets;
    (void)st; (void)cfg; (void)L1; (void)L2; (void)c; (void)n; (void)nb;
    DevBuf ws64((size_t)scan_tiles(n) * 8, st);
    {
        int fixed = (cfg.use_quals ? L1 : 0) + (cfg.paired ? sz_read(L2) + (cfg.use_quals ? L2 : 0) : 0) + 40;
        RdSize rs{c.name_off, h->lvl.as<uint8_t>(), L1, fixed, cfg.use_names};
        DevBuf S((size_t)(n + 1) * 8, st);
        exclusive_scan<uint64_t>(rs, n, S.as<uint64_t>(), S.as<uint64_t>() + n, ws64.as<uint64_t>(), st);
        // upper bound on chunks: every read adds at most 256 + sz_read(L1) + L1 + mate 2 + 40 bytes
        uint64_t max_rd = 256 + (uint64_t)sz_read(L1) + L1 + sz_read(L2) + L2 + 40;
        uint64_t cap64 = (uint64_t)n * max_rd / cfg.bucket_set_bytes + 2;
        if (cap64 > (uint64_t)n + 1) cap64 = (uint64_t)n + 1;
        if (cap64 > (1u << 24)) throw CudaError{"bucket_set_bytes too small for this many reads (more than 2^24 flush chunks)"};
        int cap = (int)cap64;
        DevBuf cstart((size_t)cap * 4, st), dn(4, st);
        SCB_LAUNCH(chunk_bounds_k, 1, 1, 0, st, S.as<uint64_t>(), n, (uint64_t)cfg.bucket_set_bytes, cstart.as<uint32_t>(), cap, dn.as<int>());
        int nch = 0;
        long long open_from = 0;
        DevBuf dof(8, st);
        SCB_CUDA(cudaMemcpyAsync(&nch, dn.p, 4, cudaMemcpyDeviceToHost, st));
        SCB_CUDA(cudaStreamSynchronize(st));
        if (nch > cap) throw CudaError{"internal: chunk capacity exceeded"};
        SCB_LAUNCH(chunk_open_from_k, 1, 1, 0, st, S.as<uint64_t>(), n, (uint64_t)cfg.bucket_set_bytes, cstart.as<uint32_t>(), nch, dof.as<long long>());
        SCB_CUDA(cudaMemcpyAsync(&open_from, dof.p, 8, cudaMemcpyDeviceToHost, st));
        SCB_CUDA(cudaStreamSynchronize(st));
        h->open_from = open_from;
        h->n_chunks = nch;
        if (nch > 1) {
            h->chunk.alloc((size_t)n * 4, st);
            SCB_LAUNCH(chunk_ids_k, (unsigned)cdiv(n, 256), 256, 0, st, cstart.as<uint32_t>(), nch, n, h->chunk.as<uint32_t>());
        } else {
            h->chunk.release();
        }
    }

}

// 4-5. sort by (chunk, bucket order, suffix key), refine ties
static void stage_sort(scb_handle *h) {
    cudaStream_t st = h->st;
    const scb_config &cfg = h->cfg;
    const int L1 = cfg.read_length[0], L2 = cfg.read_length[1];
    const Pending &c = h->cur;
    const int64_t n = c.n;
    const int nb = h->tab.n_buckets;
    (void)st; (void)cfg; (void)L1; (void)L2; (void)c; (void)n; (void)nb;
    // 4. sort by (chunk, bucket order, key prefix), stable in input order
    const int nch = h->n_chunks > 0 ? h->n_chunks : 1;
    const int seg_bits = ceil_log2((uint64_t)nch * (uint64_t)(nb + 1));
    if (seg_bits > 40) throw CudaError{"too many (chunk, bucket) segments"};
    // key prefix: enough bases that equal prefixes are rare inside a SEGMENT (one bucket of one flush chunk: the
    // unit inside which the prefix has to discriminate) of the expected size, then rounded up so that the
    // sorted bit range is a whole number of 8-bit passes; equal prefixes that do occur are refined afterwards
    // (step 5), so this only trades radix passes against refinement work. Headline shape (50M x 150, 3 flush
    // chunks, 2049 buckets): 13 + 2*13 = 39 -> 40 bits, 5 passes (sizing per bucket instead gave 48 bits, 6 passes).
    int pb;
    {
        double per_bucket = (double)std::max<int64_t>(n, 1) / ((double)(nb + 1) * (double)nch);
        int need = 6;
        while (need < 32 && std::pow(4.0, need - 6) < per_bucket) need++;
        int total_bits = std::min(64, ((seg_bits + 2 * need + 7) / 8) * 8);
        pb = std::min(L1, (total_bits - seg_bits) / 2);
    }
    DevBuf &k0 = h->srt_k0, &k1 = h->srt_k1, &v1 = h->srt_v1, &hist = h->srt_hist, &histws = h->srt_histws;
    k0.alloc((size_t)n * 8, st); k1.alloc((size_t)n * 8, st); v1.alloc((size_t)n * 4, st);
    h->perm.alloc((size_t)n * 4, st);
    SortWs ws;
    hist.alloc((size_t)SortWs::hist_elems(n) * 4, st); histws.alloc((size_t)scan_tiles(SortWs::hist_elems(n)) * 4, st);
    ws.hist = hist.as<uint32_t>(); ws.tile_ws = histws.as<uint32_t>();
    uint64_t *ka = k0.as<uint64_t>(), *kb = k1.as<uint64_t>();
    uint32_t *va = h->perm.as<uint32_t>(), *vb = v1.as<uint32_t>();
    if (n > 0) {
        SCB_LAUNCH(build_keys_pk_k, (unsigned)cdiv(n, 256), 256, 0, st, h->packed.as<uint32_t>(), h->PW, n, h->asg.as<uint32_t>(), h->endv.as<uint16_t>(),
                   h->n_chunks > 1 ? h->chunk.as<uint32_t>() : (const uint32_t *)nullptr, nb, h->tab.root_order_pos, seg_bits, pb, ka, va);
        radix_sort_pairs(&ka, &va, &kb, &vb, n, 64 - seg_bits - 2 * pb, 64, ws, st);
        if (va != h->perm.as<uint32_t>()) {  // result landed in the alternate buffer
            SCB_CUDA(cudaMemcpyAsync(h->perm.p, va, (size_t)n * 4, cudaMemcpyDeviceToDevice, st));
            va = h->perm.as<uint32_t>();
        }
    }

    SCB_CUDA(cudaEventRecord(h->stage_ev[4], st));
    // 5. refine ties with further key bases until unique or the key is exhausted
    {
        const uint64_t *cur_keys = ka;
        const uint32_t *cur_idx = h->perm.as<uint32_t>();
        const uint32_t *cur_pos = nullptr;
        int64_t m = n;
        int consumed = pb;
        DevBuf keep_keys, keep_idx, keep_pos;  // own the current compact state across rounds
        while (consumed < L1 && m > 1) {
            uint32_t t = 0, G = 0;
            DevBuf c_pos, c_idx, c_grp;
            bool have = false;
            if (cur_pos == nullptr && m >= (1 << 16)) {
                // first round over the whole flush: tied elements are rare, so instead of two full prefix sums
                // collect their positions with warp-aggregated appends (unordered), then order the short list
                DevBuf lk0((size_t)m * 8 / 4 + 64, st), lk1((size_t)m * 8 / 4 + 64, st), lv0((size_t)m + 64, st), lv1((size_t)m + 64, st), dcnt(4, st);
                const uint32_t cap = (uint32_t)(m / 4);
                SCB_CUDA(cudaMemsetAsync(dcnt.p, 0, 4, st));
                SCB_LAUNCH(tie_find_k, (unsigned)cdiv(m, 256), 256, 0, st, cur_keys, m, lk0.as<uint64_t>(), lv0.as<uint32_t>(), cap, dcnt.as<uint32_t>());
                SCB_CUDA(cudaMemcpyAsync(&t, dcnt.p, 4, cudaMemcpyDeviceToHost, st));
                SCB_CUDA(cudaStreamSynchronize(st));
                if (t == 0) break;
                if (t <= cap) {
                    uint64_t *a = lk0.as<uint64_t>(), *b = lk1.as<uint64_t>();
                    uint32_t *x = lv0.as<uint32_t>(), *y = lv1.as<uint32_t>();
                    radix_sort_pairs(&a, &x, &b, &y, t, 0, ceil_log2((uint64_t)m), ws, st);
                    c_pos.alloc((size_t)t * 4, st); c_idx.alloc((size_t)t * 4, st); c_grp.alloc((size_t)t * 4, st);
                    DevBuf hs((size_t)(t + 1) * 4, st), w32((size_t)scan_tiles(t) * 4, st);
                    HeadAtPos hp{cur_keys, x};
                    exclusive_scan<uint32_t>(hp, t, hs.as<uint32_t>(), hs.as<uint32_t>() + t, w32.as<uint32_t>(), st);
                    SCB_CUDA(cudaMemcpyAsync(&G, hs.as<uint32_t>() + t, 4, cudaMemcpyDeviceToHost, st));
                    SCB_LAUNCH(tie_gather_sparse_k, (unsigned)cdiv(t, 256), 256, 0, st, hp, hs.as<uint32_t>(), cur_idx, (int64_t)t,
                               c_pos.as<uint32_t>(), c_idx.as<uint32_t>(), c_grp.as<uint32_t>());
                    SCB_CUDA(cudaStreamSynchronize(st));
                    have = true;
                    // groups of up to 32 members (all of them on ordinary data) are finished by one warp each, on the
                    // whole remaining key; only larger groups go through the iterative rounds below
                    DevBuf gstart((size_t)(G + 1) * 4, st), big((size_t)t, st), bpos((size_t)(t + 1) * 4, st), w32b((size_t)scan_tiles(t) * 4, st);
                    DevBuf mid_list((size_t)(t / 33 + 1) * 4, st), mid_count(4, st);
                    SCB_CUDA(cudaMemsetAsync(mid_count.p, 0, 4, st));
                    SCB_LAUNCH(tie_group_starts_k, (unsigned)cdiv(t, 256), 256, 0, st, c_grp.as<uint32_t>(), (int64_t)t, G, gstart.as<uint32_t>());
                    SCB_LAUNCH(tie_small_groups_k, (unsigned)cdiv((int64_t)G * 32, 256), 256, 0, st, h->packed.as<uint32_t>(), h->PW, L1, h->endv.as<uint16_t>(),
                               gstart.as<uint32_t>(), G, c_pos.as<uint32_t>(), c_idx.as<uint32_t>(), consumed, h->perm.as<uint32_t>(), big.as<uint8_t>(),
                               mid_list.as<uint32_t>(), mid_count.as<uint32_t>(), (uint32_t)kTieMidMax);
                    SCB_LAUNCH(tie_mid_groups_k, 148 * 8, 256, 0, st,   /* grid-stride over the queued groups; 8 CTAs of 256 threads fit an SM (16 KB static smem, 48 registers) */ h->packed.as<uint32_t>(), h->PW, L1, h->endv.as<uint16_t>(), gstart.as<uint32_t>(),
                               mid_list.as<uint32_t>(), mid_count.as<uint32_t>(), c_pos.as<uint32_t>(), c_idx.as<uint32_t>(), consumed, h->perm.as<uint32_t>());
                    exclusive_scan<uint32_t>(LoadAs<uint8_t, uint32_t>{big.as<uint8_t>()}, t, bpos.as<uint32_t>(), bpos.as<uint32_t>() + t, w32b.as<uint32_t>(), st);
                    uint32_t t2 = 0;
                    SCB_CUDA(cudaMemcpyAsync(&t2, bpos.as<uint32_t>() + t, 4, cudaMemcpyDeviceToHost, st));
                    SCB_CUDA(cudaStreamSynchronize(st));
                    if (t2 == 0) break;
                    DevBuf bk((size_t)t2 * 8, st), bi((size_t)t2 * 4, st), bp((size_t)t2 * 4, st);
                    SCB_LAUNCH(tie_big_compact_k, (unsigned)cdiv(t, 256), 256, 0, st, big.as<uint8_t>(), bpos.as<uint32_t>(), c_pos.as<uint32_t>(),
                               c_idx.as<uint32_t>(), c_grp.as<uint32_t>(), (int64_t)t, bk.as<uint64_t>(), bi.as<uint32_t>(), bp.as<uint32_t>());
                    keep_keys = std::move(bk); keep_idx = std::move(bi); keep_pos = std::move(bp);
                    cur_keys = keep_keys.as<uint64_t>(); cur_idx = keep_idx.as<uint32_t>(); cur_pos = keep_pos.as<uint32_t>();
                    m = t2;
                    continue;   // `consumed` is unchanged: the large groups start their rounds from the same key offset
                }
            }
            if (!have) {
                DevBuf flag((size_t)m, st), cpos((size_t)(m + 1) * 4, st), hsum((size_t)(m + 1) * 4, st), w32((size_t)scan_tiles(m) * 4, st);
                SCB_LAUNCH(tie_flags_k, (unsigned)cdiv(m, 256), 256, 0, st, cur_keys, m, flag.as<uint8_t>());
                exclusive_scan<uint32_t>(LoadAs<uint8_t, uint32_t>{flag.as<uint8_t>()}, m, cpos.as<uint32_t>(), cpos.as<uint32_t>() + m, w32.as<uint32_t>(), st);
                exclusive_scan<uint32_t>(HeadFlag{cur_keys, flag.as<uint8_t>()}, m, hsum.as<uint32_t>(), hsum.as<uint32_t>() + m, w32.as<uint32_t>(), st);
                SCB_CUDA(cudaMemcpyAsync(&t, cpos.as<uint32_t>() + m, 4, cudaMemcpyDeviceToHost, st));
                SCB_CUDA(cudaMemcpyAsync(&G, hsum.as<uint32_t>() + m, 4, cudaMemcpyDeviceToHost, st));
                SCB_CUDA(cudaStreamSynchronize(st));
                if (t == 0) break;
                c_pos.alloc((size_t)t * 4, st); c_idx.alloc((size_t)t * 4, st); c_grp.alloc((size_t)t * 4, st);
                SCB_LAUNCH(tie_compact_k, (unsigned)cdiv(m, 256), 256, 0, st, cur_keys, flag.as<uint8_t>(), cpos.as<uint32_t>(),
                           hsum.as<uint32_t>(), cur_pos, cur_idx, m, c_pos.as<uint32_t>(), c_idx.as<uint32_t>(), c_grp.as<uint32_t>());
            }
            int grp_bits = ceil_log2(G);
            int nbases = std::min(std::min(32, (64 - grp_bits) / 2), L1 - consumed);
            DevBuf nk0((size_t)t * 8, st), nk1((size_t)t * 8, st), nv0((size_t)t * 4, st), nv1((size_t)t * 4, st);
            uint64_t *a = nk0.as<uint64_t>(), *b = nk1.as<uint64_t>();
            uint32_t *x = nv0.as<uint32_t>(), *y = nv1.as<uint32_t>();
            SCB_LAUNCH(tie_rekey_pk_k, (unsigned)cdiv(t, 256), 256, 0, st, h->packed.as<uint32_t>(), h->PW, h->endv.as<uint16_t>(), c_idx.as<uint32_t>(),
                       c_grp.as<uint32_t>(), (int64_t)t, grp_bits, consumed, nbases, a, x);
            radix_sort_pairs(&a, &x, &b, &y, t, 64 - grp_bits - 2 * nbases, 64, ws, st);
            SCB_LAUNCH(tie_writeback_k, (unsigned)cdiv(t, 256), 256, 0, st, c_pos.as<uint32_t>(), x, (int64_t)t, h->perm.as<uint32_t>());
            // next round works on the compact state
            bool in0 = (a == nk0.as<uint64_t>());
            keep_keys = in0 ? std::move(nk0) : std::move(nk1);
            keep_idx = (x == nv0.as<uint32_t>()) ? std::move(nv0) : std::move(nv1);
            keep_pos = std::move(c_pos);
            cur_keys = keep_keys.as<uint64_t>(); cur_idx = keep_idx.as<uint32_t>(); cur_pos = keep_pos.as<uint32_t>();
            m = t; consumed += nbases;
        }
    }

    h->srt_keys = ka;
    h->srt_seg_bits = seg_bits;
}

// 6. emit streams per flush chunk (the t_%03d_k.tmp contents), then the merged order if asked
static void stage_emit(scb_handle *h, int part = 3) {
    cudaStream_t st = h->st;
    const scb_config &cfg = h->cfg;
    const int64_t n = h->cur.n;
    const int nb = h->tab.n_buckets;
    const int seg_bits = h->srt_seg_bits;
    const uint64_t *ka = h->srt_keys;
    DevBuf &k0 = h->srt_k0, &k1 = h->srt_k1, &v1 = h->srt_v1;
    SortWs ws;
    ws.hist = h->srt_hist.as<uint32_t>(); ws.tile_ws = h->srt_histws.as<uint32_t>();
    if (part == 2) {   // row-dependent kernels of both orderings (the sorts and everything else ran in part 1)
        emit_order(h, h->perm.as<uint32_t>(), ka, 64 - seg_bits, seg_bits, false, h->chunked, 2);
        if (cfg.emit_merged && h->n_chunks > 1) emit_order(h, h->perm_m.as<uint32_t>(), nullptr, 0, 0, true, h->merged, 2);
        return;
    }
    SCB_CUDA(cudaEventRecord(h->stage_ev[5], st));
    emit_order(h, h->perm.as<uint32_t>(), ka, 64 - seg_bits, seg_bits, false, h->chunked, part);
    SCB_CUDA(cudaEventRecord(h->stage_ev[6], st));
    if (cfg.emit_merged && h->n_chunks > 1) {
        // merge() concatenates a bucket's pieces in chunk order (compress.cpp:104-112): a stable sort of the
        // chunk-major order by bucket order gives exactly that
        const int ob = ceil_log2((uint64_t)nb + 1);
        h->perm_m.alloc((size_t)n * 4, st);
        uint64_t *a = k0.as<uint64_t>(), *b = k1.as<uint64_t>();
        uint32_t *x = h->perm_m.as<uint32_t>(), *y = v1.as<uint32_t>();
        if (n > 0)
            SCB_LAUNCH(merged_keys_k, (unsigned)cdiv(n, 256), 256, 0, st, h->asg.as<uint32_t>(), h->perm.as<uint32_t>(), n, nb,
                       h->tab.root_order_pos, a, x);
        radix_sort_pairs(&a, &x, &b, &y, n, 0, ob, ws, st);
        if (x != h->perm_m.as<uint32_t>()) SCB_CUDA(cudaMemcpyAsync(h->perm_m.p, x, (size_t)n * 4, cudaMemcpyDeviceToDevice, st));
        emit_order(h, h->perm_m.as<uint32_t>(), a, 0, ob, true, h->merged, part);
    }

}

// 7. per-read arrays in input order
static void stage_debug(scb_handle *h) {
    cudaStream_t st = h->st;
    const scb_config &cfg = h->cfg;
    const int L1 = cfg.read_length[0], L2 = cfg.read_length[1];
    const Pending &c = h->cur;
    const int64_t n = c.n;
    const int nb = h->tab.n_buckets;
    (void)st; (void)cfg; (void)L1; (void)L2; (void)c; (void)n; (void)nb;
    if (h->sh_phase == 0) { h->dbg_bucket.alloc((size_t)n * 4, st); h->dbg_core.alloc((size_t)n * 4, st); h->dbg_end.alloc((size_t)n * 4, st); h->dbg_chunk.alloc((size_t)n * 4, st); }
    if (n > 0) {
        SCB_LAUNCH(debug_arrays_k, (unsigned)cdiv(n, 256), 256, 0, st, n, h->asg.as<uint32_t>(), nb, h->d_rank_node_id.as<int32_t>(),
                   h->d_rank_core.as<int32_t>(), h->dbg_bucket.as<int32_t>(), h->dbg_core.as<int32_t>());
        SCB_LAUNCH(widen_u16_k, (unsigned)cdiv(n, 256), 256, 0, st, h->endv.as<uint16_t>(), n, h->dbg_end.as<int32_t>());
        if (h->chunk.p) SCB_CUDA(cudaMemcpyAsync(h->dbg_chunk.p, h->chunk.p, (size_t)n * 4, cudaMemcpyDeviceToDevice, st));
        else SCB_CUDA(cudaMemsetAsync(h->dbg_chunk.p, 0, (size_t)n * 4, st));
    }

}

// =====================================================================================================
// sharded run (include/scalce_b200.h "Sharded run")
// =====================================================================================================
struct ShardTimer {   // device time of one scb_shard_* call
    scb_handle *h;
    explicit ShardTimer(scb_handle *hh) : h(hh) { cudaEventRecord(h->ev0, h->st); }
    void stop() {
        SCB_CUDA(cudaEventRecord(h->ev1, h->st));
        SCB_CUDA(cudaStreamSynchronize(h->st));
        SCB_CUDA(cudaEventElapsedTime(&h->sh_ms, h->ev0, h->ev1));
    }
};

static void shard_scan(scb_handle *h) {
    cudaStream_t st = h->st;
    ArenaScope arena_scope(&h->arena);
    h->sh_phase = 0;
    flush_begin(h, 1.25);  // room for the send side (owner sort, aux words, staged names: ~50 B per read); the receive side reuses the slab after the exchange
    const int64_t n = h->cur.n;
    h->sh_n_local = n;
    h->n_perm = 0;
    // per-read arrays of the INPUT shard survive the exchange: carve them first and remember the mark
    h->dbg_bucket.alloc((size_t)n * 4, st); h->dbg_core.alloc((size_t)n * 4, st); h->dbg_end.alloc((size_t)n * 4, st); h->dbg_chunk.alloc((size_t)n * 4, st);
    h->sh_perm.alloc((size_t)std::max<int64_t>(n, 1) * 4, st);   // send order: read by the row sends that overlap the receive side
    h->sh_local = Pending();
    h->sh_mark = h->arena.mark();
    ShardTimer tm(h);
    stage_scan(h);
    h->asg.alloc((size_t)n * 4, st);
    h->endv.alloc((size_t)n * 2, st);
    {   // rd.sz prefix sums of the shard (compress.cpp:702): every rank at once, here; only the walk along the chunk boundaries
        // (scb_shard_sizes) has to wait for the ranks before it
        const scb_config &cfg = h->cfg;
        const int L1 = cfg.read_length[0], L2 = cfg.read_length[1];
        const int fixed = (cfg.use_quals ? L1 : 0) + (cfg.paired ? sz_read(L2) + (cfg.use_quals ? L2 : 0) : 0) + 40;
        RdSize rs{h->cur.name_off, h->lvl.as<uint8_t>(), L1, fixed, cfg.use_names};
        DevBuf ws64((size_t)scan_tiles(n) * 8, st);
        h->sh_sizes.alloc((size_t)(n + 1) * 8, st);
        exclusive_scan<uint64_t>(rs, n, h->sh_sizes.as<uint64_t>(), h->sh_sizes.as<uint64_t>() + n, ws64.as<uint64_t>(), st);
        h->chunk.alloc((size_t)std::max<int64_t>(n, 1) * 4, st);
    }
    tm.stop();
    h->last_rounds = 0;
    h->sh_resolved = false; h->sh_sized = false; h->sh_aux_pending = false; h->sh_names_src = nullptr; h->sh_rows_in_place = false;
    h->sh_phase = 1;
}

static void shard_sizes(scb_handle *h, uint64_t carry_in, int32_t chunk_in, uint64_t *carry_out, int32_t *chunk_out) {
    cudaStream_t st = h->st;
    ArenaScope arena_scope(&h->arena);
    const scb_config &cfg = h->cfg;
    const int L1 = cfg.read_length[0], L2 = cfg.read_length[1];
    const Pending &c = h->cur;
    const int64_t n = c.n;
    ShardTimer tm(h);
    const uint64_t max_rd = 256 + (uint64_t)sz_read(L1) + L1 + sz_read(L2) + L2 + 40;
    uint64_t cap64 = (uint64_t)n * max_rd / cfg.bucket_set_bytes + 2;
    if (cap64 > (uint64_t)n + 1) cap64 = (uint64_t)n + 1;
    if (cap64 > (1u << 24)) throw CudaError{"bucket_set_bytes too small for this many reads (more than 2^24 flush chunks)"};
    const int cap = (int)cap64;
    DevBuf bounds((size_t)cap * 4, st), dout(16, st);
    SCB_LAUNCH(chunk_bounds_carry_k, 1, 1, 0, st, h->sh_sizes.as<uint64_t>(), n, (uint64_t)cfg.bucket_set_bytes, carry_in,
               bounds.as<uint32_t>(), cap, dout.as<unsigned long long>());
    unsigned long long o[2] = {0, 0};
    SCB_CUDA(cudaMemcpyAsync(o, dout.p, 16, cudaMemcpyDeviceToHost, st));
    SCB_CUDA(cudaStreamSynchronize(st));
    if ((int64_t)o[0] > cap) throw CudaError{"internal: chunk capacity exceeded"};
    if ((uint64_t)chunk_in + o[0] >= kAuxMaxChunks) throw CudaError{"too many flush chunks for the sharded run (>= 2^20)"};
    if (n > 0) SCB_LAUNCH(chunk_ids_global_k, (unsigned)cdiv(n, 256), 256, 0, st, bounds.as<uint32_t>(), (int)o[0], (uint32_t)chunk_in, n, h->chunk.as<uint32_t>());
    {   // what an orchestrator needs to hand whole chunks to ranks (scb_shard_chunk_layout)
        uint32_t b2[2] = {0, 0};
        if (o[0] > 0) {
            SCB_CUDA(cudaMemcpyAsync(&b2[0], bounds.as<uint32_t>(), 4, cudaMemcpyDeviceToHost, st));
            SCB_CUDA(cudaMemcpyAsync(&b2[1], bounds.as<uint32_t>() + (o[0] - 1), 4, cudaMemcpyDeviceToHost, st));
            SCB_CUDA(cudaStreamSynchronize(st));
        }
        h->sh_layout[0] = chunk_in; h->sh_layout[1] = (int64_t)o[0]; h->sh_layout[2] = n;
        h->sh_layout[3] = o[0] > 0 ? (int64_t)b2[0] : n;
        h->sh_layout[4] = o[0] > 0 ? n - (int64_t)b2[1] : 0;
    }
    tm.stop();
    h->sh_sized = true;
    *carry_out = o[1];
    *chunk_out = chunk_in + (int32_t)o[0];
}

// the engine of a sharded run depends on the core set alone, so that every rank takes the same one
static int shard_engine(scb_handle *h) {
    const int e = pick_engine(h, 0);
    if (e == kEngSeq) throw CudaError{"the sharded run needs the dense or the sparse resolve engine (SCB_RESOLVE=seq, or >= 2^24 buckets)"};
    return e;
}
static void shard_need_dense(scb_handle *h) {
    if (!dense_setup(h, h->life_total + (uint64_t)h->cur.n))
        throw CudaError{"the sharded run's dense resolve engine does not apply (more than 2^32-1 reads in one job?)"};
}

static void shard_resolve_local(scb_handle *h, uint32_t *tot_dev) {
    cudaStream_t st = h->st;
    ArenaScope arena_scope(&h->arena);
    const int64_t n = h->cur.n;
    const int nb1 = h->tab.n_buckets + 1;
    ShardTimer tm(h);
    if (n == 0) {
        SCB_CUDA(cudaMemsetAsync(tot_dev, 0, (size_t)(nb1 + 1) * 4, st));
    } else {
        h->engine = shard_engine(h);
        if ((h->life_total + (uint64_t)n) >= 0xffffffffull) throw CudaError{"more than 2^32-1 reads in one job: u32 resolve counters would overflow"};
        if (h->engine == kEngSparse) {
            sparse_setup(h, true);
            SCB_LAUNCH(life_to_base_k, (unsigned)cdiv(nb1, 256), 256, 0, st, h->d_life.as<unsigned long long>(), h->sh_base.as<uint32_t>(), nb1);
            sparse_local(h);
            SCB_LAUNCH(sub_life_k, (unsigned)cdiv(nb1 + 1, 256), 256, 0, st, h->sp_base2.as<uint32_t>(), h->d_life.as<unsigned long long>(), nb1, tot_dev);
        } else {
            shard_need_dense(h);
            const int64_t g0 = (int64_t)std::min<uint64_t>(h->life_total, (uint64_t)1 << 40);
            SCB_LAUNCH(life_to_base_k, (unsigned)cdiv(nb1, 256), 256, 0, st, h->d_life.as<unsigned long long>(), h->sh_base.as<uint32_t>(), nb1);
            if (dense_launch(h, 0, g0, dense_blocks(n, g0), nullptr) != 0) throw CudaError{"resolve: round cap hit"};
            SCB_LAUNCH(sub_life_k, (unsigned)cdiv(nb1 + 1, 256), 256, 0, st, h->sh_base.as<uint32_t>(), h->d_life.as<unsigned long long>(), nb1, tot_dev);
        }
    }
    tm.stop();
}

static void shard_resolve_round(scb_handle *h, const uint32_t *before_dev, int64_t reads_before, int first, uint32_t *tot_dev) {
    cudaStream_t st = h->st;
    ArenaScope arena_scope(&h->arena);
    const int64_t n = h->cur.n;
    const int nb1 = h->tab.n_buckets + 1;
    if (n == 0) {
        SCB_CUDA(cudaMemsetAsync(tot_dev, 0, (size_t)(nb1 + 1) * 4, st));
    } else {
        if (first) h->engine = shard_engine(h);
        if ((h->life_total + (uint64_t)reads_before + (uint64_t)n) >= 0xffffffffull) throw CudaError{"more than 2^32-1 reads in one job: u32 resolve counters would overflow"};
        if (h->engine == kEngSparse) {
            // one global round on the bucket-major pair lists: base = lifetime counts + the lower ranks' histograms under the
            // current assignment; the local histogram follows the decisions (rebuilt in the first round, then kept by differences)
            if (first) sparse_setup(h, false);
            SCB_LAUNCH(add_life_k, (unsigned)cdiv(nb1, 256), 256, 0, st, h->d_life.as<unsigned long long>(), before_dev, nb1, h->sh_base.as<uint32_t>());
            sparse_round(h, h->sh_base.as<uint32_t>(), first ? 1 : 2, 0, first != 0);
            SCB_LAUNCH(sp_copy_tot_k, (unsigned)cdiv(nb1 + 1, 256), 256, 0, st, h->sp_hist.as<uint32_t>(), h->sp_changed.as<uint32_t>(), nb1, tot_dev);
            h->last_rounds++;
        } else {
            if (first) shard_need_dense(h);
            SCB_LAUNCH(add_life_k, (unsigned)cdiv(nb1, 256), 256, 0, st, h->d_life.as<unsigned long long>(), before_dev, nb1, h->sh_base.as<uint32_t>());
            const int64_t g0 = (int64_t)std::min<uint64_t>(h->life_total, (uint64_t)1 << 40) + reads_before;
            std::vector<int64_t> blk{0, n};
            if (dense_launch(h, first ? 1 : 2, g0, blk, tot_dev) != 0) throw CudaError{"resolve: round cap hit"};
        }
    }
    h->sh_ms = 0;   // asynchronous: the caller times the round loop on its stream
}

static void shard_finalize(scb_handle *h, const uint32_t *global_tot_dev, int64_t n_global) {
    cudaStream_t st = h->st;
    ArenaScope arena_scope(&h->arena);
    const int64_t n = h->cur.n;
    const int nb = h->tab.n_buckets;
    ShardTimer tm(h);
    unsigned long long root_before = 0, root_after = 0;
    SCB_CUDA(cudaMemcpyAsync(&root_before, h->d_life.as<unsigned long long>() + nb, 8, cudaMemcpyDeviceToHost, st));
    if (n > 0) dense_finalize(h);
    SCB_LAUNCH(life_add_k, (unsigned)cdiv(nb + 1, 256), 256, 0, st, h->d_life.as<unsigned long long>(), global_tot_dev, nb);
    SCB_CUDA(cudaMemcpyAsync(&root_after, h->d_life.as<unsigned long long>() + nb, 8, cudaMemcpyDeviceToHost, st));
    h->life_total += (uint64_t)n_global;
    stage_debug(h);   // per-read arrays of the input shard (chunk ids are the global ones)
    tm.stop();
    if (h->tab.root_counts_unbucketed) h->unbucketed += (int64_t)(root_after - root_before);
    h->sh_resolved = true;
    h->sh_phase = std::max(h->sh_phase, 2);     // chunk ownership may have partitioned (3) and started its row sends (4) already
}

static void shard_bucket_hist(scb_handle *h, uint32_t *hist_dev) {
    cudaStream_t st = h->st;
    const int64_t n = h->cur.n;
    const int nb = h->tab.n_buckets;
    ShardTimer tm(h);
    SCB_CUDA(cudaMemsetAsync(hist_dev, 0, (size_t)(nb + 1) * 4, st));
    if (n > 0) {
        const size_t smem = (size_t)(nb + 1) * 4;
        const int grid = (int)std::min<int64_t>(cdiv(n, 512 * 8), 148 * 4);
        if (smem <= 160 * 1024) {
            SCB_CUDA(cudaFuncSetAttribute(bucket_hist_k<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            SCB_LAUNCH((bucket_hist_k<true>), grid, 512, smem, st, h->asg.as<uint32_t>(), n, nb, h->tab.root_order_pos, hist_dev);
        } else {
            SCB_LAUNCH((bucket_hist_k<false>), grid, 512, 0, st, h->asg.as<uint32_t>(), n, nb, h->tab.root_order_pos, hist_dev);
        }
    }
    tm.stop();
}

// dst row p <- src row perm[p] where dst is any byte address (peer memory over NVLink or local)
static void gather_rows_to(cudaStream_t st, const uint8_t *src, uint8_t *dst, const uint32_t *perm, int64_t n, int L, int grid_cap = 0) {
    if (n <= 0 || L <= 0) return;
    if (L >= 16) {
        const int64_t nchunks = (n * (int64_t)L + 15 + 15) >> 4;
        int64_t grid = cdiv(nchunks, 256 * kGatherChunks);
        if (grid_cap > 0 && grid > grid_cap) grid = grid_cap;
        SCB_LAUNCH(gather_rows16_to_k, (unsigned)grid, 256, 0, st, src, dst, perm, n, L);
    } else if ((L & 3) == 0 && (((uintptr_t)src | (uintptr_t)dst) & 3) == 0) {
        SCB_LAUNCH(gather_words_to_k, (unsigned)cdiv(n * (L / 4), 256), 256, 0, st, (const uint32_t *)src, (uint32_t *)dst, perm, n, L / 4);
    } else {
        SCB_LAUNCH(gather_rows_small_k, (unsigned)cdiv(n * L, 256), 256, 0, st, RowSrc{src, nullptr, 0, 0}, dst, perm, n, L);
    }
}

// stable partition of the local reads by owner + what the owners need to size their receive buffers
// split != null: owner = the rank whose slice of the bucket emission order holds the read's bucket; else chunk_owner[n_chunks]
// (host): owner = the rank that was handed the read's flush chunk
static void shard_partition(scb_handle *h, const int64_t *split, int G, scb_shard_xfer *out, const int32_t *chunk_owner = nullptr, int n_chunks = 0) {
    cudaStream_t st = h->st;
    ArenaScope arena_scope(&h->arena);
    const scb_config &cfg = h->cfg;
    const Pending &c = h->cur;
    const int64_t n = c.n;
    const int nb = h->tab.n_buckets;
    const int64_t n1 = std::max<int64_t>(n, 1);
    ShardTimer tm(h);
    SplitTab t;
    t.G = G;
    if (split) for (int g = 0; g <= G; g++) t.s[g] = (uint32_t)split[g];
    // one 8-bit radix pass over the destination keeps input order inside every destination
    DevBuf k0((size_t)n1 * 8, st), k1((size_t)n1 * 8, st), v0((size_t)n1 * 4, st);
    DevBuf hist((size_t)SortWs::hist_elems(n) * 4, st), histws((size_t)scan_tiles(SortWs::hist_elems(n)) * 4, st);
    SortWs ws; ws.hist = hist.as<uint32_t>(); ws.tile_ws = histws.as<uint32_t>();
    uint64_t *ka = k0.as<uint64_t>(), *kb = k1.as<uint64_t>();
    uint32_t *va = v0.as<uint32_t>(), *vb = h->sh_perm.as<uint32_t>();   // one pass: the result lands in sh_perm
    if (n > 0 && !split) {
        // whole chunks: owners are non-decreasing along the input, the send order is the input order
        std::vector<uint8_t> ow((size_t)std::max(n_chunks, 1), 0);
        for (int c2 = 0; c2 < n_chunks; c2++) ow[(size_t)c2] = (uint8_t)chunk_owner[c2];
        DevBuf dow(ow.size(), st);
        SCB_CUDA(cudaMemcpyAsync(dow.p, ow.data(), ow.size(), cudaMemcpyHostToDevice, st));
        SCB_LAUNCH(dest_keys_chunk_k, (unsigned)cdiv(n, 256), 256, 0, st, h->chunk.as<uint32_t>(), n, dow.as<uint8_t>(), ka, h->sh_perm.as<uint32_t>());
        SCB_CUDA(cudaStreamSynchronize(st));          // `ow` leaves scope
    } else if (n > 0) {
        SCB_LAUNCH(dest_keys_k, (unsigned)cdiv(n, 256), 256, 0, st, h->asg.as<uint32_t>(), n, nb, h->tab.root_order_pos, t, ka, va);
        radix_sort_pairs(&ka, &va, &kb, &vb, n, 0, std::max(1, ceil_log2((uint64_t)G)), ws, st);
        if (va != h->sh_perm.as<uint32_t>()) SCB_CUDA(cudaMemcpyAsync(h->sh_perm.p, va, (size_t)n * 4, cudaMemcpyDeviceToDevice, st));
    }
    h->sh_split_mode = split ? 0 : 1;
    const uint32_t *perm = h->sh_perm.as<uint32_t>();
    DevBuf dfirst((size_t)(G + 1) * 8, st), dnb((size_t)(G + 1) * 8, st);
    SCB_LAUNCH(dest_bounds_k, 1, kMaxRanks + 1, 0, st, ka, n, G, dfirst.as<int64_t>());
    h->sh_aux.alloc((size_t)n1 * 8, st);
    h->sh_first.assign((size_t)G + 1, 0); h->sh_nbytes.assign((size_t)G + 1, 0);
    int64_t name_bytes = 0;
    if (!split) {
        // send order = input order: the names of a destination are one contiguous piece of the input's name array, nothing is
        // staged; and nothing here depends on the tie-break, so this may run (and the row sends may start) before it. The aux
        // words need bucket and end marker: packed now if the tie-break is done, else by the first send that carries them.
        if (cfg.use_names && n > 0) SCB_LAUNCH(gather_u64_k, 1, kMaxRanks + 1, 0, st, (const uint64_t *)c.name_off, dfirst.as<int64_t>(), G + 1, dnb.as<int64_t>());
        else SCB_CUDA(cudaMemsetAsync(dnb.p, 0, (size_t)(G + 1) * 8, st));
        SCB_CUDA(cudaMemcpyAsync(h->sh_first.data(), dfirst.p, (size_t)(G + 1) * 8, cudaMemcpyDeviceToHost, st));
        SCB_CUDA(cudaMemcpyAsync(h->sh_nbytes.data(), dnb.p, (size_t)(G + 1) * 8, cudaMemcpyDeviceToHost, st));
        SCB_CUDA(cudaStreamSynchronize(st));
        const int64_t nb0 = h->sh_nbytes[0];
        for (int g = 0; g <= G; g++) h->sh_nbytes[(size_t)g] -= nb0;
        h->sh_names_src = cfg.use_names ? c.names + nb0 : nullptr;
        h->sh_aux_pending = true;
        if (h->sh_resolved && n > 0) {
            SCB_LAUNCH(pack_aux_k, (unsigned)cdiv(n, 256), 256, 0, st, perm, n, h->asg.as<uint32_t>(), h->endv.as<uint16_t>(),
                       cfg.use_names ? c.name_off : (const int64_t *)nullptr, h->chunk.as<uint32_t>(), h->sh_aux.as<uint64_t>());
            h->sh_aux_pending = false;
        }
    } else {
        if (n > 0)
            SCB_LAUNCH(pack_aux_k, (unsigned)cdiv(n, 256), 256, 0, st, perm, n, h->asg.as<uint32_t>(), h->endv.as<uint16_t>(),
                       cfg.use_names ? c.name_off : (const int64_t *)nullptr, h->chunk.as<uint32_t>(), h->sh_aux.as<uint64_t>());
        h->sh_noff.alloc((size_t)(n + 1) * 8, st);
        DevBuf ws64((size_t)scan_tiles(n) * 8, st);
        exclusive_scan<uint64_t>(AuxNameLen{h->sh_aux.as<uint64_t>()}, n, h->sh_noff.as<uint64_t>(), h->sh_noff.as<uint64_t>() + n, ws64.as<uint64_t>(), st);
        SCB_LAUNCH(gather_u64_k, 1, kMaxRanks + 1, 0, st, h->sh_noff.as<uint64_t>(), dfirst.as<int64_t>(), G + 1, dnb.as<int64_t>());
        SCB_CUDA(cudaMemcpyAsync(h->sh_first.data(), dfirst.p, (size_t)(G + 1) * 8, cudaMemcpyDeviceToHost, st));
        SCB_CUDA(cudaMemcpyAsync(h->sh_nbytes.data(), dnb.p, (size_t)(G + 1) * 8, cudaMemcpyDeviceToHost, st));
        SCB_CUDA(cudaStreamSynchronize(st));
    }
    h->sh_cnt_reads.assign((size_t)G, 0); h->sh_cnt_name_bytes.assign((size_t)G, 0);
    for (int g = 0; g < G; g++) { h->sh_cnt_reads[g] = h->sh_first[g + 1] - h->sh_first[g]; h->sh_cnt_name_bytes[g] = h->sh_nbytes[g + 1] - h->sh_nbytes[g]; }
    name_bytes = h->sh_nbytes[(size_t)G];
    if (cfg.use_names && split) {   // names are staged contiguously per destination (variable length: bulk copies move them)
        h->sh_names.alloc((size_t)name_bytes + 16, st);
        if (n > 0) SCB_LAUNCH(pack_names_k, (unsigned)cdiv(n, 256), 256, 0, st, perm, n, c.name_off, c.names, h->sh_noff.as<uint64_t>(), h->sh_names.as<uint8_t>());
        h->sh_names_src = h->sh_names.as<uint8_t>();
    }
    tm.stop();
    memset(out, 0, sizeof *out);
    out->n = n; out->name_bytes = cfg.use_names ? name_bytes : 0;
    out->aux = h->sh_aux.as<uint64_t>();
    out->names = cfg.use_names ? h->sh_names_src : nullptr;
    out->cnt_reads = h->sh_cnt_reads.data(); out->cnt_name_bytes = h->sh_cnt_name_bytes.data();
    out->packed_row_bytes = h->PW * 4;
    h->sh_G = G;
    h->sh_phase = std::max(h->sh_phase, 3);
}

// staged variant: rows gathered into local send arrays (the caller moves them, e.g. NCCL all-to-all)
static void shard_stage(scb_handle *h, scb_shard_xfer *out) {
    cudaStream_t st = h->st;
    ArenaScope arena_scope(&h->arena);
    const scb_config &cfg = h->cfg;
    const int L1 = cfg.read_length[0], L2 = cfg.read_length[1];
    const Pending &c = h->cur;
    const int64_t n = c.n;
    const uint32_t *perm = h->sh_perm.as<uint32_t>();
    const int prow = h->PW * 4;
    ShardTimer tm(h);
    h->sh_packed.alloc((size_t)n * prow + 64, st);
    gather_rows_to(st, h->packed.as<uint8_t>(), h->sh_packed.as<uint8_t>(), perm, n, prow);
    if (cfg.use_quals) { h->sh_qual1.alloc((size_t)n * L1 + 16, st); gather_rows_to(st, c.qual1, h->sh_qual1.as<uint8_t>(), perm, n, L1); }
    if (cfg.paired) {
        h->sh_seq2.alloc((size_t)n * L2 + 16, st); gather_rows_to(st, c.seq2, h->sh_seq2.as<uint8_t>(), perm, n, L2);
        if (cfg.use_quals) { h->sh_qual2.alloc((size_t)n * L2 + 16, st); gather_rows_to(st, c.qual2, h->sh_qual2.as<uint8_t>(), perm, n, L2); }
    }
    tm.stop();
    out->packed = h->sh_packed.as<uint8_t>();
    out->qual1 = cfg.use_quals ? h->sh_qual1.as<uint8_t>() : nullptr;
    out->seq2 = cfg.paired ? h->sh_seq2.as<uint8_t>() : nullptr;
    out->qual2 = (cfg.paired && cfg.use_quals) ? h->sh_qual2.as<uint8_t>() : nullptr;
    h->sh_phase = 4;
}

// fused pack + send: every row gather writes straight into its owner's receive buffer (peer memory over
// NVLink, or this rank's own buffer). Destinations are visited in rotated order (rank+1, rank+2, ...) so that at
// any moment the ranks target different owners.
static void shard_send(scb_handle *h, int rank, const scb_shard_peer *peers, int what /* 1 small arrays, 2 rows, 3 both */, bool async) {
    const scb_config &cfg = h->cfg;
    const int L1 = cfg.read_length[0], L2 = cfg.read_length[1];
    const Pending &c = h->sh_phase >= 5 ? h->sh_local : h->cur;   // after import `cur` is the received set
    const int G = h->sh_G;
    const int prow = h->PW * 4;
    // async: on a side stream with a capped grid (NVLink-bound work needs few SMs), so that the receive side's sort
    // runs next to it; scb_shard_send_wait joins
    cudaStream_t st = async ? h->st_aux[0] : h->st;
    // CTAs of the overlapped row sends: 4 per SM measured best next to the sort (2 GPUs, 50M x 150: step 59.0 / 55.6 / 54.4 / 55.9 ms
    // at 1 / 2 / 4 per SM / uncapped; profiles/r02_multi_gpu_ab.txt)
    const int cap = async ? 148 * 4 : 0;
    if (async) { SCB_CUDA(cudaEventRecord(h->ev_fork, h->st)); SCB_CUDA(cudaStreamWaitEvent(st, h->ev_fork, 0)); }
    cudaEventRecord(async ? h->ev_a0 : h->ev_s0, st);
    if ((what & 1) && h->sh_aux_pending) {     // partitioned before the tie-break (chunk ownership): bucket and end marker exist only now
        if (!h->sh_resolved) throw CudaError{"sharded run: the aux words need scb_shard_finalize before they are sent"};
        const int64_t n = c.n;
        if (n > 0)
            SCB_LAUNCH(pack_aux_k, (unsigned)cdiv(n, 256), 256, 0, st, h->sh_perm.as<uint32_t>(), n, h->asg.as<uint32_t>(), h->endv.as<uint16_t>(),
                       cfg.use_names ? c.name_off : (const int64_t *)nullptr, h->chunk.as<uint32_t>(), h->sh_aux.as<uint64_t>());
        h->sh_aux_pending = false;
    }
    // chunk ownership: a destination's reads are one contiguous range of the input, every array moves as one copy (copy engines,
    // no SMs); bucket ranges: row gathers by the owner-sorted order
    const bool contiguous = h->sh_split_mode == 1;
    auto rows = [&](const uint8_t *src, void *dst_base, int64_t row_off, int64_t f, int64_t ng, int L) {
        if (contiguous) SCB_CUDA(cudaMemcpyAsync((uint8_t *)dst_base + row_off * L, src + f * (int64_t)L, (size_t)ng * L, cudaMemcpyDeviceToDevice, st));
        else gather_rows_to(st, src, (uint8_t *)dst_base + row_off * L, h->sh_perm.as<uint32_t>() + f, ng, L, cap);
    };
    for (int k = 1; k <= G; k++) {
        const int g = (rank + k) % G;
        const int64_t ng = h->sh_cnt_reads[g];
        if (ng == 0) continue;
        const scb_shard_peer &pp = peers[g];
        const int64_t f = h->sh_first[g];
        if (what & 1) {
            SCB_CUDA(cudaMemcpyAsync((uint8_t *)pp.aux + pp.row_off * 8, h->sh_aux.as<uint64_t>() + f, (size_t)ng * 8, cudaMemcpyDeviceToDevice, st));
            rows(h->packed.as<uint8_t>(), pp.packed, pp.row_off, f, ng, prow);
            if (cfg.use_names && h->sh_cnt_name_bytes[g] > 0)
                SCB_CUDA(cudaMemcpyAsync((uint8_t *)pp.names + pp.name_off, h->sh_names_src + h->sh_nbytes[g], (size_t)h->sh_cnt_name_bytes[g],
                                         cudaMemcpyDeviceToDevice, st));
        }
        if ((what & 2) && contiguous && g == rank) {
            // the rank's own rows stay where they are: the emit reads them from the input (RowSrc), so most of the "exchange" of
            // chunk ownership - a local copy of ~80-90 % of the quality / mate-2 bytes - does not happen at all
            h->sh_rows_in_place = true; h->sh_self_lo = pp.row_off; h->sh_rank = rank;
        } else if (what & 2) {
            if (cfg.use_quals) rows(c.qual1, pp.qual1, pp.row_off, f, ng, L1);
            if (cfg.paired) {
                rows(c.seq2, pp.seq2, pp.row_off, f, ng, L2);
                if (cfg.use_quals) rows(c.qual2, pp.qual2, pp.row_off, f, ng, L2);
            }
        }
    }
    SCB_CUDA(cudaEventRecord(async ? h->ev_a1 : h->ev_s1, st));
    if (!async) {
        SCB_CUDA(cudaStreamSynchronize(st));
        SCB_CUDA(cudaEventElapsedTime(&h->sh_ms, h->ev_s0, h->ev_s1));
    }
    if (h->sh_phase < 4) h->sh_phase = 4;
}
static void shard_send_wait(scb_handle *h) {
    SCB_CUDA(cudaStreamSynchronize(h->st_aux[0]));
    SCB_CUDA(cudaEventElapsedTime(&h->sh_ms, h->ev_a0, h->ev_a1));
    if (!h->sh_rows_in_place) h->sh_local = Pending();     // else the emit still reads the rank's own rows from it
}

// persistent receive buffers (plain cudaMalloc so that they can be exported over CUDA IPC), grown with headroom
static void shard_recv_reserve(scb_handle *h, const int64_t *need, void **ptrs, int32_t *changed) {
    *changed = 0;
    SCB_CUDA(cudaStreamSynchronize(h->st));
    for (int k = 0; k < 6; k++) {
        const size_t want = (size_t)std::max<int64_t>(need[k], 0) + 256;
        if (need[k] > 0 && h->rx_cap[k] < want) {
            // peers may still hold an IPC mapping of the old array: it is freed at the next import, after they
            // have re-mapped (CUDA leaves freeing an array that is still imported elsewhere undefined)
            if (h->rx[k]) h->rx_retired.push_back(h->rx[k]);
            h->rx[k] = nullptr; h->rx_cap[k] = 0;
            const size_t cap = want + want / 8;
            SCB_CUDA(cudaMalloc(&h->rx[k], cap));
            h->rx_cap[k] = cap;
            *changed = 1;
        }
        ptrs[k] = h->rx[k];
    }
}

static void shard_import(scb_handle *h, const scb_shard_xfer *in, int32_t n_chunks_global) {
    cudaStream_t st = h->st;
    ArenaScope arena_scope(&h->arena);
    const scb_config &cfg = h->cfg;
    const int L1 = cfg.read_length[0];
    const int nb = h->tab.n_buckets;
    const int64_t n = in->n;
    h->PW = (L1 + 15) / 16;
    if (in->packed_row_bytes != h->PW * 4) throw CudaError{"import: packed row size does not match the read length"};
    if (n >= (1ll << 31)) throw CudaError{"import: more than 2^31-1 reads on one rank"};
    SCB_CUDA(cudaStreamSynchronize(st));
    for (void *r : h->rx_retired) cudaFree(r);
    h->rx_retired.clear();
    h->arena.rewind(h->sh_mark);   // everything of the scan/resolve side is dead now (the send arrays were delivered)
    ShardTimer tm(h);
    Pending imp;
    imp.n = n; imp.borrowed = true;
    imp.qual1 = in->qual1; imp.names = in->names; imp.seq2 = in->seq2; imp.qual2 = in->qual2; imp.name_bytes = in->name_bytes;
    if (h->sh_rows_in_place) {
        // received rows [lo, hi) are this rank's own reads [first, first + cnt) of its input: row r there is input row r - lo + first
        const int L2 = cfg.read_length[1];
        const int64_t lo = h->sh_self_lo, cnt = h->sh_cnt_reads[(size_t)h->sh_rank], first = h->sh_first[(size_t)h->sh_rank];
        const Pending &own = h->cur;
        imp.own_lo = lo; imp.own_hi = lo + cnt;
        auto shifted = [&](const uint8_t *p, int L) { return p ? (const uint8_t *)((intptr_t)p + (intptr_t)(first - lo) * L) : nullptr; };
        imp.own_qual1 = shifted(own.qual1, L1); imp.own_seq2 = shifted(own.seq2, L2); imp.own_qual2 = shifted(own.qual2, L2);
    }
    h->packed.borrow((void *)in->packed, (size_t)n * h->PW * 4);
    h->asg.alloc((size_t)std::max<int64_t>(n, 1) * 4, st); h->endv.alloc((size_t)std::max<int64_t>(n, 1) * 2, st); h->lvl.alloc((size_t)std::max<int64_t>(n, 1), st);
    h->n_chunks = n_chunks_global;
    if (n_chunks_global > 1) h->chunk.alloc((size_t)std::max<int64_t>(n, 1) * 4, st); else h->chunk.release();
    if (n > 0)
        SCB_LAUNCH(unpack_aux_k, (unsigned)cdiv(n, 256), 256, 0, st, in->aux, n, nb, h->d_rank_level.as<uint8_t>(), h->asg.as<uint32_t>(),
                   h->endv.as<uint16_t>(), h->lvl.as<uint8_t>(), n_chunks_global > 1 ? h->chunk.as<uint32_t>() : (uint32_t *)nullptr);
    if (cfg.use_names) {
        h->sh_name_off.alloc((size_t)(n + 1) * 8, st);
        DevBuf ws64((size_t)scan_tiles(n) * 8, st);
        exclusive_scan<uint64_t>(AuxNameLen{in->aux}, n, h->sh_name_off.as<uint64_t>(), h->sh_name_off.as<uint64_t>() + n, ws64.as<uint64_t>(), st);
        imp.name_off = h->sh_name_off.as<int64_t>();
    }
    h->sh_local = std::move(h->cur);   // the rank's own input: overlapping row sends still read it
    h->cur = std::move(imp);
    h->n_perm = n;
    tm.stop();
    h->sh_phase = 5;
}

static void shard_finish(scb_handle *h, int what /* 1 = sort, 2 = emit, 3 = both; 4 / 8 = the emit in two parts (emit_order) */) {
    ArenaScope arena_scope(&h->arena);
    ShardTimer tm(h);
    if (what & 1) { stage_meta(h); stage_sort(h); h->srt_keys_valid = true; }
    if (what & 2) stage_emit(h);
    if (what & 4) stage_emit(h, 1);
    if (what & 8) stage_emit(h, 2);
    tm.stop();
    if (what & (2 | 8)) {
        h->sh_phase = 6;   // stage_debug keys on != 0; reset by the next flush / scan
        if (h->sh_rows_in_place) { SCB_CUDA(cudaStreamSynchronize(h->st)); h->sh_local = Pending(); h->sh_rows_in_place = false; }
    }
}

// ---- the transform on one GPU ------------------------------------------------------------------------------
// closed_only (scb_flush_closed): only the flush chunks that are complete are emitted; the reads of the open chunk are taken out
// of the populations again and stay pending, so that the next flush decides them exactly as if nothing had happened - the chunk
// boundaries, chunk contents and lifetime counts of a job fed in pieces are those of the reference fed the whole input
// (compress.cpp:702-713 carries total_size across reads and input files).
static void run_flush(scb_handle *h, bool closed_only = false) {
    cudaStream_t st = h->st;
    Pending tail;                 // reads of the open chunk, copied out of the flush workspace (pooled memory)
    {
    ArenaScope arena_scope(&h->arena);
    h->sh_phase = 0;
    flush_begin(h, 1.0);
    h->n_perm = h->cur.n;
    SCB_CUDA(cudaEventRecord(h->ev0, st));
    SCB_CUDA(cudaEventRecord(h->stage_ev[0], st));
    stage_scan(h);
    SCB_CUDA(cudaEventRecord(h->stage_ev[1], st));
    stage_resolve(h);
    stage_meta(h);
    SCB_CUDA(cudaEventRecord(h->stage_ev[2], st));
    stage_chunks(h);
    const int64_t n_all = h->cur.n;
    int64_t keep = n_all;
    if (closed_only && h->open_from < n_all) {
        keep = h->open_from;
        const int64_t m = n_all - keep;
        const int nb = h->tab.n_buckets;
        DevBuf droots(8, st);
        SCB_CUDA(cudaMemsetAsync(droots.p, 0, 8, st));
        SCB_LAUNCH(uncommit_tail_k, (unsigned)cdiv(m, 256), 256, 0, st, h->asg.as<uint32_t>(), keep, n_all, h->d_life.as<unsigned long long>(), nb, droots.as<unsigned long long>());
        unsigned long long roots = 0;
        SCB_CUDA(cudaMemcpyAsync(&roots, droots.p, 8, cudaMemcpyDeviceToHost, st));
        SCB_CUDA(cudaStreamSynchronize(st));
        h->life_total -= (uint64_t)m;
        if (h->tab.root_counts_unbucketed) h->unbucketed -= (int64_t)roots;
        h->n_chunks -= 1;
        h->cur.n = keep;              // everything below works on the closed chunks only
        h->n_perm = keep;
        h->n_last = keep;
    }
    SCB_CUDA(cudaEventRecord(h->stage_ev[3], st));
    stage_sort(h);
    stage_emit(h);
    SCB_CUDA(cudaEventRecord(h->stage_ev[7], st));
    stage_debug(h);
    SCB_CUDA(cudaEventRecord(h->stage_ev[8], st));
    SCB_CUDA(cudaEventRecord(h->ev1, st));
    SCB_CUDA(cudaStreamSynchronize(st));
    for (int k = 0; k < SCB_N_STAGES; k++) SCB_CUDA(cudaEventElapsedTime(&h->stage_ms[k], h->stage_ev[k], h->stage_ev[k + 1]));
    if (keep < n_all) {
        // the open chunk's reads leave the flush workspace (the next flush re-uses it from the start) as an owned pending batch
        const scb_config &cfg = h->cfg;
        const int L1 = cfg.read_length[0], L2 = cfg.read_length[1];
        const Pending &c = h->cur;
        const int64_t m = n_all - keep;
        g_arena = nullptr;            // pooled allocations from here on
        tail.n = m;
        auto cp = [&](DevBuf &d, const uint8_t *src, size_t bytes) { d.alloc(bytes, st); SCB_CUDA(cudaMemcpyAsync(d.p, src, bytes, cudaMemcpyDeviceToDevice, st)); };
        cp(tail.b_seq1, c.seq1 + keep * L1, (size_t)m * L1);
        if (cfg.use_quals) cp(tail.b_qual1, c.qual1 + keep * L1, (size_t)m * L1);
        if (cfg.paired) { cp(tail.b_seq2, c.seq2 + keep * L2, (size_t)m * L2); if (cfg.use_quals) cp(tail.b_qual2, c.qual2 + keep * L2, (size_t)m * L2); }
        if (cfg.use_names) {
            int64_t o2[2] = {0, 0};
            SCB_CUDA(cudaMemcpyAsync(&o2[0], c.name_off + keep, 8, cudaMemcpyDeviceToHost, st));
            SCB_CUDA(cudaMemcpyAsync(&o2[1], c.name_off + n_all, 8, cudaMemcpyDeviceToHost, st));
            SCB_CUDA(cudaStreamSynchronize(st));
            tail.name_bytes = o2[1] - o2[0];
            cp(tail.b_names, c.names + o2[0], (size_t)tail.name_bytes);
            tail.b_off.alloc((size_t)(m + 1) * 8, st);
            SCB_LAUNCH(rebase_off_k, (unsigned)cdiv(m + 1, 256), 256, 0, st, c.name_off + keep, m + 1, tail.b_off.as<int64_t>());
        }
        SCB_CUDA(cudaStreamSynchronize(st));
        tail.seq1 = tail.b_seq1.as<uint8_t>(); tail.qual1 = tail.b_qual1.as<uint8_t>(); tail.seq2 = tail.b_seq2.as<uint8_t>(); tail.qual2 = tail.b_qual2.as<uint8_t>();
        tail.names = tail.b_names.as<uint8_t>(); tail.name_off = tail.b_off.as<int64_t>();
    }
    }   // arena scope
    if (tail.n > 0) h->pending.insert(h->pending.begin(), std::move(tail));
}

static void fill_result(scb_handle *h, scb_result *out) {
    out->n_reads = h->cur.n;
    out->n_chunks = h->n_chunks;
    out->n_buckets_nonempty = (int32_t)(h->cfg.emit_merged && h->n_chunks > 1 ? h->merged.n_seg : (h->n_chunks <= 1 ? h->chunked.n_seg : 0));
    const bool alias = !(h->cfg.emit_merged && h->n_chunks > 1);
    for (int k = 0; k < SCB_N_STREAMS; k++) {
        out->data[k] = h->chunked.data[k].as<uint8_t>();
        out->chunk_off[k] = h->chunked.chunk_off[k].data();
        if (h->cfg.emit_merged) {
            const EmitOut &m = alias ? h->chunked : h->merged;
            out->merged[k] = m.data[k].as<uint8_t>();
            out->merged_size[k] = m.size[k];
        }
    }
    out->bucket_id = h->dbg_bucket.as<int32_t>(); out->core_idx = h->dbg_core.as<int32_t>();
    out->end = h->dbg_end.as<int32_t>(); out->chunk = h->dbg_chunk.as<int32_t>();
    out->perm = h->perm.as<uint32_t>();
}

// ---- (f2) FASTQ text -> pending batch (parse.cuh) ----------------------------------------------------------------------------
struct ParsedMate { DevBuf text, tile_cnt, tile_first, line_end, name_len, name_off, err; int64_t n = 0; const uint8_t *t = nullptr; int64_t bytes = 0; };

static int64_t parse_lines(scb_handle *h, ParsedMate &m, const uint8_t *text, int64_t bytes, int location) {
    cudaStream_t st = h->st;
    m.bytes = bytes;
    if (location == 0) {
        m.text.alloc((size_t)bytes + 16, st);
        SCB_CUDA(cudaMemcpyAsync(m.text.p, text, (size_t)bytes, cudaMemcpyHostToDevice, st));
        m.t = m.text.as<uint8_t>();
    } else m.t = text;
    if (bytes == 0) { m.n = 0; return 0; }
    const int64_t tiles = cdiv(bytes, kParseTile);
    m.tile_cnt.alloc((size_t)tiles * 4, st); m.tile_first.alloc((size_t)(tiles + 1) * 8, st);
    SCB_LAUNCH(parse_count_nl_k, (unsigned)tiles, 256, 0, st, m.t, bytes, m.tile_cnt.as<uint32_t>());
    DevBuf ws64((size_t)scan_tiles(tiles) * 8, st);
    exclusive_scan<uint64_t>(NlCount{m.tile_cnt.as<uint32_t>()}, tiles, m.tile_first.as<uint64_t>(), m.tile_first.as<uint64_t>() + tiles, ws64.as<uint64_t>(), st);
    uint64_t nl = 0; uint8_t last = 0;
    SCB_CUDA(cudaMemcpyAsync(&nl, m.tile_first.as<uint64_t>() + tiles, 8, cudaMemcpyDeviceToHost, st));
    SCB_CUDA(cudaMemcpyAsync(&last, m.t + bytes - 1, 1, cudaMemcpyDeviceToHost, st));
    SCB_CUDA(cudaStreamSynchronize(st));
    const int64_t lines = (int64_t)nl + (last != '\n' ? 1 : 0);         // a last line without its newline still ends at the end of the text
    if (lines % 4 != 0) throw CudaError{"FASTQ text does not hold a whole number of 4-line records"};
    m.line_end.alloc((size_t)std::max<int64_t>(lines, 1) * 8, st);
    SCB_LAUNCH(parse_line_ends_k, (unsigned)tiles, 256, 0, st, m.t, bytes, m.tile_first.as<uint64_t>(), m.line_end.as<int64_t>(), lines);
    if (last != '\n') SCB_CUDA(cudaMemcpyAsync(m.line_end.as<int64_t>() + lines - 1, &bytes, 8, cudaMemcpyHostToDevice, st));
    m.n = lines / 4;
    return m.n;
}

static void submit_fastq(scb_handle *h, const uint8_t *text1, int64_t bytes1, const uint8_t *text2, int64_t bytes2, int location, const int32_t *phred, int64_t *n_out) {
    cudaStream_t st = h->st;
    const scb_config &cfg = h->cfg;
    const int L[2] = {cfg.read_length[0], cfg.read_length[1]};
    ParsedMate pm[2];
    const int mates = cfg.paired ? 2 : 1;
    for (int m = 0; m < mates; m++) parse_lines(h, pm[m], m ? text2 : text1, m ? bytes2 : bytes1, location);
    const int64_t n = pm[0].n;
    if (cfg.paired && pm[1].n != n) throw CudaError{"the two FASTQ texts hold different numbers of records"};
    *n_out = n;
    if (n == 0) return;
    int64_t total = n;
    for (auto &p : h->pending) total += p.n;
    if (total >= (1ll << 31)) throw CudaError{"more than 2^31-1 reads pending; flush first"};
    Pending p;
    p.n = n;
    DevBuf derr(4, st);
    SCB_CUDA(cudaMemsetAsync(derr.p, 0, 4, st));
    for (int m = 0; m < mates; m++) {
        pm[m].name_len.alloc((size_t)n * 4, st);
        ParseRec pr{pm[m].t, pm[m].bytes, pm[m].line_end.as<int64_t>(), n, L[m], pm[m].name_len.as<uint32_t>(), derr.as<uint32_t>()};
        SCB_LAUNCH(parse_records_k, (unsigned)cdiv(n, 256), 256, 0, st, pr);
    }
    if (cfg.use_names) {
        p.b_off.alloc((size_t)(n + 1) * 8, st);
        DevBuf ws64((size_t)scan_tiles(n) * 8, st);
        exclusive_scan<uint64_t>(NameLenU{pm[0].name_len.as<uint32_t>()}, n, p.b_off.as<uint64_t>(), p.b_off.as<uint64_t>() + n, ws64.as<uint64_t>(), st);
        uint64_t nb = 0;
        SCB_CUDA(cudaMemcpyAsync(&nb, p.b_off.as<uint64_t>() + n, 8, cudaMemcpyDeviceToHost, st));
        SCB_CUDA(cudaStreamSynchronize(st));
        p.name_bytes = (int64_t)nb;
        p.b_names.alloc((size_t)nb + 16, st);
    }
    p.b_seq1.alloc((size_t)n * L[0], st);
    if (cfg.use_quals) p.b_qual1.alloc((size_t)n * L[0], st);
    if (cfg.paired) { p.b_seq2.alloc((size_t)n * L[1], st); if (cfg.use_quals) p.b_qual2.alloc((size_t)n * L[1], st); }
    for (int m = 0; m < mates; m++) {
        ParseCopy pc;
        pc.text = pm[m].t; pc.line_end = pm[m].line_end.as<int64_t>(); pc.n = n; pc.L = L[m]; pc.phred = phred[m];
        pc.name_off = (m == 0 && cfg.use_names) ? p.b_off.as<uint64_t>() : nullptr;
        pc.seq = m ? p.b_seq2.as<uint8_t>() : p.b_seq1.as<uint8_t>();
        pc.qual = cfg.use_quals ? (m ? p.b_qual2.as<uint8_t>() : p.b_qual1.as<uint8_t>()) : nullptr;
        pc.names = (m == 0 && cfg.use_names) ? p.b_names.as<uint8_t>() : nullptr;
        pc.err = derr.as<uint32_t>();
        SCB_LAUNCH(parse_copy_k, (unsigned)cdiv(n * 32, 256), 256, 0, st, pc);
    }
    uint32_t err = 0;
    SCB_CUDA(cudaMemcpyAsync(&err, derr.p, 4, cudaMemcpyDeviceToHost, st));
    SCB_CUDA(cudaStreamSynchronize(st));
    if (err) {
        throw CudaError{std::string("FASTQ text rejected:") + ((err & 1) ? " a name line does not start with '@';" : "") + ((err & 2) ? " a read line does not have the configured length;" : "") +
                        ((err & 4) ? " a quality line does not have the configured length;" : "") + ((err & 8) ? " a name is longer than 255 bytes;" : "") +
                        ((err & 16) ? " a quality symbol lies outside [offset, offset + 80);" : "")};
    }
    if (cfg.use_quals) {   // output_quality's input-order statistics (qualities.cpp:186-199), per mate
        for (int m = 0; m < mates; m++) {
            if (!h->q_freq3[m].p) {
                h->q_freq3[m].alloc((size_t)kAcDepth * kAcDepth * 8, st); h->q_freq4[m].alloc((size_t)kAcDepth * kAcDepth * kAcDepth * 8, st);
                SCB_CUDA(cudaMemsetAsync(h->q_freq3[m].p, 0, h->q_freq3[m].bytes, st));
                SCB_CUDA(cudaMemsetAsync(h->q_freq4[m].p, 0, h->q_freq4[m].bytes, st));
            }
            const uint8_t *sym = m ? p.b_qual2.as<uint8_t>() : p.b_qual1.as<uint8_t>();
            const int64_t tot = n * L[m];
            int dev_sms = kSMs;
            SCB_CUDA(cudaDeviceGetAttribute(&dev_sms, cudaDevAttrMultiProcessorCount, cfg.device));
            const int grid = (int)std::min<int64_t>((int64_t)dev_sms * 4, cdiv(cdiv(tot, kStatRun), 256));
            SCB_LAUNCH(parse_qstats_k, grid, 256, 0, st, sym, tot, h->q_prev[m][0], h->q_prev[m][1], h->q_freq3[m].as<unsigned long long>(), h->q_freq4[m].as<unsigned long long>());
            uint8_t lastq[2] = {0, 0};
            if (tot >= 2) SCB_CUDA(cudaMemcpyAsync(lastq, sym + tot - 2, 2, cudaMemcpyDeviceToHost, st));
            else SCB_CUDA(cudaMemcpyAsync(lastq + 1, sym + tot - 1, 1, cudaMemcpyDeviceToHost, st));
            SCB_CUDA(cudaStreamSynchronize(st));
            if (tot >= 2) { h->q_prev[m][0] = lastq[0]; h->q_prev[m][1] = lastq[1]; }
            else { h->q_prev[m][0] = h->q_prev[m][1]; h->q_prev[m][1] = lastq[1]; }
            h->q_seen[m] += (uint64_t)tot;
        }
    }
    p.seq1 = p.b_seq1.as<uint8_t>(); p.qual1 = p.b_qual1.as<uint8_t>(); p.seq2 = p.b_seq2.as<uint8_t>(); p.qual2 = p.b_qual2.as<uint8_t>();
    p.names = p.b_names.as<uint8_t>(); p.name_off = p.b_off.as<int64_t>();
    SCB_CUDA(cudaStreamSynchronize(st));
    h->pending.push_back(std::move(p));
}

// ---- (f3) .scalcer body from merged meta + stream 1; (f4) the inverse (container.cuh) -----------------------------------
static void segment_table(scb_handle *h, const uint8_t *meta_dev, int64_t nseg, DevBuf &seg_core, DevBuf &seg_reads, DevBuf &seg_bytes, DevBuf &in_start) {
    cudaStream_t st = h->st;
    const scb_config &cfg = h->cfg;
    const int L1 = cfg.read_length[0];
    const int rsz = 8 + 8 * (3 + 2 * cfg.paired);
    const int64_t n1 = std::max<int64_t>(nseg, 1);
    seg_core.alloc((size_t)n1 * 4, st); seg_reads.alloc((size_t)n1 * 8, st); seg_bytes.alloc((size_t)n1 * 8, st); in_start.alloc((size_t)(n1 + 1) * 8, st);
    if (nseg > 0)
        SCB_LAUNCH(ct_segments_k, (unsigned)cdiv(nseg, 256), 256, 0, st, MetaRecs{meta_dev, rsz}, nseg, h->d_core_len.as<int32_t>(), L1, L1 > 255 ? 2 : 1,
                   seg_core.as<int32_t>(), seg_reads.as<int64_t>(), seg_bytes.as<int64_t>());
    DevBuf ws64((size_t)scan_tiles(nseg) * 8, st);
    exclusive_scan<uint64_t>(SegBytes{seg_bytes.as<int64_t>()}, nseg, in_start.as<uint64_t>(), in_start.as<uint64_t>() + nseg, ws64.as<uint64_t>(), st);
}

static void assemble_reads(scb_handle *h, int chunk) {
    cudaStream_t st = h->st;
    const int rsz = 8 + 8 * (3 + 2 * h->cfg.paired);
    const uint8_t *s1, *meta; int64_t b1, bm;
    if (chunk < 0) {
        const EmitOut &m = (h->n_chunks > 1) ? h->merged : h->chunked;
        s1 = m.data[1].as<uint8_t>(); b1 = m.size[1]; meta = m.data[3].as<uint8_t>(); bm = m.size[3];
    } else {
        const auto &c1 = h->chunked.chunk_off[1], &c3 = h->chunked.chunk_off[3];
        s1 = h->chunked.data[1].as<uint8_t>() + c1[(size_t)chunk]; b1 = c1[(size_t)chunk + 1] - c1[(size_t)chunk];
        meta = h->chunked.data[3].as<uint8_t>() + c3[(size_t)chunk]; bm = c3[(size_t)chunk + 1] - c3[(size_t)chunk];
    }
    const int64_t nseg = bm / rsz;
    segment_table(h, meta, nseg, h->as_seg_core, h->as_seg_reads, h->as_seg_bytes, h->as_in_start);
    const int64_t body = b1 + 12 * nseg;
    h->as_body.alloc((size_t)body + 16, st);
    if (body > 0)
        SCB_LAUNCH(ct_assemble_k, (unsigned)cdiv(cdiv(body, 16), 256), 256, 0, st, s1, h->as_in_start.as<uint64_t>(), nseg, h->as_seg_core.as<int32_t>(),
                   h->as_seg_reads.as<int64_t>(), h->as_body.as<uint8_t>(), body);
    SCB_CUDA(cudaStreamSynchronize(st));
    h->as_bytes = body; h->as_nseg = nseg;
}

static void inverse_reads(scb_handle *h, const uint8_t *stream, const int32_t *seg_core, const int64_t *seg_reads, int64_t nseg, const uint8_t *quals,
                          int mate, int phred, int location, uint8_t *seq_out, uint8_t *qual_out, int64_t *n_out) {
    cudaStream_t st = h->st;
    const scb_config &cfg = h->cfg;
    const int L = cfg.read_length[mate ? 1 : 0];
    const int sz_meta = cfg.read_length[0] > 255 ? 2 : 1;
    const int64_t n1 = std::max<int64_t>(nseg, 1);
    // the segment table is small: always staged through the host (it is what the container's inline headers hold)
    std::vector<int32_t> hc((size_t)n1, 0); std::vector<int64_t> hr((size_t)n1, 0);
    if (nseg > 0) {
        if (location == 1) {
            SCB_CUDA(cudaMemcpyAsync(hc.data(), seg_core, (size_t)nseg * 4, cudaMemcpyDeviceToHost, st));
            SCB_CUDA(cudaMemcpyAsync(hr.data(), seg_reads, (size_t)nseg * 8, cudaMemcpyDeviceToHost, st));
            SCB_CUDA(cudaStreamSynchronize(st));
        } else { memcpy(hc.data(), seg_core, (size_t)nseg * 4); memcpy(hr.data(), seg_reads, (size_t)nseg * 8); }
    }
    std::vector<uint64_t> in_start((size_t)n1 + 1, 0), read_start((size_t)n1 + 1, 0);
    const int32_t ncores = (int32_t)h->tab.cores.size();
    for (int64_t s2 = 0; s2 < nseg; s2++) {
        const int32_t c = hc[(size_t)s2];
        if (c != SCB_ROOT_ID && (c < 0 || c >= ncores)) throw CudaError{"inverse: core index out of range in the segment table"};
        if (hr[(size_t)s2] < 0) throw CudaError{"inverse: negative read count in the segment table"};
        const int lv = c == SCB_ROOT_ID ? 0 : (int)h->tab.cores[(size_t)c].size();
        read_start[(size_t)s2 + 1] = read_start[(size_t)s2] + (uint64_t)hr[(size_t)s2];
        in_start[(size_t)s2 + 1] = in_start[(size_t)s2] + (uint64_t)hr[(size_t)s2] * (uint64_t)(sz_read(cfg.read_length[0] - lv) + sz_meta);
    }
    const int64_t n = (int64_t)read_start[(size_t)nseg];
    *n_out = n;
    if (n == 0) return;
    const int64_t sbytes = mate ? n * sz_read(L) : (int64_t)in_start[(size_t)nseg];
    DevBuf d_stream, d_quals, d_seq, d_qual, d_core, d_is, d_rs;
    const uint8_t *ps = stream, *pq = quals;
    uint8_t *po = seq_out, *pqo = qual_out;
    if (location == 0) {
        d_stream.alloc((size_t)sbytes + 16, st);
        SCB_CUDA(cudaMemcpyAsync(d_stream.p, stream, (size_t)sbytes, cudaMemcpyHostToDevice, st));
        ps = d_stream.as<uint8_t>();
        if (quals) { d_quals.alloc((size_t)n * L, st); SCB_CUDA(cudaMemcpyAsync(d_quals.p, quals, (size_t)n * L, cudaMemcpyHostToDevice, st)); pq = d_quals.as<uint8_t>(); }
        d_seq.alloc((size_t)n * L, st); po = d_seq.as<uint8_t>();
        if (quals && qual_out) { d_qual.alloc((size_t)n * L, st); pqo = d_qual.as<uint8_t>(); }
    }
    if (!quals) pqo = nullptr;
    if (mate) {
        SCB_LAUNCH(ct_inverse2_k, (unsigned)cdiv(n * ((L + 15) / 16), 256), 256, 0, st, ps, n, L, pq, phred, po, pqo);
    } else {
        upload(d_core, hc, st); upload(d_is, in_start, st); upload(d_rs, read_start, st);
        InvParams ip;
        ip.stream1 = ps; ip.in_start = d_is.as<uint64_t>(); ip.read_start = d_rs.as<uint64_t>(); ip.seg_core = d_core.as<int32_t>();
        ip.core_len = h->d_core_len.as<int32_t>(); ip.core_off = h->d_core_off.as<uint64_t>(); ip.core_chars = h->d_core_chars.as<uint8_t>();
        ip.nseg = nseg; ip.n = n; ip.L = L; ip.sz_meta = sz_meta; ip.quals = pq; ip.phred = phred; ip.seq_out = po; ip.qual_out = pqo;
        SCB_LAUNCH(ct_inverse_k, (unsigned)cdiv(n * ((L + 15) / 16), 256), 256, 0, st, ip);
    }
    if (location == 0) {
        SCB_CUDA(cudaMemcpyAsync(seq_out, po, (size_t)n * L, cudaMemcpyDeviceToHost, st));
        if (pqo) SCB_CUDA(cudaMemcpyAsync(qual_out, pqo, (size_t)n * L, cudaMemcpyDeviceToHost, st));
    }
    SCB_CUDA(cudaStreamSynchronize(st));
}

}  // namespace scb

// =====================================================================================================
// C ABI
// =====================================================================================================
#define SCB_TRY try {
#define SCB_CATCH                                                          \
    }                                                                      \
    catch (scb::CudaError & e) { scb::g_last_error = e.msg; return SCB_ECUDA; } \
    catch (std::bad_alloc &) { scb::g_last_error = "host allocation failed"; return SCB_ENOMEM; }

extern "C" {

int scb_abi_version(void) { return SCB_ABI_VERSION; }
const char *scb_last_error(void) { return scb::g_last_error.c_str(); }

int scb_create(const char *const *cores, int32_t n_cores, const scb_config *cfg, scb_handle **out) {
    if (n_cores < 0 || (n_cores > 0 && !cores)) { scb::g_last_error = "bad core array"; return SCB_EINVAL; }
    std::vector<std::string> v;
    v.reserve((size_t)n_cores);
    for (int32_t i = 0; i < n_cores; i++) {
        if (!cores[i]) { scb::g_last_error = "null core string"; return SCB_EINVAL; }
        v.emplace_back(cores[i]);
    }
    return scb::create_common(v, cfg, out);
}

int scb_create_from_file(const char *path, const scb_config *cfg, scb_handle **out) {
    if (!path) { scb::g_last_error = "null path"; return SCB_EINVAL; }
    std::vector<std::string> v;
    std::string err = scb::load_core_file(path, v);
    if (!err.empty()) { scb::g_last_error = err; return SCB_EIO; }
    return scb::create_common(v, cfg, out);
}

int scb_table_dryrun(const char *const *cores, int32_t n_cores, int32_t *n_states, int32_t *n_buckets, int32_t *root_order_pos,
                     int32_t *core_node_id /* [n_cores] or NULL */) {
    if (n_cores < 0 || (n_cores > 0 && !cores)) { scb::g_last_error = "bad core array"; return SCB_EINVAL; }
    std::vector<std::string> v;
    for (int32_t i = 0; i < n_cores; i++) v.emplace_back(cores[i] ? cores[i] : "");
    scb::CoreTable t;
    std::string err = scb::build_core_table(v, t);
    if (!err.empty()) { scb::g_last_error = err; return SCB_EINVAL; }
    if (n_states) *n_states = t.n_states;
    if (n_buckets) *n_buckets = t.n_buckets;
    if (root_order_pos) *root_order_pos = t.root_order_pos;
    if (core_node_id)
        for (int32_t i = 0; i < n_cores; i++) core_node_id[i] = t.rank_node_id[(size_t)t.core_to_rank[(size_t)i]];
    return SCB_OK;
}

int scb_table_info(const scb_handle *h, int32_t *n_cores, int32_t *n_states, int32_t *n_buckets, int32_t *smem_resident) {
    if (!h) { scb::g_last_error = "null handle"; return SCB_EINVAL; }
    if (n_cores) *n_cores = (int32_t)h->tab.cores.size();
    if (n_states) *n_states = h->tab.n_states;
    if (n_buckets) *n_buckets = h->tab.n_buckets;
    if (smem_resident) *smem_resident = h->smem_resident ? 1 : 0;
    return SCB_OK;
}

const char *scb_core(const scb_handle *h, int32_t idx) {
    if (!h || idx < 0 || idx >= (int32_t)h->tab.cores.size()) return nullptr;
    return h->tab.cores[(size_t)idx].c_str();
}

int scb_submit(scb_handle *h, const scb_batch *b) {
    if (!h || !b) { scb::g_last_error = "null argument"; return SCB_EINVAL; }
    const scb_config &cfg = h->cfg;
    if (b->n < 0 || (!b->seq1 && b->n > 0)) { scb::g_last_error = "bad batch"; return SCB_EINVAL; }
    if (b->n == 0) return SCB_OK;
    if ((cfg.use_quals && !b->qual1) || (cfg.use_names && (!b->names || !b->name_off)) ||
        (cfg.paired && (!b->seq2 || (cfg.use_quals && !b->qual2)))) { scb::g_last_error = "batch is missing an array the configuration requires"; return SCB_EINVAL; }
    int64_t total = b->n;
    for (auto &p : h->pending) total += p.n;
    if (total >= (1ll << 31)) { scb::g_last_error = "more than 2^31-1 reads pending; flush first"; return SCB_EINVAL; }
    SCB_TRY
    SCB_CUDA(cudaSetDevice(cfg.device));
    cudaStream_t st = h->st;
    const int L1 = cfg.read_length[0], L2 = cfg.read_length[1];
    scb::Pending p;
    p.n = b->n;
    if (b->location == 1) {
        p.borrowed = true;
        p.seq1 = b->seq1; p.qual1 = b->qual1; p.seq2 = b->seq2; p.qual2 = b->qual2; p.names = b->names; p.name_off = b->name_off;
        if (cfg.use_names) {
            int64_t o[2];
            SCB_CUDA(cudaMemcpyAsync(&o[0], b->name_off, 8, cudaMemcpyDeviceToHost, st));
            SCB_CUDA(cudaMemcpyAsync(&o[1], b->name_off + b->n, 8, cudaMemcpyDeviceToHost, st));
            SCB_CUDA(cudaStreamSynchronize(st));
            if (o[0] != 0) { scb::g_last_error = "device batches must have name_off[0] == 0"; return SCB_EINVAL; }
            p.name_bytes = o[1];
            // the same contract as for host batches: every name 0..255 bytes (the reference's length byte, names.cpp:48-62)
            scb::DevBuf rng(16, st);
            SCB_CUDA(cudaMemsetAsync(rng.p, 0, 16, st));
            SCB_LAUNCH(scb::name_len_range_k, (unsigned)std::max<int64_t>(1, std::min<int64_t>(scb::cdiv(b->n, 256), 148 * 8)), 256, 0, st, b->name_off, b->n, rng.as<long long>());
            long long r2[2] = {0, 0};
            SCB_CUDA(cudaMemcpyAsync(r2, rng.p, 16, cudaMemcpyDeviceToHost, st));
            SCB_CUDA(cudaStreamSynchronize(st));
            if (r2[0] > 255 || r2[1] > 0) { scb::g_last_error = "name length outside 0..255"; return SCB_EINVAL; }
        }
    } else {
        auto up = [&](scb::DevBuf &d, const void *src, size_t bytes) {
            d.alloc(bytes, st);
            SCB_CUDA(cudaMemcpyAsync(d.p, src, bytes, cudaMemcpyHostToDevice, st));
        };
        up(p.b_seq1, b->seq1, (size_t)b->n * L1);
        if (cfg.use_quals) up(p.b_qual1, b->qual1, (size_t)b->n * L1);
        if (cfg.paired) { up(p.b_seq2, b->seq2, (size_t)b->n * L2); if (cfg.use_quals) up(p.b_qual2, b->qual2, (size_t)b->n * L2); }
        if (cfg.use_names) {
            int64_t o0 = b->name_off[0];
            p.name_bytes = b->name_off[b->n] - o0;
            // host-side validation of the offsets runs while the sequence / quality copies above are in flight (pinned sources)
            bool bad = p.name_bytes < 0;
            for (int64_t i = 0; i < b->n && !bad; i++) {
                int64_t len = b->name_off[i + 1] - b->name_off[i];
                bad = len < 0 || len > 255;
            }
            if (bad) {
                SCB_CUDA(cudaStreamSynchronize(st));   // the copies read the caller's buffers: finish them before returning
                scb::g_last_error = "name length outside 0..255";
                return SCB_EINVAL;
            }
            up(p.b_names, b->names + o0, (size_t)p.name_bytes);
            if (o0 == 0) up(p.b_off, b->name_off, (size_t)(b->n + 1) * 8);
            else {
                std::vector<int64_t> tmp((size_t)b->n + 1);
                for (int64_t i = 0; i <= b->n; i++) tmp[(size_t)i] = b->name_off[i] - o0;
                up(p.b_off, tmp.data(), tmp.size() * 8);
                SCB_CUDA(cudaStreamSynchronize(st));
            }
        }
        p.seq1 = p.b_seq1.as<uint8_t>(); p.qual1 = p.b_qual1.as<uint8_t>(); p.seq2 = p.b_seq2.as<uint8_t>(); p.qual2 = p.b_qual2.as<uint8_t>();
        p.names = p.b_names.as<uint8_t>(); p.name_off = p.b_off.as<int64_t>();
        // pageable sources must not be reused by the caller before the copy is done
        SCB_CUDA(cudaStreamSynchronize(st));
    }
    h->pending.push_back(std::move(p));
    SCB_CATCH
    return SCB_OK;
}

int scb_flush(scb_handle *h, scb_result *out) {
    if (!h || !out) { scb::g_last_error = "null argument"; return SCB_EINVAL; }
    memset(out, 0, sizeof *out);
    if (h->pending.empty()) h->pending.emplace_back();
    SCB_TRY
    SCB_CUDA(cudaSetDevice(h->cfg.device));
    scb::run_flush(h);
    float ms = 0;
    SCB_CUDA(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    out->device_ms = ms;
    SCB_CATCH
    scb::fill_result(h, out);
    return SCB_OK;
}

int scb_flush_closed(scb_handle *h, scb_result *out) {
    if (!h || !out) { scb::g_last_error = "null argument"; return SCB_EINVAL; }
    memset(out, 0, sizeof *out);
    if (h->cfg.emit_merged) { scb::g_last_error = "scb_flush_closed emits flush chunks only (emit_merged = 0): the merge across flushes is merge()'s (compress.cpp:488-522)"; return SCB_ESTATE; }
    if (h->pending.empty()) h->pending.emplace_back();
    SCB_TRY
    SCB_CUDA(cudaSetDevice(h->cfg.device));
    scb::run_flush(h, true);
    float ms = 0;
    SCB_CUDA(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    out->device_ms = ms;
    SCB_CATCH
    scb::fill_result(h, out);
    return SCB_OK;
}

int scb_set_stream(scb_handle *h, void *cuda_stream, int32_t enable) {
    if (!h) { scb::g_last_error = "null handle"; return SCB_EINVAL; }
    cudaSetDevice(h->cfg.device);
    cudaStreamSynchronize(h->st);
    h->st = enable ? (cudaStream_t)cuda_stream : h->st_own;
    return SCB_OK;
}

int scb_shard_info(const scb_handle *h, int32_t *n_cols, int32_t *root_order_pos) {
    if (!h) { scb::g_last_error = "null handle"; return SCB_EINVAL; }
    if (n_cols) *n_cols = h->tab.n_buckets + 1;
    if (root_order_pos) *root_order_pos = h->tab.root_order_pos;
    return SCB_OK;
}

#define SCB_SHARD_ENTER(min_phase)                                                                         \
    if (!h) { scb::g_last_error = "null handle"; return SCB_EINVAL; }                                     \
    if (h->sh_phase < (min_phase)) { scb::g_last_error = "sharded run: call order violated"; return SCB_ESTATE; } \
    SCB_TRY                                                                                                \
    SCB_CUDA(cudaSetDevice(h->cfg.device));

int scb_shard_scan(scb_handle *h, int64_t *n_local) {
    if (!h) { scb::g_last_error = "null handle"; return SCB_EINVAL; }
    if ((uint32_t)h->tab.n_buckets + 1 >= scb::kAuxMaxBuckets) { scb::g_last_error = "sharded run: more than 2^24-2 buckets"; return SCB_EINVAL; }
    if (h->pending.empty()) h->pending.emplace_back();
    SCB_TRY
    SCB_CUDA(cudaSetDevice(h->cfg.device));
    scb::shard_scan(h);
    SCB_CATCH
    if (n_local) *n_local = h->sh_n_local;
    return SCB_OK;
}
int scb_shard_sizes(scb_handle *h, uint64_t carry_in, int32_t chunk_in, uint64_t *carry_out, int32_t *chunk_out) {
    if (!carry_out || !chunk_out) { scb::g_last_error = "null argument"; return SCB_EINVAL; }
    SCB_SHARD_ENTER(1)
    scb::shard_sizes(h, carry_in, chunk_in, carry_out, chunk_out);
    SCB_CATCH
    return SCB_OK;
}
int scb_shard_resolve_local(scb_handle *h, uint32_t *tot_dev) {
    SCB_SHARD_ENTER(1)
    scb::shard_resolve_local(h, tot_dev);
    SCB_CATCH
    return SCB_OK;
}
int scb_shard_resolve_round(scb_handle *h, const uint32_t *before_dev, int64_t reads_before, int32_t first, uint32_t *tot_dev) {
    SCB_SHARD_ENTER(1)
    scb::shard_resolve_round(h, before_dev, reads_before, first, tot_dev);
    SCB_CATCH
    return SCB_OK;
}
int scb_shard_finalize(scb_handle *h, const uint32_t *global_tot_dev, int64_t n_global) {
    SCB_SHARD_ENTER(1)
    if (!h->sh_sized) { scb::g_last_error = "sharded run: scb_shard_sizes must precede scb_shard_finalize"; return SCB_ESTATE; }
    scb::shard_finalize(h, global_tot_dev, n_global);
    SCB_CATCH
    return SCB_OK;
}
int scb_shard_bucket_hist(scb_handle *h, uint32_t *hist_dev) {
    SCB_SHARD_ENTER(2)
    if (!h->sh_resolved) { scb::g_last_error = "sharded run: call order violated (scb_shard_finalize first)"; return SCB_ESTATE; }
    scb::shard_bucket_hist(h, hist_dev);
    SCB_CATCH
    return SCB_OK;
}
static int shard_split_ok(const scb_handle *h, const int64_t *split, int32_t n_ranks) {
    if (split[0] != 0 || split[n_ranks] != h->tab.n_buckets + 1) { scb::g_last_error = "split must cover [0, n_cols]"; return 0; }
    for (int g = 0; g < n_ranks; g++) if (split[g] > split[g + 1]) { scb::g_last_error = "split must be non-decreasing"; return 0; }
    return 1;
}
int scb_shard_partition(scb_handle *h, const int64_t *split, int32_t n_ranks, scb_shard_xfer *out) {
    if (!split || !out || n_ranks < 1 || n_ranks > scb::kMaxRanks) { scb::g_last_error = "bad argument (1..64 ranks)"; return SCB_EINVAL; }
    SCB_SHARD_ENTER(2)
    if (!h->sh_resolved) { scb::g_last_error = "sharded run: call order violated (scb_shard_finalize first)"; return SCB_ESTATE; }
    if (!shard_split_ok(h, split, n_ranks)) return SCB_EINVAL;
    scb::shard_partition(h, split, n_ranks, out);
    SCB_CATCH
    return SCB_OK;
}
// callable from scb_shard_sizes on: ownership by flush chunk does not depend on the tie-break
int scb_shard_partition_chunks(scb_handle *h, const int32_t *chunk_owner, int32_t n_chunks, int32_t n_ranks, scb_shard_xfer *out) {
    if (!chunk_owner || !out || n_chunks < 1 || n_ranks < 1 || n_ranks > scb::kMaxRanks) { scb::g_last_error = "bad argument (1..64 ranks, >= 1 chunk)"; return SCB_EINVAL; }
    for (int c = 0; c < n_chunks; c++)
        if (chunk_owner[c] < 0 || chunk_owner[c] >= n_ranks || (c > 0 && chunk_owner[c] < chunk_owner[c - 1])) {
            scb::g_last_error = "chunk owners must be ranks and must not decrease along the chunk order"; return SCB_EINVAL;
        }
    SCB_SHARD_ENTER(1)
    if (!h->sh_sized || h->sh_phase >= 5) { scb::g_last_error = "sharded run: scb_shard_partition_chunks belongs between scb_shard_sizes and the import"; return SCB_ESTATE; }
    {
        const int64_t c0 = h->sh_layout[0], nn = h->sh_layout[1], n = h->sh_layout[2], tail = h->sh_layout[4];
        const int64_t last = nn > 0 ? (tail > 0 ? c0 + nn : c0 + nn - 1) : c0;          // largest chunk id a local read carries
        if (n > 0 && last >= n_chunks) { scb::g_last_error = "chunk owners do not cover this rank's chunks"; return SCB_EINVAL; }
    }
    scb::shard_partition(h, nullptr, n_ranks, out, chunk_owner, n_chunks);
    SCB_CATCH
    return SCB_OK;
}
int scb_shard_chunk_layout(const scb_handle *h, int64_t *out5) {
    if (!h || !out5) { scb::g_last_error = "bad argument"; return SCB_EINVAL; }
    if (h->sh_phase < 1 || !h->sh_sized) { scb::g_last_error = "sharded run: scb_shard_sizes must precede scb_shard_chunk_layout"; return SCB_ESTATE; }
    for (int k = 0; k < 5; k++) out5[k] = h->sh_layout[k];
    return SCB_OK;
}
int scb_shard_split_mode(const scb_handle *h) { return h ? h->sh_split_mode : -1; }
namespace scb { static int64_t chunk_owners(const std::vector<int64_t> &lay, int G, int n_chunks, std::vector<int32_t> &owner); }
int scb_shard_chunk_owners(const int64_t *layouts, int32_t n_ranks, int32_t n_chunks, int32_t *chunk_owner, int64_t *max_load) {
    if (!layouts || !chunk_owner || n_ranks < 1 || n_ranks > scb::kMaxRanks || n_chunks < 1) { scb::g_last_error = "bad argument (1..64 ranks, >= 1 chunk)"; return SCB_EINVAL; }
    std::vector<int32_t> ow;
    const int64_t ml = scb::chunk_owners(std::vector<int64_t>(layouts, layouts + (size_t)n_ranks * 5), n_ranks, n_chunks, ow);
    for (int c = 0; c < n_chunks; c++) chunk_owner[c] = ow[(size_t)c];
    if (max_load) *max_load = ml;
    return SCB_OK;
}
int scb_shard_pack(scb_handle *h, const int64_t *split, int32_t n_ranks, scb_shard_xfer *out) {
    if (!split || !out || n_ranks < 1 || n_ranks > scb::kMaxRanks) { scb::g_last_error = "bad argument (1..64 ranks)"; return SCB_EINVAL; }
    SCB_SHARD_ENTER(2)
    if (!h->sh_resolved) { scb::g_last_error = "sharded run: call order violated (scb_shard_finalize first)"; return SCB_ESTATE; }
    if (!shard_split_ok(h, split, n_ranks)) return SCB_EINVAL;
    scb::shard_partition(h, split, n_ranks, out);
    const float ms = h->sh_ms;
    scb::shard_stage(h, out);
    h->sh_ms += ms;
    SCB_CATCH
    return SCB_OK;
}
int scb_shard_send(scb_handle *h, int32_t rank, int32_t n_ranks, const scb_shard_peer *peers, int32_t what, int32_t async) {
    if (!peers || (what & ~3) || what == 0) { scb::g_last_error = "bad argument"; return SCB_EINVAL; }
    SCB_SHARD_ENTER(3)
    if (n_ranks != h->sh_G || rank < 0 || rank >= n_ranks) { scb::g_last_error = "rank / n_ranks do not match scb_shard_partition"; return SCB_EINVAL; }
    scb::shard_send(h, rank, peers, what, async != 0);
    SCB_CATCH
    return SCB_OK;
}
int scb_shard_send_wait(scb_handle *h) {
    SCB_SHARD_ENTER(3)
    scb::shard_send_wait(h);
    SCB_CATCH
    return SCB_OK;
}
int scb_shard_recv_reserve(scb_handle *h, const int64_t *need_bytes, void **ptrs, int32_t *changed) {
    if (!h || !need_bytes || !ptrs || !changed) { scb::g_last_error = "null argument"; return SCB_EINVAL; }
    SCB_TRY
    SCB_CUDA(cudaSetDevice(h->cfg.device));
    scb::shard_recv_reserve(h, need_bytes, ptrs, changed);
    SCB_CATCH
    return SCB_OK;
}
int scb_ipc_export(scb_handle *h, const void *dev_ptr, uint8_t *handle64) {
    if (!h || !dev_ptr || !handle64) { scb::g_last_error = "null argument"; return SCB_EINVAL; }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
    SCB_TRY
    SCB_CUDA(cudaSetDevice(h->cfg.device));
    cudaIpcMemHandle_t mh;
    SCB_CUDA(cudaIpcGetMemHandle(&mh, (void *)dev_ptr));
    memcpy(handle64, &mh, 64);
    SCB_CATCH
    return SCB_OK;
}
int scb_ipc_open(scb_handle *h, const uint8_t *handle64, void **out) {
    if (!h || !handle64 || !out) { scb::g_last_error = "null argument"; return SCB_EINVAL; }
    SCB_TRY
    SCB_CUDA(cudaSetDevice(h->cfg.device));
    cudaIpcMemHandle_t mh;
    memcpy(&mh, handle64, 64);
    SCB_CUDA(cudaIpcOpenMemHandle(out, mh, cudaIpcMemLazyEnablePeerAccess));
    SCB_CATCH
    return SCB_OK;
}
int scb_ipc_close(scb_handle *h, void *peer_ptr) {
    if (!h || !peer_ptr) return SCB_OK;
    SCB_TRY
    SCB_CUDA(cudaSetDevice(h->cfg.device));
    SCB_CUDA(cudaIpcCloseMemHandle(peer_ptr));
    SCB_CATCH
    return SCB_OK;
}
int scb_shard_import(scb_handle *h, const scb_shard_xfer *in, int32_t n_chunks_global) {
    if (!in || in->n < 0 || n_chunks_global < 0) { scb::g_last_error = "bad argument"; return SCB_EINVAL; }
    SCB_SHARD_ENTER(4)
    if (!h->sh_resolved || h->sh_aux_pending) { scb::g_last_error = "sharded run: call order violated (the tie-break and the aux-word send precede the import)"; return SCB_ESTATE; }
    scb::shard_import(h, in, n_chunks_global);
    SCB_CATCH
    return SCB_OK;
}
int scb_shard_finish_sort(scb_handle *h) {
    SCB_SHARD_ENTER(5)
    scb::shard_finish(h, 1);
    SCB_CATCH
    return SCB_OK;
}
int scb_shard_finish(scb_handle *h, scb_result *out) {
    if (!out) { scb::g_last_error = "null argument"; return SCB_EINVAL; }
    memset(out, 0, sizeof *out);
    SCB_SHARD_ENTER(5)
    scb::shard_finish(h, h->srt_keys_valid ? 2 : 3);
    h->srt_keys_valid = false;
    out->device_ms = h->sh_ms;
    SCB_CATCH
    scb::fill_result(h, out);
    return SCB_OK;
}
float scb_shard_last_ms(const scb_handle *h) { return h ? h->sh_ms : 0.f; }

int scb_copy_stream(scb_handle *h, int32_t stream, int32_t chunk, void *dst, int64_t dst_bytes) {
    if (!h || stream < 0 || stream >= SCB_N_STREAMS || (!dst && dst_bytes > 0)) { scb::g_last_error = "bad argument"; return SCB_EINVAL; }
    const uint8_t *src; int64_t bytes;
    if (chunk < 0) {
        if (!h->cfg.emit_merged) { scb::g_last_error = "merged output was not requested (emit_merged = 0)"; return SCB_ESTATE; }
        const scb::EmitOut &m = (h->n_chunks > 1) ? h->merged : h->chunked;
        src = m.data[stream].as<uint8_t>(); bytes = m.size[stream];
    } else {
        if (chunk >= h->n_chunks) { scb::g_last_error = "chunk out of range"; return SCB_EINVAL; }
        const auto &co = h->chunked.chunk_off[stream];
        src = h->chunked.data[stream].as<uint8_t>() + co[(size_t)chunk]; bytes = co[(size_t)chunk + 1] - co[(size_t)chunk];
    }
    if (dst_bytes < bytes) { scb::g_last_error = "destination too small"; return SCB_EINVAL; }
    SCB_TRY
    SCB_CUDA(cudaSetDevice(h->cfg.device));
    if (bytes > 0) SCB_CUDA(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDeviceToHost, h->st));
    SCB_CUDA(cudaStreamSynchronize(h->st));
    SCB_CATCH
    return SCB_OK;
}

int scb_copy_debug(scb_handle *h, int32_t *bucket_id, int32_t *core_idx, int32_t *end, int32_t *chunk, uint32_t *perm) {
    if (!h) { scb::g_last_error = "null handle"; return SCB_EINVAL; }
    SCB_TRY
    SCB_CUDA(cudaSetDevice(h->cfg.device));
    const size_t b = (size_t)h->n_last * 4, bp = (size_t)h->n_perm * 4;
    {
        if (bucket_id && b) SCB_CUDA(cudaMemcpyAsync(bucket_id, h->dbg_bucket.p, b, cudaMemcpyDeviceToHost, h->st));
        if (core_idx && b) SCB_CUDA(cudaMemcpyAsync(core_idx, h->dbg_core.p, b, cudaMemcpyDeviceToHost, h->st));
        if (end && b) SCB_CUDA(cudaMemcpyAsync(end, h->dbg_end.p, b, cudaMemcpyDeviceToHost, h->st));
        if (chunk && b) SCB_CUDA(cudaMemcpyAsync(chunk, h->dbg_chunk.p, b, cudaMemcpyDeviceToHost, h->st));
        if (perm && bp) SCB_CUDA(cudaMemcpyAsync(perm, h->perm.p, bp, cudaMemcpyDeviceToHost, h->st));
    }
    SCB_CUDA(cudaStreamSynchronize(h->st));
    SCB_CATCH
    return SCB_OK;
}

int64_t scb_unbucketed(const scb_handle *h) { return h ? h->unbucketed : -1; }

int64_t scb_lifetime_count(scb_handle *h, int32_t core_idx) {
    if (!h) return -1;
    int32_t r = h->tab.n_buckets;
    if (core_idx >= 0) {
        if (core_idx >= (int32_t)h->tab.core_to_rank.size()) return -1;
        r = h->tab.core_to_rank[(size_t)core_idx];
    }
    unsigned long long v = 0;
    cudaSetDevice(h->cfg.device);
    if (cudaMemcpy(&v, h->d_life.as<unsigned long long>() + r, 8, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    return (int64_t)v;
}

int64_t scb_kernel_launches(const scb_handle *) { return scb::g_launches.load(); }

int scb_reset_counts(scb_handle *h) {
    if (!h) { scb::g_last_error = "null handle"; return SCB_EINVAL; }
    SCB_TRY
    SCB_CUDA(cudaSetDevice(h->cfg.device));
    SCB_CUDA(cudaMemsetAsync(h->d_life.p, 0, ((size_t)h->tab.n_buckets + 1) * 8, h->st));
    SCB_CUDA(cudaStreamSynchronize(h->st));
    SCB_CATCH
    h->life_total = 0;
    h->unbucketed = 0;
    for (int m = 0; m < 2; m++) {
        if (h->q_freq3[m].p) { cudaMemsetAsync(h->q_freq3[m].p, 0, h->q_freq3[m].bytes, h->st); cudaMemsetAsync(h->q_freq4[m].p, 0, h->q_freq4[m].bytes, h->st); }
        h->q_prev[m][0] = h->q_prev[m][1] = 500; h->q_seen[m] = 0;
    }
    cudaStreamSynchronize(h->st);
    return SCB_OK;
}

int scb_resolve_rounds(const scb_handle *h) { return h ? h->last_rounds : -1; }

int64_t scb_device_bytes(const scb_handle *h) { return h ? (int64_t)h->arena.peak : -1; }

int scb_resolve_engine(const scb_handle *h) { return h ? scb::pick_engine(h, h->life_total) : -1; }

int scb_stage_ms(const scb_handle *h, float *out, int32_t cap) {
    if (!h || !out) return SCB_EINVAL;
    for (int k = 0; k < SCB_N_STAGES && k < cap; k++) out[k] = h->stage_ms[k];
    return SCB_N_STAGES;
}

// ---- the sharded flush as one call: scalce_b200/shard.py::ShardedTransform.flush in C++ ------------------------------------
namespace scb {
struct CommError { std::string msg; };
static void comm_check(int rc, const char *what) { if (rc != 0) throw CommError{std::string("scb_comm ") + what + " failed"}; }
static std::vector<int64_t> comm_allgather_i64(const scb_comm *cm, const std::vector<int64_t> &mine) {
    std::vector<int64_t> all(mine.size() * (size_t)cm->n_ranks);
    comm_check(cm->allgather(cm->ctx, mine.data(), all.data(), (int64_t)(mine.size() * 8), 0, nullptr), "allgather (host)");
    return all;
}
// contiguous slices of the bucket emission order with about equal read counts (a bucket is never divided)
static std::vector<int64_t> balanced_split(const std::vector<uint64_t> &hist, int G) {
    const int64_t ncols = (int64_t)hist.size();
    std::vector<uint64_t> cum(hist.size());
    uint64_t run = 0;
    for (size_t i = 0; i < hist.size(); i++) { run += hist[i]; cum[i] = run; }
    std::vector<int64_t> split{0};
    for (int g = 1; g < G; g++) {
        const uint64_t target = run * (uint64_t)g / (uint64_t)G;
        int64_t k = (int64_t)(std::upper_bound(cum.begin(), cum.end(), target) - cum.begin());   // first bucket whose inclusive count exceeds the target
        k = std::min(std::max(k, split.back()), ncols);
        split.push_back(k);
    }
    split.push_back(ncols);
    return split;
}
// Whole flush chunks to ranks. lay[g] = scb_shard_chunk_layout of rank g. A chunk that lies inside one rank stays there; a
// chunk that spans ranks goes to the least loaded rank it touches (ties: the rank holding most of it). Chunks and shards are
// both contiguous in the input, so owners never decrease along the chunk order. Returns the largest load (reads).
static int64_t chunk_owners(const std::vector<int64_t> &lay, int G, int n_chunks, std::vector<int32_t> &owner) {
    owner.assign((size_t)std::max(n_chunks, 1), 0);
    std::vector<int64_t> load((size_t)G, 0);
    std::vector<std::vector<std::pair<int, int64_t>>> touch((size_t)n_chunks);      // (rank, reads) of the chunks at shard edges
    for (int g = 0; g < G; g++) {
        const int64_t c0 = lay[(size_t)g * 5], nn = lay[(size_t)g * 5 + 1], n = lay[(size_t)g * 5 + 2], head = lay[(size_t)g * 5 + 3], tail = lay[(size_t)g * 5 + 4];
        if (c0 < n_chunks && head > 0) touch[(size_t)c0].push_back({g, head});
        if (nn > 0 && c0 + nn < n_chunks && tail > 0) touch[(size_t)(c0 + nn)].push_back({g, tail});
        for (int64_t c = c0 + 1; c < c0 + nn && c < n_chunks; c++) owner[(size_t)c] = g;                 // chunks that start and end inside the shard
        load[(size_t)g] += n - head - (nn > 0 ? tail : 0);
    }
    int prev = 0;
    for (int c = 0; c < n_chunks; c++) {
        auto &t = touch[(size_t)c];
        if (!t.empty()) {
            int best = -1; int64_t best_cnt = 0, total = 0;
            for (auto &e : t) {
                total += e.second;
                if (e.first < prev) continue;                  // keeps the owners monotone whatever the loads say
                if (best < 0 || load[(size_t)e.first] < load[(size_t)best] || (load[(size_t)e.first] == load[(size_t)best] && e.second > best_cnt)) { best = e.first; best_cnt = e.second; }
            }
            if (best < 0) best = t.back().first;
            owner[(size_t)c] = best;
            load[(size_t)best] += total;
        } else if (owner[(size_t)c] < prev) {
            owner[(size_t)c] = prev;                           // a chunk without reads (cannot happen below n_chunks): anywhere monotone
        }
        prev = owner[(size_t)c];
    }
    return *std::max_element(load.begin(), load.end());
}
}  // namespace scb

#define SCB_FL(call) do { int _rc = (call); if (_rc != SCB_OK) return _rc; } while (0)

int scb_shard_flush(scb_handle *h, const scb_comm *cm, scb_result *out) {
    using namespace scb;
    if (!h || !cm || !out || !cm->allgather || !cm->barrier || cm->n_ranks < 1 || cm->n_ranks > kMaxRanks || cm->rank < 0 || cm->rank >= cm->n_ranks) {
        g_last_error = "bad argument"; return SCB_EINVAL;
    }
    const int G = cm->n_ranks, r = cm->rank;
    const scb_config &cfg = h->cfg;
    const int L1 = cfg.read_length[0], L2 = cfg.read_length[1];
    float *ms = h->fl_ms;
    for (int k = 0; k < SCB_N_SHARD_PHASES; k++) { ms[k] = 0.f; h->fl_wall[k] = 0.f; }
    auto wall_last = std::chrono::steady_clock::now();
    // wall[k]: host time from the end of the phase before to the end of phase k (device work + collectives + waiting for other ranks)
    auto wall = [&](int k) { const auto now = std::chrono::steady_clock::now(); h->fl_wall[k] += std::chrono::duration<float, std::milli>(now - wall_last).count(); wall_last = now; };
    h->fl_rounds = 0;
    enum { P_SCAN, P_CHUNKS, P_RESOLVE, P_ROUNDS, P_FINALIZE, P_HIST, P_PACK, P_EXCH, P_IMPORT, P_SORT, P_ROWS, P_EMIT };
    auto lap = [&](int k) { ms[k] += scb_shard_last_ms(h); wall(k); };
    try {
        SCB_CUDA(cudaSetDevice(cfg.device));
        cudaStream_t st = h->st;
        int32_t ncols = 0, rootpos = 0;
        SCB_FL(scb_shard_info(h, &ncols, &rootpos));
        // ---- scan ------------------------------------------------------------------------------------------------
        int64_t n_local = 0;
        SCB_FL(scb_shard_scan(h, &n_local)); lap(P_SCAN);
        const std::vector<int64_t> ns = comm_allgather_i64(cm, {n_local});
        std::vector<int64_t> before(1, 0);
        for (int g = 0; g < G; g++) before.push_back(before.back() + ns[(size_t)g]);
        const int64_t n_global = before.back();
        // ---- flush chunks along the global order: the ranks take turns --------------------------------------------
        uint64_t carry = 0; int32_t chunk = 0;
        for (int g = 0; g < G; g++) {
            if (g == r) { SCB_FL(scb_shard_sizes(h, carry, chunk, &carry, &chunk)); lap(P_CHUNKS); }
            const std::vector<int64_t> cc = comm_allgather_i64(cm, {(int64_t)carry, (int64_t)chunk});
            carry = (uint64_t)cc[(size_t)g * 2]; chunk = (int32_t)cc[(size_t)g * 2 + 1];
        }
        const int32_t n_chunks = chunk + (carry > 0 ? 1 : 0);
        // ---- who emits what. Chunk-major output shards by flush chunk: a rank is handed whole chunks, near its own shard, and
        //      only the reads of the chunks at shard edges cross NVLink. That needs enough chunks to balance the ranks, and it
        //      leaves no rank a contiguous piece of the bucket-major merged stream - so with emit_merged, or with few chunks, the
        //      bucket emission order is cut into ranges instead and (G-1)/G of the reads move. SCB_SHARD_SPLIT=chunks|buckets forces.
        std::vector<int32_t> owner;
        bool by_chunk = false;
        {
            int64_t lay5[5] = {0, 0, 0, 0, 0};
            SCB_FL(scb_shard_chunk_layout(h, lay5));
            const std::vector<int64_t> lay = comm_allgather_i64(cm, std::vector<int64_t>(lay5, lay5 + 5));
            const int64_t max_load = n_chunks > 0 ? chunk_owners(lay, G, n_chunks, owner) : 0;
            by_chunk = G > 1 && n_chunks >= 2 && !cfg.emit_merged && (double)max_load * G <= 1.6 * (double)n_global;
            if (const char *e = getenv("SCB_SHARD_SPLIT")) {
                if (!strcmp(e, "chunks") && n_chunks >= 1 && !cfg.emit_merged) by_chunk = true;
                if (!strcmp(e, "buckets")) by_chunk = false;
            }
        }
        // ---- chunk ownership: nothing of the partition depends on the tie-break, so the receive arrays are sized and published and
        //      the quality / mate-2 rows start crossing NVLink (copy engines) now, under the tie-break ---------------------------------
        scb_shard_xfer x;
        memset(&x, 0, sizeof x);
        std::vector<scb_shard_peer> peers((size_t)G);
        scb_shard_xfer y;
        memset(&y, 0, sizeof y);
        auto setup_exchange = [&]() -> int {
            // ---- exchange: what every rank receives from every rank -----------------------------------------------------------
            std::vector<int64_t> mine((size_t)2 * G);
            for (int g = 0; g < G; g++) { mine[(size_t)g] = x.cnt_reads[g]; mine[(size_t)G + g] = x.cnt_name_bytes[g]; }
            const std::vector<int64_t> mat = comm_allgather_i64(cm, mine);      // mat[s * 2G + g] = reads s -> g
            int64_t n_recv = 0, nb_recv = 0;
            for (int s2 = 0; s2 < G; s2++) { n_recv += mat[(size_t)s2 * 2 * G + r]; nb_recv += mat[(size_t)s2 * 2 * G + G + r]; }
            const int prow = x.packed_row_bytes;
            const int64_t need[6] = {n_recv * 8, n_recv * prow + 64, cfg.use_quals ? n_recv * L1 : 0, cfg.use_names ? nb_recv + 16 : 0,
                                     cfg.paired ? n_recv * L2 : 0, (cfg.paired && cfg.use_quals) ? n_recv * L2 : 0};
            void *ptrs[6]; int32_t changed = 0;
            SCB_FL(scb_shard_recv_reserve(h, need, ptrs, &changed));
            // publish the receive arrays: raw pointers between threads of one process, CUDA IPC handles between processes
            const bool have_table = (int)h->fl_table.size() == G;
            const std::vector<int64_t> anyc = comm_allgather_i64(cm, {(int64_t)((changed || !have_table) ? 1 : 0)});
            bool any_changed = false;
            for (int g = 0; g < G; g++) any_changed |= anyc[(size_t)g] != 0;
            if (any_changed) {
                h->fl_table.assign((size_t)G, std::vector<void *>(6, nullptr));
                if (cm->same_process) {
                    std::vector<int64_t> pm(6);
                    for (int k = 0; k < 6; k++) pm[(size_t)k] = (int64_t)(uintptr_t)ptrs[k];
                    const std::vector<int64_t> pa = comm_allgather_i64(cm, pm);
                    for (int g = 0; g < G; g++) for (int k = 0; k < 6; k++) h->fl_table[(size_t)g][(size_t)k] = (void *)(uintptr_t)pa[(size_t)g * 6 + k];
                } else {
                    std::vector<uint8_t> hb(6 * 72, 0), ha((size_t)G * 6 * 72);
                    for (int k = 0; k < 6; k++) if (ptrs[k]) { SCB_FL(scb_ipc_export(h, ptrs[k], &hb[(size_t)k * 72])); hb[(size_t)k * 72 + 64] = 1; }
                    comm_check(cm->allgather(cm->ctx, hb.data(), ha.data(), (int64_t)hb.size(), 0, nullptr), "allgather (host)");
                    for (int g = 0; g < G; g++)
                        for (int k = 0; k < 6; k++) {
                            const uint8_t *e = &ha[((size_t)g * 6 + k) * 72];
                            if (g == r) { h->fl_table[(size_t)g][(size_t)k] = ptrs[k]; continue; }
                            if (!e[64]) continue;
                            auto key = std::make_pair(g, k);
                            auto it = h->fl_ipc.find(key);
                            if (it == h->fl_ipc.end() || memcmp(it->second.first.data(), e, 64) != 0) {
                                if (it != h->fl_ipc.end()) scb_ipc_close(h, it->second.second);
                                void *mp = nullptr;
                                SCB_FL(scb_ipc_open(h, e, &mp));
                                h->fl_ipc[key] = std::make_pair(std::vector<uint8_t>(e, e + 64), mp);
                            }
                            h->fl_table[(size_t)g][(size_t)k] = h->fl_ipc[key].second;
                        }
                }
            } else {
                for (int k = 0; k < 6; k++) h->fl_table[(size_t)r][(size_t)k] = ptrs[k];
            }
            for (int g = 0; g < G; g++) {
                scb_shard_peer &pg = peers[(size_t)g];
                void **t = h->fl_table[(size_t)g].data();
                pg.aux = t[0]; pg.packed = t[1]; pg.qual1 = t[2]; pg.names = t[3]; pg.seq2 = t[4]; pg.qual2 = t[5];
                pg.row_off = 0; pg.name_off = 0;
                for (int s2 = 0; s2 < r; s2++) { pg.row_off += mat[(size_t)s2 * 2 * G + g]; pg.name_off += mat[(size_t)s2 * 2 * G + G + g]; }
            }
            y.n = n_recv; y.name_bytes = nb_recv; y.packed_row_bytes = prow;
            y.aux = (const uint64_t *)ptrs[0]; y.packed = (const uint8_t *)ptrs[1]; y.qual1 = (const uint8_t *)ptrs[2]; y.names = (const uint8_t *)ptrs[3];
            y.seq2 = (const uint8_t *)ptrs[4]; y.qual2 = (const uint8_t *)ptrs[5];
            return SCB_OK;
        };
        if (by_chunk) {
            SCB_FL(scb_shard_partition_chunks(h, owner.data(), n_chunks, G, &x)); lap(P_PACK);
            SCB_FL(setup_exchange());
            // rank 0 is about to resolve its shard alone: its small memsets and copies would queue behind 2.3 ms of row copies on the
            // copy engines (measured: the stream sat idle that long before the tie-break kernel), so ITS rows leave after that, under
            // the joint rounds; every other rank only waits for rank 0 now and sends at once
            if (r != 0) SCB_FL(scb_shard_send(h, r, G, peers.data(), 2, 1));
            wall(P_PACK);          // wall time of the partition phase includes the exchange set-up (counts, receive arrays, IPC handles)
        }
        // ---- tie-break ------------------------------------------------------------------------------------------------
        const int RW = ncols + 1;
        DevBuf tot((size_t)RW * 4, st), all((size_t)G * RW * 4, st), bf((size_t)ncols * 4, st), gtot((size_t)ncols * 4, st), dchg(8, st);
        SCB_CUDA(cudaMemsetAsync(tot.p, 0, (size_t)RW * 4, st));
        if (r == 0) {
            SCB_FL(scb_shard_resolve_local(h, tot.as<uint32_t>())); lap(P_RESOLVE);
            if (by_chunk) SCB_FL(scb_shard_send(h, r, G, peers.data(), 2, 1));
        }
        comm_check(cm->allgather(cm->ctx, tot.p, all.p, (int64_t)RW * 4, 1, st), "allgather (device)");
        int rounds = 0;
        if (G > 1) {
            cudaEvent_t e0, e1;
            SCB_CUDA(cudaEventCreate(&e0)); SCB_CUDA(cudaEventCreate(&e1));
            SCB_CUDA(cudaEventRecord(e0, st));
            for (bool first = true;; first = false) {
                if (r > 0) {
                    if (first) SCB_LAUNCH(sh_guess_k, (unsigned)cdiv(ncols, 256), 256, 0, st, all.as<uint32_t>(), (unsigned long long)before[(size_t)r], (unsigned long long)ns[0], ncols, bf.as<uint32_t>());
                    else SCB_LAUNCH(sh_sum_rows_k, (unsigned)cdiv(ncols, 256), 256, 0, st, all.as<uint32_t>(), RW, 0, r, ncols, bf.as<uint32_t>());
                    SCB_FL(scb_shard_resolve_round(h, bf.as<uint32_t>(), before[(size_t)r], first ? 1 : 0, tot.as<uint32_t>()));
                }
                comm_check(cm->allgather(cm->ctx, tot.p, all.p, (int64_t)RW * 4, 1, st), "allgather (device)");
                rounds++;
                SCB_LAUNCH(sh_changed_k, 1, 1, 0, st, all.as<uint32_t>(), RW, G, ncols, dchg.as<unsigned long long>());
                unsigned long long chg = 0;
                SCB_CUDA(cudaMemcpyAsync(&chg, dchg.p, 8, cudaMemcpyDeviceToHost, st));
                SCB_CUDA(cudaStreamSynchronize(st));
                if (!first && chg == 0) break;            // a round other than the first that changed nothing anywhere: the sequential assignment
                if (rounds >= kRdMaxRounds) throw CudaError{"joint rounds: round cap hit"};
            }
            SCB_CUDA(cudaEventRecord(e1, st));
            SCB_CUDA(cudaStreamSynchronize(st));
            SCB_CUDA(cudaEventElapsedTime(&ms[P_ROUNDS], e0, e1));
            wall(P_ROUNDS);
            cudaEventDestroy(e0); cudaEventDestroy(e1);
        }
        h->fl_rounds = rounds;
        SCB_LAUNCH(sh_sum_rows_k, (unsigned)cdiv(ncols, 256), 256, 0, st, all.as<uint32_t>(), RW, 0, G, ncols, gtot.as<uint32_t>());
        SCB_CUDA(cudaStreamSynchronize(st));
        SCB_FL(scb_shard_finalize(h, gtot.as<uint32_t>(), n_global)); lap(P_FINALIZE);
        if (!by_chunk) {
            // ---- bucket ranges: owners follow from the global bucket histogram ---------------------------------------------------------
            DevBuf hist((size_t)ncols * 4, st), allh((size_t)G * ncols * 4, st);
            SCB_FL(scb_shard_bucket_hist(h, hist.as<uint32_t>())); lap(P_HIST);
            comm_check(cm->allgather(cm->ctx, hist.p, allh.p, (int64_t)ncols * 4, 1, st), "allgather (device)");
            // per-rank histograms fit u32; their sum over the ranks may not: add on the host in 64 bits
            std::vector<uint32_t> hh((size_t)G * ncols);
            SCB_CUDA(cudaMemcpyAsync(hh.data(), allh.p, hh.size() * 4, cudaMemcpyDeviceToHost, st));
            SCB_CUDA(cudaStreamSynchronize(st));
            std::vector<uint64_t> gh((size_t)ncols, 0);
            for (int g = 0; g < G; g++) for (int c2 = 0; c2 < ncols; c2++) gh[(size_t)c2] += hh[(size_t)g * ncols + c2];
            const std::vector<int64_t> split = balanced_split(gh, G);
            SCB_FL(scb_shard_partition(h, split.data(), G, &x)); lap(P_PACK);
            SCB_FL(setup_exchange());
            wall(P_PACK);
        }
        // what the receive side needs to SORT goes first; the quality / mate-2 rows cross NVLink on a side stream (bucket ranges: while
        // the received reads are sorted; chunk ownership: since before the tie-break) and are only awaited before the emit
        SCB_FL(scb_shard_send(h, r, G, peers.data(), 1, 0)); lap(P_EXCH);
        comm_check(cm->barrier(cm->ctx), "barrier");
        SCB_FL(scb_shard_import(h, &y, n_chunks)); lap(P_IMPORT);
        if (!by_chunk) SCB_FL(scb_shard_send(h, r, G, peers.data(), 2, 1));
        SCB_FL(scb_shard_finish_sort(h)); lap(P_SORT);
        SCB_FL(scb_shard_send_wait(h)); lap(P_ROWS);
        comm_check(cm->barrier(cm->ctx), "barrier");           // every rank's row writes have landed
        SCB_FL(scb_shard_finish(h, out)); lap(P_EMIT);
    } catch (scb::CudaError &e) { scb::g_last_error = e.msg; return SCB_ECUDA; }
    catch (scb::CommError &e) { scb::g_last_error = e.msg; return SCB_ECUDA; }
    catch (std::bad_alloc &) { scb::g_last_error = "host allocation failed"; return SCB_ENOMEM; }
    return SCB_OK;
}

int64_t scb_shard_n_local(const scb_handle *h) { return h ? h->sh_n_local : -1; }

int scb_shard_flush_wall(const scb_handle *h, float *wall_ms, int32_t cap) {
    if (!h || !wall_ms) return SCB_EINVAL;
    for (int k = 0; k < SCB_N_SHARD_PHASES && k < cap; k++) wall_ms[k] = h->fl_wall[k];
    return SCB_N_SHARD_PHASES;
}

int scb_shard_flush_stats(const scb_handle *h, float *phase_ms, int32_t cap, int32_t *rounds) {
    if (!h) return SCB_EINVAL;
    for (int k = 0; k < SCB_N_SHARD_PHASES && k < cap; k++) if (phase_ms) phase_ms[k] = h->fl_ms[k];
    if (rounds) *rounds = h->fl_rounds;
    return SCB_N_SHARD_PHASES;
}

int scb_submit_fastq(scb_handle *h, const uint8_t *text1, int64_t bytes1, const uint8_t *text2, int64_t bytes2, int32_t location,
                     const int32_t *phred_offset, int64_t *n_records) {
    if (!h || !phred_offset || !n_records || bytes1 < 0 || (bytes1 > 0 && !text1) || (location != 0 && location != 1)) { scb::g_last_error = "bad argument"; return SCB_EINVAL; }
    if (h->cfg.paired && (bytes2 < 0 || (bytes2 > 0 && !text2))) { scb::g_last_error = "paired configuration: the second FASTQ text is missing"; return SCB_EINVAL; }
    SCB_TRY
    SCB_CUDA(cudaSetDevice(h->cfg.device));
    scb::submit_fastq(h, text1, bytes1, text2, bytes2, location, phred_offset, n_records);
    SCB_CATCH
    return SCB_OK;
}

int scb_quality_stats(scb_handle *h, int32_t mate, uint64_t *freq3, uint64_t *freq4) {
    if (!h || (mate != 0 && mate != 1)) { scb::g_last_error = "bad argument"; return SCB_EINVAL; }
    const size_t n3 = (size_t)scb::kAcDepth * scb::kAcDepth, n4 = n3 * scb::kAcDepth;
    SCB_TRY
    SCB_CUDA(cudaSetDevice(h->cfg.device));
    if (!h->q_freq3[mate].p) {
        if (freq3) memset(freq3, 0, n3 * 8);
        if (freq4) memset(freq4, 0, n4 * 8);
        return SCB_OK;
    }
    if (freq3) SCB_CUDA(cudaMemcpyAsync(freq3, h->q_freq3[mate].p, n3 * 8, cudaMemcpyDeviceToHost, h->st));
    if (freq4) SCB_CUDA(cudaMemcpyAsync(freq4, h->q_freq4[mate].p, n4 * 8, cudaMemcpyDeviceToHost, h->st));
    SCB_CUDA(cudaStreamSynchronize(h->st));
    // the first symbol of a job sets every ac_freq4 entry to 1 (qualities.cpp:192-197); the counts come on top
    if (freq4 && h->q_seen[mate] > 0) for (size_t i = 0; i < n4; i++) freq4[i] += 1;
    SCB_CATCH
    return SCB_OK;
}

int scb_assemble_reads(scb_handle *h, int32_t chunk, const uint8_t **body_dev, int64_t *body_bytes, int64_t *n_segments) {
    if (!h || !body_bytes) { scb::g_last_error = "null argument"; return SCB_EINVAL; }
    if (chunk < 0 && !h->cfg.emit_merged && h->n_chunks > 1) { scb::g_last_error = "merged output was not requested (emit_merged = 0)"; return SCB_ESTATE; }
    if (chunk >= h->n_chunks && !(chunk == 0 && h->n_chunks == 0)) { scb::g_last_error = "chunk out of range"; return SCB_EINVAL; }
    if (h->chunked.chunk_off[1].empty()) { scb::g_last_error = "nothing was flushed yet"; return SCB_ESTATE; }
    SCB_TRY
    SCB_CUDA(cudaSetDevice(h->cfg.device));
    scb::assemble_reads(h, chunk);
    SCB_CATCH
    if (body_dev) *body_dev = h->as_body.as<uint8_t>();
    *body_bytes = h->as_bytes;
    if (n_segments) *n_segments = h->as_nseg;
    return SCB_OK;
}

int scb_copy_assembled(scb_handle *h, void *dst, int64_t dst_bytes, int32_t *seg_core, int64_t *seg_reads) {
    if (!h) { scb::g_last_error = "null handle"; return SCB_EINVAL; }
    if (dst && dst_bytes < h->as_bytes) { scb::g_last_error = "destination too small"; return SCB_EINVAL; }
    SCB_TRY
    SCB_CUDA(cudaSetDevice(h->cfg.device));
    if (dst && h->as_bytes > 0) SCB_CUDA(cudaMemcpyAsync(dst, h->as_body.p, (size_t)h->as_bytes, cudaMemcpyDeviceToHost, h->st));
    if (seg_core && h->as_nseg > 0) SCB_CUDA(cudaMemcpyAsync(seg_core, h->as_seg_core.p, (size_t)h->as_nseg * 4, cudaMemcpyDeviceToHost, h->st));
    if (seg_reads && h->as_nseg > 0) SCB_CUDA(cudaMemcpyAsync(seg_reads, h->as_seg_reads.p, (size_t)h->as_nseg * 8, cudaMemcpyDeviceToHost, h->st));
    SCB_CUDA(cudaStreamSynchronize(h->st));
    SCB_CATCH
    return SCB_OK;
}

int scb_inverse_reads(scb_handle *h, const uint8_t *stream, const int32_t *seg_core, const int64_t *seg_reads, int64_t n_segments, const uint8_t *quals,
                      int32_t mate, int32_t phred_offset, int32_t location, uint8_t *seq_out, uint8_t *qual_out, int64_t *n_reads_out) {
    if (!h || !n_reads_out || n_segments < 0 || (n_segments > 0 && (!stream || !seg_core || !seg_reads || !seq_out)) || (mate != 0 && mate != 1) ||
        (location != 0 && location != 1) || (mate == 1 && !h->cfg.paired)) { scb::g_last_error = "bad argument"; return SCB_EINVAL; }
    SCB_TRY
    SCB_CUDA(cudaSetDevice(h->cfg.device));
    scb::inverse_reads(h, stream, seg_core, seg_reads, n_segments, quals, mate, phred_offset, location, seq_out, qual_out, n_reads_out);
    SCB_CATCH
    return SCB_OK;
}

void scb_destroy(scb_handle *h) {
    if (!h) return;
    cudaSetDevice(h->cfg.device);
    if (h->st) cudaStreamSynchronize(h->st);
    cudaStream_t st = h->st_own;
    cudaEvent_t e0 = h->ev0, e1 = h->ev1;
    for (auto &e : h->stage_ev) if (e) cudaEventDestroy(e);
    for (auto &a : h->st_aux) if (a) { cudaStreamSynchronize(a); cudaStreamDestroy(a); }
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    for (auto &e : h->ev_join) if (e) cudaEventDestroy(e);
    if (h->ev_s0) cudaEventDestroy(h->ev_s0);
    if (h->ev_s1) cudaEventDestroy(h->ev_s1);
    if (h->ev_a0) cudaEventDestroy(h->ev_a0);
    if (h->ev_a1) cudaEventDestroy(h->ev_a1);
    for (auto &kv : h->fl_ipc) cudaIpcCloseMemHandle(kv.second.second);
    for (auto &r : h->rx) if (r) cudaFree(r);
    for (void *r : h->rx_retired) cudaFree(r);
    h->arena.destroy();
    delete h;  // pooled DevBufs free stream-ordered
    if (st) { cudaStreamSynchronize(st); cudaStreamDestroy(st); }
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
}

}  // extern "C"
