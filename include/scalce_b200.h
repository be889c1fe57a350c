/*
 * scalce_b200.h - C ABI of the B200-native SCALCE boosting transform
 * (core scan -> stateful bucket pick -> bucket histogram -> stable reorder of sequence,
 * quality and name payloads).
 *
 * The reference (sfu-compbio/scalce) has no plugin or FFI layer; its seam is the set of free
 * functions in reads.h:87-94 driven per read from compress.cpp:600-717 and flushed by
 * dump_trie (compress.cpp:524-552). This header is the batch form of that seam: plain
 * pointers and sizes, no C++/torch types. Each entry point names what it replaces.
 *
 * Conventions
 *   - every function returns 0 on success, a negative SCB_E* code on failure; the text of the
 *     last failure is available from scb_last_error(). The reference's convention is
 *     ERROR() -> "(ERROR) ..." on stderr + exit(1) (const.h:77-81); the CLI glue maps a non-zero
 *     return to that.
 *   - there is NO CPU fallback: if no sm_100 device is usable, scb_create fails with
 *     SCB_ENODEVICE.
 *   - bit-exactness contract: results equal the reference run with -T 1 (its only deterministic
 *     mode): per-read bucket, end marker, flush chunk, final permutation and every stream byte.
 *   - one submitting thread per handle.
 */
#ifndef SCALCE_B200_H_
#define SCALCE_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SCB_ABI_VERSION 2

#define SCB_OK 0
#define SCB_EINVAL -1    /* bad argument / unsupported configuration */
#define SCB_ENODEVICE -2 /* no usable CUDA device (no fallback exists) */
#define SCB_ECUDA -3     /* CUDA runtime error, see scb_last_error() */
#define SCB_ENOMEM -4    /* host or device allocation failed */
#define SCB_ESTATE -5    /* call order violated */
#define SCB_EIO -6       /* core-set file could not be read */

/* id/core written for the "no core" (root) bucket: MAXBIN-1, const.h:94, reads.cpp:161-164 */
#define SCB_ROOT_ID ((1 << 30) - 1)

/* longest read: what the reference's line buffer holds (fgets into MAXLINE = 2500 bytes: text + newline + NUL, const.h:87) */
#define SCB_MAX_READ_LENGTH 2498

/* stream numbering = temp-file numbering t_%03d_<k>.tmp (compress.cpp:527-533, reads.cpp:91-180) */
enum { SCB_S_NAMES = 0, SCB_S_READS = 1, SCB_S_QUALS = 2, SCB_S_META = 3, SCB_S_READS2 = 4, SCB_S_QUALS2 = 5, SCB_N_STREAMS = 6 };

typedef struct scb_handle scb_handle;

/* Replaces the globals the path reads: read_length[2], _use_names, _use_second_file,
 * _compress_qualities, _max_bucket_set_size (const.h:100-119, main.cpp:62-80). */
typedef struct scb_config {
    int32_t read_length[2];    /* L of mate 1, mate 2 (0 when single-end), 1..SCB_MAX_READ_LENGTH; fixed-length build only */
    int32_t use_names;         /* 1: names are payload (stream 0); 0: -n mode, one 0 byte per read in size accounting */
    int32_t paired;            /* _use_second_file */
    int32_t use_quals;         /* _compress_qualities (0 with -Q / -f) */
    int32_t device;            /* CUDA device ordinal */
    uint64_t bucket_set_bytes; /* -B, flush threshold of compress.cpp:708 (default 4 GiB) */
    int32_t emit_merged;       /* 1: additionally produce the post-merge order (merge(), compress.cpp:488-522) */
    int32_t reserved;
} scb_config;

/* One batch of parsed reads in input order: what thread() holds per read after the parse and
 * after output_name / output_quality (compress.cpp:614-699), as a structure of arrays.
 *   seq*   ASCII bases, n rows of read_length[m] bytes, no terminator (any case; N and anything
 *          that is not C/G/T maps to A, const.cpp:47-49)
 *   qual*  the bytes output_quality produced (q - offset, 0 under 'N'; qualities.cpp:177-204);
 *          opaque payload to this library. NULL when use_quals == 0.
 *   names  concatenated name characters as output_name keeps them (after '@', up to the first
 *          space; names.cpp:48-62) WITHOUT the length byte; name_off[n+1] byte offsets into it.
 *          A name longer than 255 bytes is an error (the reference's length byte wraps).
 *          NULL when use_names == 0.
 * location: 0 = host memory (pageable or pinned), 1 = device memory on cfg.device. Device arrays are read in whole 16-byte
 * granules: each must be readable up to the next 16-byte boundary past its last byte (true of anything cudaMalloc or a
 * framework's caching allocator returns; a sub-range ending in the middle of a larger array is fine too). */
typedef struct scb_batch {
    int64_t n;
    const uint8_t *seq1, *qual1;
    const uint8_t *names;
    const int64_t *name_off;
    const uint8_t *seq2, *qual2;
    int32_t location;
    int32_t reserved;
} scb_batch;

/* Result of a flush. Streams are byte-identical to what bin_dump writes into the temp files:
 * stream k of flush chunk c is data[k][chunk_off[k][c] .. chunk_off[k][c+1]).
 * Pointers are DEVICE pointers owned by the handle, valid until the next submit/flush/destroy;
 * chunk_off arrays are host memory owned by the handle. Use scb_copy_stream for host copies. */
typedef struct scb_result {
    int64_t n_reads;
    int32_t n_chunks;
    int32_t n_buckets_nonempty;       /* distinct non-empty buckets over the whole flush; filled when the flush has one chunk or
                                         emit_merged is set (it is the number of merged meta records), 0 otherwise */
    const uint8_t *data[SCB_N_STREAMS];
    const int64_t *chunk_off[SCB_N_STREAMS];
    /* post-merge order (emit_merged): one set of streams, meta already merged */
    const uint8_t *merged[SCB_N_STREAMS];
    int64_t merged_size[SCB_N_STREAMS];
    /* per-read arrays in INPUT order (device): bucket id as written to meta (BFS node id,
     * SCB_ROOT_ID for none), core index (-1 for none), end marker, flush chunk; and the final
     * permutation in OUTPUT order (perm[j] = input index of the j-th emitted read). */
    const int32_t *bucket_id, *core_idx, *end, *chunk;
    const uint32_t *perm;
    /* device-side time of the last flush in milliseconds (CUDA events on the handle's stream) */
    float device_ms;
} scb_result;

/* Replaces read_patterns_from_file + prepare_aho_automata (reads.cpp:379-410, 270-324) for a
 * core set given as strings; core index = position in the array. */
int scb_create(const char *const *cores, int32_t n_cores, const scb_config *cfg, scb_handle **out);
/* Same, loading the set from disk: text (one core per whitespace-separated token, -P file,
 * reads.cpp:388-394) or patterns.bin records (reads.cpp:338-369); format is sniffed. */
int scb_create_from_file(const char *path, const scb_config *cfg, scb_handle **out);
/* Host-only: compiles the core set into the automaton tables without touching a device and reports
 * their shape: states (trie nodes incl. root), buckets (distinct cores), where aho_output would emit
 * the root bucket (== n_buckets unless some base starts no core), and each core's BFS node id - the
 * "id" bin_dump writes (reads.cpp:166, 296). Needs no GPU; used by the CPU-side tests. */
int scb_table_dryrun(const char *const *cores, int32_t n_cores, int32_t *n_states, int32_t *n_buckets, int32_t *root_order_pos,
                     int32_t *core_node_id);
/* Number of cores / automaton states / whether the transition table is shared-memory resident. */
int scb_table_info(const scb_handle *h, int32_t *n_cores, int32_t *n_states, int32_t *n_buckets, int32_t *smem_resident);
/* patterns[i] (reads.h:48): NUL-terminated core string by index, owned by the handle. */
const char *scb_core(const scb_handle *h, int32_t idx);

/* Replaces the per-read calls aho_search / output_read / aho_trie_bucket and the size accounting
 * (compress.cpp:673-706): appends the batch, in order, to the pending set. */
int scb_submit(scb_handle *h, const scb_batch *batch);
/* Replaces dump_trie / aho_output / bin_prepare / bin_dump (compress.cpp:524-552,
 * reads.cpp:466-499, 600-634, 91-180) for everything pending; lifetime bucket counts
 * (aho_trie::bin_size, reads.h:82) persist in the handle across flushes. */
int scb_flush(scb_handle *h, scb_result *out);
/* Streaming form for jobs that do not fit one flush (inputs beyond HBM, or more than 2^31-1 reads): emits only the flush chunks
 * that are COMPLETE and keeps the reads of the open chunk pending - they are taken out of the bucket populations again and are
 * decided anew by the next flush, together with what is submitted in between. Chunk boundaries, chunk contents, in-chunk order
 * and lifetime counts of a job fed through any sequence of scb_submit / scb_flush_closed calls and one final scb_flush are
 * therefore those of one scb_flush over the whole input, i.e. the reference's (compress.cpp:702-713 carries total_size across
 * reads and input files; a caller cannot align plain flushes with chunk boundaries because rd.sz depends on the core picked on
 * the device). out->n_reads = reads emitted (0 when no chunk closed yet); chunk numbers restart at 0 in every result: the caller
 * numbers its temp files consecutively. Needs emit_merged = 0. Not part of the sharded run (scb_shard_sizes carries the sum). */
int scb_flush_closed(scb_handle *h, scb_result *out);
/* Device-to-host copy of one stream slice (chunk >= 0) or of the merged stream (chunk = -1). */
int scb_copy_stream(scb_handle *h, int32_t stream, int32_t chunk, void *dst, int64_t dst_bytes);
/* Copies per-read arrays of the last flush to host; any pointer may be NULL. */
int scb_copy_debug(scb_handle *h, int32_t *bucket_id, int32_t *core_idx, int32_t *end, int32_t *chunk, uint32_t *perm);
/* unbuck() (reads.cpp:502): reads flushed from the root bucket so far. */
int64_t scb_unbucketed(const scb_handle *h);
/* Lifetime count of a core's bucket (bin_size), -1 = root. */
int64_t scb_lifetime_count(scb_handle *h, int32_t core_idx);
/* Number of this library's kernel launches in this PROCESS so far, over all handles (an atomic process-wide counter;
 * the handle argument is ignored). bench.py's gpu_launches is a difference of two readings. */
int64_t scb_kernel_launches(const scb_handle *h);
/* Device time of each stage of the last flush, CUDA events on the handle's stream, milliseconds:
 * 0 scan, 1 resolve, 2 size accounting + chunk ids, 3 key build + sort, 4 tie refinement,
 * 5 emit (per-chunk streams), 6 merged order + emit, 7 per-read arrays. Returns SCB_N_STAGES. */
#define SCB_N_STAGES 8
int scb_stage_ms(const scb_handle *h, float *out, int32_t cap);
/* Forgets all lifetime bucket populations and the unbucketed count (a new compression job with the
 * same core set); keeps the automaton and the device workspace. */
int scb_reset_counts(scb_handle *h);
/* Fixed-point rounds the parallel tie-break needed in the last flush (0 = sequential engine used). */
int scb_resolve_rounds(const scb_handle *h);
/* High-water mark of the flush workspace (bytes of device memory handed out by the handle's slab in one flush, inputs
 * copied by scb_submit not included): what to size a flush against the GPU's memory with. */
int64_t scb_device_bytes(const scb_handle *h);
/* Which tie-break engine serves this core set (all are exact, aho_search's population compare, reads.cpp:420-421):
 * 0 dense (per-warp population rows in shared memory, <= ~6.4k buckets), 1 sparse (bucket-major candidate lists in
 * global memory, core sets of production size: the reference sizes patterns[] for 5-10 M cores, reads.cpp:336, 385),
 * 2 sequential (one warp; jobs of >= 2^32 - 1 reads). The sharded run works with 0 and 1. */
int scb_resolve_engine(const scb_handle *h);

/* =====================================================================================================
 * Sharded run: ONE flush whose input is the concatenation of the ranks' submissions in rank order
 * (one handle per GPU, one process per GPU). The reference has no counterpart (single process); the
 * contract is that concatenating the ranks' outputs in rank order gives, per flush chunk and per
 * stream, exactly what ONE handle fed the whole input would have produced (= the reference at -T 1).
 * The library does the per-GPU work; the caller moves bytes between ranks (NCCL all-gather of a few
 * KB per resolve round, one all-to-all of the payload) - scalce_b200/shard.py is that caller.
 * Call order on every rank:
 *   scb_submit...                                    this rank's shard, in input order
 *   scb_shard_scan                                   core scan (aho_search minus the populations)
 *   scb_shard_sizes                                  rd.sz accounting; ranks in order, each passing
 *                                                    (carry, chunk) on to the next (compress.cpp:702-713)
 *   scb_shard_resolve_local  (first rank)            tie-break over its own shard, then
 *   scb_shard_resolve_round  (the others, repeated)  one global fixed-point round per call, until a round
 *                                                    other than the first changes nothing on any rank
 *   scb_shard_finalize                               bucket / end marker per read, lifetime counts
 *   scb_shard_bucket_hist -> all-reduce -> split of the bucket emission order into contiguous slices
 *   scb_shard_pack -> all-to-all -> scb_shard_import
 *   scb_shard_finish                                 stable sort + emit of the owned bucket slice
 * Restrictions: fewer than 2^24 buckets. Every rank owns ALL
 * flush chunks of its buckets, so emit_merged works per rank.
 * ===================================================================================================== */

/* Device arrays of one side of the exchange, destination-major (send) or source-major (receive); rows keep
 * input order inside a destination/source. aux: one u64 per read (bucket rank | end << 24 | name length
 * << 36 | flush chunk << 44); packed: 2-bit rows of packed_row_bytes; names: name bytes without length
 * bytes. cnt_* are host arrays [n_ranks] owned by the handle (send side only). */
typedef struct scb_shard_xfer {
    int64_t n;
    int64_t name_bytes;
    const uint64_t *aux;
    const uint8_t *packed, *qual1, *names, *seq2, *qual2;
    const int64_t *cnt_reads, *cnt_name_bytes;
    int32_t packed_row_bytes;
    int32_t reserved;
} scb_shard_xfer;

/* enable = 1: runs the handle's work on the caller's CUDA stream (cudaStream_t; NULL = the legacy default
 * stream) instead of its own, so that the caller's events and collectives order with it.
 * enable = 0: back to the handle's own stream. */
int scb_set_stream(scb_handle *h, void *cuda_stream, int32_t enable);
/* Number of bucket columns a resolve histogram has: n_buckets + 1 (root last); tot arrays hold one more
 * word (the changed count). Emission position of the root bucket in *root_order_pos. */
int scb_shard_info(const scb_handle *h, int32_t *n_cols, int32_t *root_order_pos);
/* Scan of everything pending (aho_search's walk, reads.cpp:413-429, without the population compare). */
int scb_shard_scan(scb_handle *h, int64_t *n_local);
/* rd.sz + sizeof(bin_node) accounting and flush-chunk ids numbered along the global order. */
int scb_shard_sizes(scb_handle *h, uint64_t carry_in, int32_t chunk_in, uint64_t *carry_out, int32_t *chunk_out);
/* Tie-break of a shard that starts the input order (populations = lifetime counts). tot_dev: device
 * u32[n_cols + 1], receives the shard's bucket histogram (rank order) and 0 in the last word. */
int scb_shard_resolve_local(scb_handle *h, uint32_t *tot_dev);
/* One global round for a later shard. before_dev: device u32[n_cols] populations of all lower ranks'
 * shards under the current global assignment (a guess in the first round); reads_before: their read
 * count; first: 1 for the first round. tot_dev as above, last word = decisions that changed. */
int scb_shard_resolve_round(scb_handle *h, const uint32_t *before_dev, int64_t reads_before, int32_t first, uint32_t *tot_dev);
/* Commits the converged assignment. global_tot_dev: device u32[n_cols], bucket histogram summed over all
 * ranks (adds to the lifetime counts of EVERY rank's handle); n_global: reads of the whole flush.
 * Exception: the no-core (root) bucket is counted by every rank for its own shard only - nothing reads it for the
 * tie-break - so after a sharded flush scb_unbucketed() and scb_lifetime_count(h, -1) are rank-local: sum them over
 * the ranks for the value unbuck() (reads.cpp:502) would report. */
int scb_shard_finalize(scb_handle *h, const uint32_t *global_tot_dev, int64_t n_global);
/* Local bucket histogram in emission order, device u32[n_cols] (overwritten). */
int scb_shard_bucket_hist(scb_handle *h, uint32_t *hist_dev);
/* Stable partition of the local reads by owner; split[g] = first emission position owned by rank g,
 * split[0] = 0, split[n_ranks] = n_cols (host array). Fills *out with the send arrays. */
int scb_shard_pack(scb_handle *h, const int64_t *split, int32_t n_ranks, scb_shard_xfer *out);
/* Fused pack + exchange over peer memory (the production path; scb_shard_pack + a caller-side all-to-all is the
 * staged alternative):
 *   scb_shard_partition   owner of every local read, per-owner counts (out->cnt_*), aux words and names staged
 *   (ranks all-gather the counts; every rank sizes its receive arrays and publishes them)
 *   scb_shard_recv_reserve  persistent receive arrays of this rank (cudaMalloc; grown with headroom). need_bytes
 *                         / ptrs are indexed 0 aux, 1 packed, 2 qual1, 3 names, 4 seq2, 5 qual2; *changed = 1 when
 *                         any array moved (its IPC handle must be re-published)
 *   scb_ipc_export / scb_ipc_open / scb_ipc_close   CUDA IPC plumbing for those arrays (64-byte handles)
 *   scb_shard_send        row gathers that write straight into the owners' receive arrays: peers[g] holds owner
 *                         g's arrays as mapped in THIS process and the row / name-byte offset at which this
 *                         rank's reads start there. Returns when this rank's writes are complete; a barrier over
 *                         the ranks then makes every receive array complete.
 * followed by scb_shard_import of the rank's own receive arrays. */
typedef struct scb_shard_peer {
    void *aux, *packed, *qual1, *names, *seq2, *qual2;
    int64_t row_off, name_off;
} scb_shard_peer;
int scb_shard_partition(scb_handle *h, const int64_t *split, int32_t n_ranks, scb_shard_xfer *out);
/* The other way to hand out the emit work: whole FLUSH CHUNKS instead of bucket ranges. The output of the transform is
 * chunk-major (one set of temp files per flush chunk, compress.cpp:708-713), chunks and shards are both contiguous pieces of
 * the input, so a rank that is handed the chunks lying in (or mostly in) its own shard keeps most of its reads and only the
 * reads of the chunks at shard edges cross NVLink - against (n_ranks-1)/n_ranks of all reads for bucket ranges. After the
 * exchange a rank holds every read of its chunks and emits them exactly as one handle would; the other ranks' pieces of those
 * chunks are empty, so "rank-order concatenation per chunk" still describes the job's output.
 *   scb_shard_chunk_layout      after scb_shard_sizes: out5 = {first chunk id of this shard, chunks that START in it, reads,
 *                               reads belonging to the first chunk id, reads belonging to the last started chunk}
 *   scb_shard_partition_chunks  replaces scb_shard_partition; chunk_owner[n_chunks] (host) = rank that emits each chunk,
 *                               non-decreasing (then the send order is the input order, every array of a destination is one
 *                               contiguous copy and no owner sort is needed). Nothing of it depends on the tie-break: it may
 *                               be called right after scb_shard_sizes, and scb_shard_send(what = 2) - the quality / mate-2
 *                               rows, most of the bytes - may then run UNDER the tie-break. The aux words (bucket, end
 *                               marker) are packed by the first scb_shard_send that carries them (what & 1), which
 *                               therefore needs scb_shard_finalize before it.
 * The bucket-major merged stream (emit_merged) has no contiguous per-rank piece under this ownership: orchestrators keep
 * bucket ranges when it is requested. scb_shard_split_mode: ownership used by the handle's last sharded flush (0 bucket
 * ranges, 1 flush chunks). scb_shard_flush picks chunks when emit_merged = 0 and the chunks balance the ranks within 1.6x. */
int scb_shard_chunk_layout(const scb_handle *h, int64_t *out5);
int scb_shard_partition_chunks(scb_handle *h, const int32_t *chunk_owner, int32_t n_chunks, int32_t n_ranks, scb_shard_xfer *out);
int scb_shard_split_mode(const scb_handle *h);
/* The owner rule scb_shard_flush applies (host arithmetic, no device): layouts = the ranks' scb_shard_chunk_layout outputs in
 * rank order [n_ranks x 5]. A chunk inside one shard stays with that rank; a chunk spanning shards goes to the least loaded rank
 * it touches (ties: the one holding most of it). *max_load = reads of the busiest rank. */
int scb_shard_chunk_owners(const int64_t *layouts, int32_t n_ranks, int32_t n_chunks, int32_t *chunk_owner, int64_t *max_load);
int scb_shard_recv_reserve(scb_handle *h, const int64_t *need_bytes, void **ptrs, int32_t *changed);
/* what: 1 = aux words + 2-bit rows + names (all the receive side needs to SORT), 2 = quality / mate-2 rows, 3 = both.
 * async = 0: returns when this rank's writes are complete. async = 1: the writes run on a side stream with a small
 * grid and the call returns at once - so that, after a first synchronous send of `what = 1`, a barrier and
 * scb_shard_import, the row exchange overlaps scb_shard_finish_sort; scb_shard_send_wait joins it (then a barrier,
 * then scb_shard_finish emits). */
int scb_shard_send(scb_handle *h, int32_t rank, int32_t n_ranks, const scb_shard_peer *peers, int32_t what, int32_t async);
int scb_shard_send_wait(scb_handle *h);
/* Sort + tie refinement of the imported reads (needs aux, 2-bit rows and names only); optional: scb_shard_finish
 * runs it if it was not called. */
int scb_shard_finish_sort(scb_handle *h);
int scb_ipc_export(scb_handle *h, const void *dev_ptr, uint8_t *handle64);
int scb_ipc_open(scb_handle *h, const uint8_t *handle64, void **out);
int scb_ipc_close(scb_handle *h, void *peer_ptr);
/* Adopts the received arrays (caller-owned device memory, must stay valid until the next submit; the
 * packed array needs 64 readable bytes past its last row). */
int scb_shard_import(scb_handle *h, const scb_shard_xfer *in, int32_t n_chunks_global);
/* Sort + emit of the owned slice; result as scb_flush (per-read arrays refer to the rank's INPUT shard,
 * perm to the received order). */
int scb_shard_finish(scb_handle *h, scb_result *out);
/* Device milliseconds of the last scb_shard_* call (CUDA events on the handle's stream). */
float scb_shard_last_ms(const scb_handle *h);

/* ---- the whole sharded flush as ONE call (the C++ orchestrator; scalce_b200/shard.py is the same sequence in Python) ----
 * Communication between the ranks is the caller's: MPI, NCCL (libscalce_b200_nccl.so wraps an ncclComm_t into this struct),
 * or threads of one process. The library needs two collectives:
 *   allgather  every rank contributes `bytes` from `send` and receives n_ranks * bytes in rank order in `recv`.
 *              device = 1: both buffers are device memory of the calling rank's GPU and the operation must be ordered on
 *              `stream` (a cudaStream_t: the handle's stream); device = 0: host memory, blocking.
 *   barrier    all ranks arrive (host side); every rank has synchronised its stream before it calls this.
 * same_process = 1: the ranks are threads of one process on GPUs with peer access (or one GPU): device pointers are
 * exchanged as they are; otherwise the receive arrays are published through CUDA IPC (scb_ipc_export / scb_ipc_open).
 * Both callbacks return 0 on success. */
typedef struct scb_comm {
    int32_t rank, n_ranks;
    int32_t same_process;
    int32_t reserved;
    void *ctx;
    int (*allgather)(void *ctx, const void *send, void *recv, int64_t bytes, int32_t device, void *stream);
    int (*barrier)(void *ctx);
} scb_comm;
/* Everything between scb_submit of this rank's shard and the rank's slice of the output: scan, flush chunks along the global
 * order, joint tie-break (one all-gather of the bucket histograms per round), bucket-range split, fused pack + exchange over
 * peer memory with the row exchange overlapped by the sort, emit. Collective: every rank calls it once per flush with its own
 * handle. Result as scb_shard_finish. scb_shard_flush_stats: device milliseconds of the phases of the last call, in the order
 * scan, chunks, resolve (rank 0 alone), resolve_rounds (joint), finalize, hist, pack, exchange, import, sort, exchange_rows,
 * emit; *rounds = joint rounds. */
#define SCB_N_SHARD_PHASES 12
int scb_shard_flush(scb_handle *h, const scb_comm *comm, scb_result *out);
int scb_shard_flush_stats(const scb_handle *h, float *phase_ms, int32_t cap, int32_t *rounds);
/* Host wall time per phase of the last scb_shard_flush, same order: from the end of the phase before to the end of the phase,
 * i.e. device work + collectives + waiting for the other ranks. Their sum is the duration of the call. */
int scb_shard_flush_wall(const scb_handle *h, float *wall_ms, int32_t cap);
/* Reads of this rank's own input shard in the last sharded flush (the per-read arrays of scb_result refer to them). */
int64_t scb_shard_n_local(const scb_handle *h);

/* =====================================================================================================
 * The transform's two neighbours that SURVEY.md 8(f) ranks next, on the device.
 * ===================================================================================================== */
/* (f2) The host front end on the device: FASTQ text -> one pending batch, replacing for a whole text at once the parse loop of
 * thread() (compress.cpp:614-671: four lines per record, '@' name line, read and quality lines of read_length characters),
 * output_name (names.cpp:48-62: the characters after '@' up to the first space) and output_quality at lossy percentage 0
 * (qualities.cpp:177-204: quality - phred_offset, 0 under an upper-case 'N'), including output_quality's input-order context
 * statistics ac_freq3 / ac_freq4 (what the arithmetic coder's model is built from, arithmetic.cpp) carried across calls.
 * text1 / text2 (text2 only in a paired configuration; its name lines are skipped as the reference does): plain FASTQ, any
 * number of whole records; location 0 = host, 1 = device. phred_offset[2]: per mate, what quality_mapping_init detected
 * (qualities.cpp:99-104). Errors (SCB_ECUDA with the reason in scb_last_error): a record count that is not whole, a name line
 * without '@', a read or quality line of another length (the reference exits on those, compress.cpp:629-636), a name longer
 * than 255 bytes, a quality symbol outside [offset, offset + 80) (AC_DEPTH, arithmetic.h:47). A batch submitted this way and
 * batches submitted through scb_submit may be mixed; only this entry point feeds the statistics.
 * scb_quality_stats: ac_freq3 [80*80] and ac_freq4 [80*80*80] of a mate as the reference holds them after the same input
 * (uint64 each; either pointer may be NULL); scb_reset_counts clears them. */
int scb_submit_fastq(scb_handle *h, const uint8_t *text1, int64_t bytes1, const uint8_t *text2, int64_t bytes2, int32_t location,
                     const int32_t *phred_offset, int64_t *n_records);
int scb_quality_stats(scb_handle *h, int32_t mate, uint64_t *freq3, uint64_t *freq4);
/* (f3) Bucket-record assembly of the .scalcer body, replacing the per-bucket loop of combine_and_compress_with_split
 * (compress.cpp:345-384) for mate 1: for every non-empty bucket `int32 core, int64 n_reads` (n_reads = tR / record size,
 * compress.cpp:371-376) followed by that bucket's packed reads + end markers - everything the reference writes to
 * <out>_1.scalcer after its 16 header bytes (magic, _no_ac, read_length; compress.cpp:289-291), as ONE contiguous
 * buffer. chunk = -1: from the merged streams of the last flush (emit_merged, or a single flush chunk); chunk >= 0: from
 * that flush chunk's streams. *body_dev (may be NULL) receives a device pointer owned by the handle, valid until the next
 * scb_assemble_reads / scb_destroy. scb_copy_assembled copies the body and/or the segment table (core index - SCB_ROOT_ID
 * for the no-core bucket - and reads per bucket record) to host; any pointer may be NULL. */
int scb_assemble_reads(scb_handle *h, int32_t chunk, const uint8_t **body_dev, int64_t *body_bytes, int64_t *n_segments);
int scb_copy_assembled(scb_handle *h, void *dst, int64_t dst_bytes, int32_t *seg_core, int64_t *seg_reads);
/* (f4) The decompress-side inverse (decompress.cpp:331-352): bucket-ordered packed reads -> the ASCII rows the
 * decompressor writes, in stream order. stream: mate 0 = packed reads + end markers WITHOUT the inline bucket headers
 * (stream 1 / the .scalcer body minus its 12-byte records); mate 1 = stream 4 (mate-2 reads: unrotated, no end marker).
 * seg_core / seg_reads [n_segments]: core index and reads of every bucket record, in stream order (what the inline headers
 * hold). quals (may be NULL): the quality payload in the same order, n x L bytes; where it is 0 the base becomes 'N'
 * (decompress.cpp:350-351) and qual_out (may be NULL) receives payload + phred_offset. seq_out / qual_out: n x L bytes,
 * row pitch L, no newlines. location: 0 = all pointers host, 1 = all device (on cfg.device). */
int scb_inverse_reads(scb_handle *h, const uint8_t *stream, const int32_t *seg_core, const int64_t *seg_reads, int64_t n_segments, const uint8_t *quals,
                      int32_t mate, int32_t phred_offset, int32_t location, uint8_t *seq_out, uint8_t *qual_out, int64_t *n_reads_out);

/* aho_trie_free (reads.cpp:505-535). */
void scb_destroy(scb_handle *h);

const char *scb_last_error(void);
int scb_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* SCALCE_B200_H_ */
