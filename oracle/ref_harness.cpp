// ref_harness.cpp - in-memory driver around the UNMODIFIED reference objects
// (reads.o names.o qualities.o buffio.o arithmetic.o const.o compiled from /root/reference by
// oracle/Makefile). TEST INFRASTRUCTURE ONLY: used to time the reference's own transform on the
// host cores (bench.py cpu_baseline, kind "reference") and to cross-check the oracle port.
//
// It repeats, per read, exactly the calls thread() makes after its parse step
// (compress.cpp:673-715): aho_search, output_name, output_read, output_quality,
// aho_trie_bucket, payload memcpy, flush by dump_trie-equivalent (aho_output into six
// buffered_files under `tmp_dir`). refh_run is single-threaded (the reference's only deterministic
// mode, what the parity checks use); refh_run_mt repeats thread()'s locking with T pthreads (-T T) to
// time the reference with all host cores - its output is not deterministic (unlocked read of bin_size
// in aho_search, reads.cpp:420-425) and is never compared with anything.
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <sys/time.h>

#include "arithmetic.h"
#include "buffio.h"
#include "const.h"
#include "names.h"
#include "qualities.h"
#include "reads.h"

// the globals main.cpp defines (main.cpp:62-80); main.o is not linked here
int _quality_sample_lines = 100000;
int _quality_lossy_percentage = 0;
char _use_second_file = 0;
char _is_fasta = 0;
char _use_names = 1;
uint64_t _file_buffer_size = 128 * 1024 * 1024;
uint64_t _max_bucket_set_size = 4LL * 1024LL * 1024LL * 1024LL;
char _temp_directory[MAXLINE] = "__temp__";
char _library_name[MAXLINE] = "";
char _pattern_path[MAXLINE];
int _split_reads = 0;
int _compression_mode = IO_SYS;
char _interleave = 0;
int64_t _time_elapsed = 0;
int _thread_count = 1;
int _decompress = 0;
int _no_ac = 1;
int _compress_qualities = 1;
int32_t read_length[2];
int64_t reads_count = 0;

static double now_s() {
    struct timeval t;
    gettimeofday(&t, 0);
    return t.tv_sec + t.tv_usec * 1e-6;
}

static aho_trie *g_trie = 0;
extern uint8_t *_pool;   // reads.cpp:56, allocated by prepare_aho_automata (reads.cpp:322)

static void dump(int fl, const char *dir, int nf) {  // dump_trie, compress.cpp:524-552
    buffered_file fp[7];
    char path[MAXLINE];
    for (int i = 0; i < nf; i++) {
        snprintf(path, MAXLINE, "%s/t_%03d_%d.tmp", dir, fl, i);
        f_init(&fp[i], IO_SYS);
        f_open(&fp[i], path, IO_WRITE);
    }
    aho_output(g_trie, fp);
    for (int i = 0; i < nf; i++) f_close(fp + i);
}

extern "C" {

// output_quality's input-order statistics (qualities.cpp:186-199) are only kept with arithmetic coding on (_no_ac = 0): the
// CPU test that pins the oracle's restatement of them switches it on, runs refh_run and reads the reference's globals back.
void refh_set_no_ac(int v) { _no_ac = v; }
void refh_get_stats(int mate, uint64_t *f3, uint64_t *f4) {
    memcpy(f3, ac_freq3[mate], sizeof(uint64_t) * AC_DEPTH * AC_DEPTH);
    memcpy(f4, ac_freq4[mate], sizeof(uint64_t) * AC_DEPTH * AC_DEPTH * AC_DEPTH);
}

// Loads the core set (text file, -P) and builds the automaton; returns seconds spent.
double refh_init(const char *cores_path, int L1, int L2, int paired, int use_names, uint64_t bucket_set_bytes) {
    read_length[0] = L1;
    read_length[1] = L2;
    _use_second_file = paired ? 1 : 0;
    _use_names = use_names ? 1 : 0;
    _max_bucket_set_size = bucket_set_bytes;
    if (g_trie && _pool) { free(_pool); _pool = 0; }   // a repeated init allocates a new payload pool; drop the old one
    double t0 = now_s();
    g_trie = read_patterns_from_file(cores_path);
    return now_s() - t0;
}

// Runs n reads given as FASTQ lines in memory: seq/qual rows of L bytes (ASCII, no newline),
// names as "@name" strings via offsets (without '@'). Writes chunk files into tmp_dir.
// Returns seconds spent in the per-read loop + flushes; *n_chunks_out = temp files written.
double refh_run(int64_t n, const uint8_t *seq1, const uint8_t *qual1, const uint8_t *names, const int64_t *name_off,
                const uint8_t *seq2, const uint8_t *qual2, int phred_offset, const char *tmp_dir, int *n_chunks_out,
                int32_t *bucket_id_out, int32_t *end_out) {
    const int L1 = read_length[0], L2 = read_length[1];
    quality_mapping qmap[2];
    for (int m = 0; m < 2; m++) {
        qmap[m].offset = phred_offset;
        for (int c = 0; c < 128; c++) qmap[m].values[c] = c;
    }
    static char read[MAXLINE], name[MAXLINE], qual[MAXLINE], read2[MAXLINE], qual2b[MAXLINE];
    static uint8_t out[MAXLINE * 5];
    read_data rd;
    rd.data = out;
    uint64_t total_size = 0;
    int temp_file_count = 0;
    const int nf = 4 + 2 * _use_second_file;
    double t0 = now_s();
    for (int64_t i = 0; i < n; i++) {
        // what the parse step leaves in the line buffers (compress.cpp:614-671); not part of the path
        memcpy(read, seq1 + i * L1, L1); read[L1] = '\n'; read[L1 + 1] = 0;
        memcpy(qual, qual1 + i * L1, L1); qual[L1] = '\n'; qual[L1 + 1] = 0;
        int nl = (int)(name_off[i + 1] - name_off[i]);
        name[0] = '@'; memcpy(name + 1, names + name_off[i], nl); name[nl + 1] = '\n'; name[nl + 2] = 0;
        if (_use_second_file) {
            memcpy(read2, seq2 + i * L2, L2); read2[L2] = '\n'; read2[L2 + 1] = 0;
            memcpy(qual2b, qual2 + i * L2, L2); qual2b[L2] = '\n'; qual2b[L2 + 1] = 0;
        }
        aho_trie *bucket;
        int p = aho_search(read, g_trie, &bucket);                       // compress.cpp:673
        rd.sz = output_name(name, rd.data);
        if (p != -1) {
            rd.sz += output_read(read, rd.data + rd.sz, p - bucket->level + 1, bucket->level);
            rd.end = p + 1;
        } else {
            rd.sz += output_read(read, rd.data + rd.sz, 0, 0);
            rd.end = 0;
        }
        if (_compress_qualities) rd.sz += output_quality(qual, read, qmap + 0, rd.data + rd.sz, 0);
        rd.of = rd.sz;
        if (_use_second_file) {
            rd.sz += output_read(read2, rd.data + rd.sz, 0, 0);
            if (_compress_qualities) rd.sz += output_quality(qual2b, read2, qmap + 1, rd.data + rd.sz, 1);
        }
        bin_node *bn = aho_trie_bucket(bucket, &rd);
        total_size += rd.sz + sizeof(bin_node);
        memcpy(bn->data.data, rd.data, rd.sz);
        if (bucket_id_out) bucket_id_out[i] = bucket->output >= 0 ? bucket->id : MAXBIN - 1;
        if (end_out) end_out[i] = rd.end;
        if (total_size >= _max_bucket_set_size) {
            dump(temp_file_count++, tmp_dir, nf);
            total_size = 0;
        }
    }
    if (total_size) dump(temp_file_count++, tmp_dir, nf);
    double dt = now_s() - t0;
    if (n_chunks_out) *n_chunks_out = temp_file_count;
    return dt;
}

// ---- the same loop with T threads, locks as in thread() (compress.cpp:600-717) ----------------------
struct mt_job {
    int64_t n, next;
    const uint8_t *seq1, *qual1, *names, *seq2, *qual2;
    const int64_t *name_off;
    quality_mapping qmap[2];
    uint64_t total_size;
    int temp_file_count, nf;
    const char *tmp_dir;
    pthread_spinlock_t r_spin, w_spin;
};

static void *mt_worker(void *arg) {
    mt_job *J = (mt_job *)arg;
    const int L1 = read_length[0], L2 = read_length[1];
    char *read = new char[MAXLINE], *name = new char[MAXLINE], *qual = new char[MAXLINE], *read2 = new char[MAXLINE], *qual2b = new char[MAXLINE];
    uint8_t *out = new uint8_t[MAXLINE * 5];
    read_data rd;
    rd.data = out;
    while (1) {
        // the parse step holds r_spin in the reference (compress.cpp:614-671); here it is the copy into line buffers
        pthread_spin_lock(&J->r_spin);
        const int64_t i = J->next;
        if (i >= J->n) { pthread_spin_unlock(&J->r_spin); break; }
        J->next = i + 1;
        memcpy(read, J->seq1 + i * L1, L1); read[L1] = '\n'; read[L1 + 1] = 0;
        memcpy(qual, J->qual1 + i * L1, L1); qual[L1] = '\n'; qual[L1 + 1] = 0;
        int nl = (int)(J->name_off[i + 1] - J->name_off[i]);
        name[0] = '@'; memcpy(name + 1, J->names + J->name_off[i], nl); name[nl + 1] = '\n'; name[nl + 2] = 0;
        if (_use_second_file) {
            memcpy(read2, J->seq2 + i * L2, L2); read2[L2] = '\n'; read2[L2 + 1] = 0;
            memcpy(qual2b, J->qual2 + i * L2, L2); qual2b[L2] = '\n'; qual2b[L2 + 1] = 0;
        }
        pthread_spin_unlock(&J->r_spin);

        aho_trie *bucket;
        int p = aho_search(read, g_trie, &bucket);                       // compress.cpp:673, no lock
        rd.sz = output_name(name, rd.data);
        if (p != -1) {
            rd.sz += output_read(read, rd.data + rd.sz, p - bucket->level + 1, bucket->level);
            rd.end = p + 1;
        } else {
            rd.sz += output_read(read, rd.data + rd.sz, 0, 0);
            rd.end = 0;
        }
        pthread_spin_lock(&J->w_spin);                                   // compress.cpp:688-704
        if (_compress_qualities) rd.sz += output_quality(qual, read, J->qmap + 0, rd.data + rd.sz, 0);
        rd.of = rd.sz;
        if (_use_second_file) {
            rd.sz += output_read(read2, rd.data + rd.sz, 0, 0);
            if (_compress_qualities) rd.sz += output_quality(qual2b, read2, J->qmap + 1, rd.data + rd.sz, 1);
        }
        bin_node *bn = aho_trie_bucket(bucket, &rd);
        J->total_size += rd.sz + sizeof(bin_node);
        pthread_spin_unlock(&J->w_spin);
        memcpy(bn->data.data, rd.data, rd.sz);
        if (J->total_size >= _max_bucket_set_size) {                     // compress.cpp:708-715
            pthread_spin_lock(&J->w_spin);
            if (J->total_size >= _max_bucket_set_size) {
                dump(J->temp_file_count++, J->tmp_dir, J->nf);
                J->total_size = 0;
            }
            pthread_spin_unlock(&J->w_spin);
        }
    }
    delete[] read; delete[] name; delete[] qual; delete[] read2; delete[] qual2b; delete[] out;
    return 0;
}

// Same inputs as refh_run; `threads` pthreads as the reference's -T. Returns seconds (loop + flushes).
double refh_run_mt(int64_t n, const uint8_t *seq1, const uint8_t *qual1, const uint8_t *names, const int64_t *name_off,
                   const uint8_t *seq2, const uint8_t *qual2, int phred_offset, const char *tmp_dir, int *n_chunks_out, int threads) {
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    mt_job J;
    J.n = n; J.next = 0;
    J.seq1 = seq1; J.qual1 = qual1; J.names = names; J.name_off = name_off; J.seq2 = seq2; J.qual2 = qual2;
    for (int m = 0; m < 2; m++) {
        J.qmap[m].offset = phred_offset;
        for (int c = 0; c < 128; c++) J.qmap[m].values[c] = c;
    }
    J.total_size = 0; J.temp_file_count = 0; J.nf = 4 + 2 * _use_second_file; J.tmp_dir = tmp_dir;
    pthread_spin_init(&J.r_spin, 0);
    pthread_spin_init(&J.w_spin, 0);
    _thread_count = threads;
    pthread_t th[256];
    double t0 = now_s();
    for (int t = 0; t < threads; t++) pthread_create(&th[t], 0, mt_worker, &J);
    for (int t = 0; t < threads; t++) pthread_join(th[t], 0);
    if (J.total_size) dump(J.temp_file_count++, tmp_dir, J.nf);
    double dt = now_s() - t0;
    _thread_count = 1;
    pthread_spin_destroy(&J.r_spin);
    pthread_spin_destroy(&J.w_spin);
    if (n_chunks_out) *n_chunks_out = J.temp_file_count;
    return dt;
}

int refh_unbucketed(void) { return unbuck(); }

}  // extern "C"
