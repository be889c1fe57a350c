// scb_glue.h - the INTEGRATION.md glue in compilable form: what a maintainer adds to compress.cpp to put libscalce_b200.so
// behind the scalce CLI. make_dropin.py textually inserts this header and five call sites into a COPY of the reference's
// compress.cpp (the copy lives under oracle/_ref/dropin/, never in the repository) and links it with the reference's other
// objects, unmodified. The result, oracle/_ref/scalce_scb, is the reference CLI with only the boosting transform replaced
// (compress.cpp:673-715 per-read search / bucket insert, 524-552 + 799-801 flush, 732-733 core-set load, 821 free, 834 unbuck).
// TEST INFRASTRUCTURE: tests/test_gpu_dropin.py byte-compares its .scalce{n,r,q} files with the unmodified CLI's.
//
// Included in the middle of compress.cpp, after its globals: MAXLINE, ERROR, LOG, read_length, _use_names, _use_second_file,
// _compress_qualities, _max_bucket_set_size, _temp_directory, _thread_count, temp_file_count, file_reads, patterns,
// output_name, output_quality, buffered_file and the f_* functions are the reference's own.
#pragma once
#include <vector>

#include "scalce_b200.h"

static scb_handle *g_scb = 0;                  // replaces `aho_trie *trie`
struct ScbSoa {                                // one batch of parsed reads, filled by the parse loop
    std::vector<uint8_t> seq1, qual1, names, seq2, qual2;
    std::vector<int64_t> name_off;
    int64_t n;
    ScbSoa() : name_off(1, 0), n(0) {}
};
static ScbSoa g_soa;
static const int64_t SCB_BATCH_READS = 1 << 20;

static void scb_check(int rc) {
    if (rc) ERROR("%s\n", scb_last_error());
}

// replaces read_patterns_from_file / read_patterns (compress.cpp:732-733); runs after get_quality_stats because the
// library wants the read lengths at creation. Also fills patterns[] (reads.h:48), which the container writer
// (compress.cpp:373) and nothing else of compress() reads.
extern char _binary_patterns_bin_start;
extern char _binary_patterns_bin_end;
static void scb_glue_create(const char *pattern_path) {
    if (_thread_count != 1) ERROR("the GPU transform reproduces the reference at -T 1 (its only deterministic mode): run with -T 1\n");
    scb_config cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.read_length[0] = read_length[0];
    cfg.read_length[1] = _use_second_file ? read_length[1] : 0;
    cfg.use_names = _use_names ? 1 : 0;
    cfg.paired = _use_second_file ? 1 : 0;
    cfg.use_quals = _compress_qualities ? 1 : 0;
    cfg.bucket_set_bytes = _max_bucket_set_size;
    cfg.device = 0;
    cfg.emit_merged = 0;                       // merge() stays the reference's
    char tmp[MAXLINE];
    if (!pattern_path[0]) {                    // the embedded core set (Makefile:67-68): hand it over as a file
        snprintf(tmp, MAXLINE, "%s/patterns_embedded.bin", _temp_directory);
        FILE *f = fopen(tmp, "wb");
        if (!f) ERROR("cannot write %s\n", tmp);
        fwrite(&_binary_patterns_bin_start, 1, (size_t)(&_binary_patterns_bin_end - &_binary_patterns_bin_start), f);
        fclose(f);
        pattern_path = tmp;
    }
    scb_check(scb_create_from_file(pattern_path, &cfg, &g_scb));
    if (pattern_path == tmp) remove(tmp);
    int32_t nc = 0;
    scb_check(scb_table_info(g_scb, &nc, 0, 0, 0));
    patterns = (char **)malloc(sizeof(char *) * (size_t)(nc + 1));
    for (int32_t i = 0; i < nc; i++) patterns[i] = strdup(scb_core(g_scb, i));
}

static void scb_glue_submit() {                // replaces compress.cpp:673-706 for a whole batch
    if (!g_soa.n) return;
    scb_batch b;
    memset(&b, 0, sizeof b);
    b.n = g_soa.n;
    b.seq1 = g_soa.seq1.data();
    b.qual1 = _compress_qualities ? g_soa.qual1.data() : 0;
    b.names = _use_names ? g_soa.names.data() : 0;
    b.name_off = _use_names ? g_soa.name_off.data() : 0;
    b.seq2 = _use_second_file ? g_soa.seq2.data() : 0;
    b.qual2 = (_use_second_file && _compress_qualities) ? g_soa.qual2.data() : 0;
    b.location = 0;
    scb_check(scb_submit(g_scb, &b));
    g_soa = ScbSoa();
}

// per read, after the parse (compress.cpp:614-671): output_name (names.cpp:48-62) and output_quality
// (qualities.cpp:177-204, including its input-order statistics) stay the reference's; their bytes are payload
static void scb_glue_append(char *name, char *read, char *qual, char *read2, char *qual2, quality_mapping *qm) {
    uint8_t tmp[MAXLINE];
    int nl = output_name(name, tmp);
    if (_use_names) {
        g_soa.names.insert(g_soa.names.end(), tmp + 1, tmp + nl);      // without the length byte
        g_soa.name_off.push_back((int64_t)g_soa.names.size());
    }
    g_soa.seq1.insert(g_soa.seq1.end(), (uint8_t *)read, (uint8_t *)read + read_length[0]);
    if (_compress_qualities) {
        int nq = output_quality(qual, read, qm + 0, tmp, 0);
        g_soa.qual1.insert(g_soa.qual1.end(), tmp, tmp + nq);
    }
    if (_use_second_file) {
        g_soa.seq2.insert(g_soa.seq2.end(), (uint8_t *)read2, (uint8_t *)read2 + read_length[1]);
        if (_compress_qualities) {
            int nq = output_quality(qual2, read2, qm + 1, tmp, 1);
            g_soa.qual2.insert(g_soa.qual2.end(), tmp, tmp + nq);
        }
    }
    file_reads++;
    if (++g_soa.n == SCB_BATCH_READS) scb_glue_submit();
}

// replaces every dump_trie (compress.cpp:708-715, 799-801): one flush at end of input, chunk c stream k -> t_%03d_<k>.tmp
static void scb_glue_flush() {
    scb_glue_submit();
    scb_result r;
    scb_check(scb_flush(g_scb, &r));
    int nf = 4 + 2 * (_use_second_file ? 1 : 0);
    std::vector<char> buf;
    for (int c = 0; c < r.n_chunks; c++)
        for (int k = 0; k < nf; k++) {
            int64_t bytes = r.chunk_off[k][c + 1] - r.chunk_off[k][c];
            buf.resize(bytes ? (size_t)bytes : 1);
            scb_check(scb_copy_stream(g_scb, k, c, buf.data(), bytes));
            char path[MAXLINE];
            snprintf(path, MAXLINE, "%s/t_%03d_%d.tmp", _temp_directory, c, k);
            buffered_file f;
            f_init(&f, IO_SYS);
            f_open(&f, path, IO_WRITE);
            if (bytes) f_write(&f, buf.data(), bytes);
            f_close(&f);
        }
    temp_file_count = r.n_chunks;              // merge() and combine_and_compress_with_split() run unchanged
}

static int scb_glue_unbucketed() { return (int)scb_unbucketed(g_scb); }
static void scb_glue_destroy() { scb_destroy(g_scb); g_scb = 0; }
