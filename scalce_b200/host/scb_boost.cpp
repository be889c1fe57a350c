// scb_boost.cpp - C++ host side of the boundary: FASTQ in, the reference's temp files out, the transform itself
// through the C ABI (include/scalce_b200.h). This is the glue INTEGRATION.md describes, as a standalone program:
//
//   parse            compress.cpp:614-671  (name line, read line, '+' line, quality line; mates in lock step)
//   sample           qualities.cpp:60-104  (phred offset from the first 100000 quality lines; read length)
//   name payload     names.cpp:48-62       (text after '@' up to the first space or newline)
//   quality payload  qualities.cpp:177-204 (q - offset, 0 under an 'N' base; lossy percentage 0)
//   search/bucket    compress.cpp:673-706  -> scb_submit per batch of reads
//   flush            compress.cpp:524-552, 708-715, 799-801 -> scb_flush, chunk c stream k -> t_%03d_<k>.tmp
//   container        compress.cpp:262-379 (combine_and_compress_with_split in raw mode, -c no -A): the merged streams
//                    -> PREFIX_<mate>.scalce{n,r,q}   (--container PREFIX; no entropy coding here - that stays the host's)
//
// It links libscalce_b200.so only; there is no CPU path for the transform (scb_create fails without a B200).
// --assemble DIR reads merged_<k>.tmp from DIR instead of running the transform and only writes the container (no GPU).
// --dump-soa DIR stops after the host stages and writes the batch arrays (what scb_submit would get): that mode
// needs no GPU and is what the CPU tests check (tests/test_host_cpu.py).
#include <cerrno>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/scalce_b200.h"

namespace {

[[noreturn]] void die(const std::string &m) {   // the reference's convention, const.h:77-81
    fprintf(stderr, "(ERROR) %s\n", m.c_str());
    exit(1);
}
void scb_check(int rc) { if (rc) die(scb_last_error()); }
void write_file(const std::string &path, const void *p, size_t bytes);

struct LineReader {   // plain-text FASTQ; lines without the terminator
    FILE *f = nullptr;
    std::vector<char> buf;
    explicit LineReader(const char *path) : buf(1 << 16) {
        f = fopen(path, "rb");
        if (!f) die(std::string("cannot open ") + path + ": " + strerror(errno));
        setvbuf(f, nullptr, _IOFBF, 8 << 20);
    }
    ~LineReader() { if (f) fclose(f); }
    bool next(std::string &out) {
        out.clear();
        while (fgets(buf.data(), (int)buf.size(), f)) {
            size_t l = strlen(buf.data());
            if (l && buf[l - 1] == '\n') { out.append(buf.data(), l - 1); if (!out.empty() && out.back() == '\r') out.pop_back(); return true; }
            out.append(buf.data(), l);
        }
        return !out.empty();
    }
    void rewind_file() { rewind(f); }
};

struct Soa {   // one batch, structure of arrays (scb_batch)
    std::vector<uint8_t> seq1, qual1, names, seq2, qual2;
    std::vector<int64_t> name_off{0};
    int64_t n = 0;
    void clear() { seq1.clear(); qual1.clear(); names.clear(); seq2.clear(); qual2.clear(); name_off.assign(1, 0); n = 0; }
};

struct Sample { int offset = 64; int L = 0; };

// quality_mapping_init, qualities.cpp:60-104: histogram of the quality characters of the first `lines` records;
// offset 33 if any character in [33, 64) occurs, else 64; the read length is that of the last sampled record
Sample sample_file(const char *path, int lines) {
    LineReader r(path);
    std::string a, b, c, d;
    long stat[256] = {0};
    Sample s;
    for (int i = 0; i < lines; i++) {
        if (!r.next(a) || !r.next(b) || !r.next(c) || !r.next(d)) break;
        for (unsigned char ch : d) stat[ch]++;
        s.L = (int)d.size();
    }
    for (int i = 33; i < 64; i++) if (stat[i]) { s.offset = 33; break; }
    return s;
}

std::vector<uint8_t> read_file(const std::string &path) {
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) die("cannot open " + path + ": " + strerror(errno));
    std::vector<uint8_t> v;
    uint8_t buf[1 << 16];
    size_t r;
    while ((r = fread(buf, 1, sizeof buf, f)) > 0) v.insert(v.end(), buf, buf + r);
    fclose(f);
    return v;
}

template <typename T> void put(std::vector<uint8_t> &o, T v) { const uint8_t *p = (const uint8_t *)&v; o.insert(o.end(), p, p + sizeof(T)); }

// combine_and_compress_with_split, compress.cpp:262-379, raw mode (_no_ac = 1, IO_SYS): per mate three files.
//   .scalcen  magic, use_names byte, [name index 0 + library name when names are off], then the name records
//   .scalcer  magic, int32 no_ac, int32 read length, then per bucket [mate 1 only: int32 core, int64 reads] + its packed reads
//   .scalceq  magic, int64 phred offset, then the quality bytes
// streams: the merged (bucket-order) streams 0..5; meta records: int32 id, int32 core, int64 lN, lR, lQ [, lR2, lQ2].
// core_len(core) = length of core `core` (patterns[core], compress.cpp:373); the root bucket has core = MAXBIN - 1.
template <typename CoreLen>
void write_containers(const std::string &prefix, const std::vector<uint8_t> *streams, CoreLen core_len, int L1, int L2, bool paired, bool use_names,
                      int64_t phred_offset, const std::string &library) {
    static const uint8_t magic[8] = {'s', 'c', 'a', 'l', 'c', 'e', '2', '2'};
    const int nlen = 3 + 2 * (paired ? 1 : 0);
    const size_t rsz = 8 + 8 * (size_t)nlen;
    const int sz_meta = L1 > 255 ? 2 : 1;
    const std::vector<uint8_t> &meta = streams[3];
    if (meta.size() % rsz) die("meta stream is not a whole number of records");
    for (int mate = 0; mate < (paired ? 2 : 1); mate++) {
        std::vector<uint8_t> fn(magic, magic + 8), fr(magic, magic + 8), fq(magic, magic + 8);
        fn.push_back(use_names ? 1 : 0);
        if (!use_names) { put<int64_t>(fn, 0); fn.insert(fn.end(), library.begin(), library.end()); }
        put<int32_t>(fr, 1);                                   // _no_ac
        put<int32_t>(fr, mate ? L2 : L1);
        put<int64_t>(fq, phred_offset);
        const std::vector<uint8_t> &sr = streams[mate ? 4 : 1], &sq = streams[mate ? 5 : 2], &sn = streams[0];
        size_t pn = 0, pr = 0, pq = 0;
        for (size_t o = 0; o < meta.size(); o += rsz) {
            int32_t core;
            int64_t len[5] = {0, 0, 0, 0, 0};
            memcpy(&core, &meta[o + 4], 4);
            memcpy(len, &meta[o + 8], 8 * (size_t)nlen);
            const int64_t lN = len[0];
            int64_t lR = len[1], lQ = len[2];
            if (mate) { lR = len[3]; lQ = len[4]; }
            else {
                const int clen = core == (1 << 30) - 1 ? 0 : core_len(core);
                const int64_t per = (((L1 - clen) >> 2) + (((L1 - clen) & 3) != 0)) + sz_meta;
                put<int32_t>(fr, core);
                put<int64_t>(fr, lR / per);
            }
            if (pr + (size_t)lR > sr.size() || pq + (size_t)lQ > sq.size() || (use_names && pn + (size_t)lN > sn.size())) die("meta records run past the streams");
            fr.insert(fr.end(), sr.begin() + pr, sr.begin() + pr + lR); pr += (size_t)lR;
            fq.insert(fq.end(), sq.begin() + pq, sq.begin() + pq + lQ); pq += (size_t)lQ;
            if (use_names) { fn.insert(fn.end(), sn.begin() + pn, sn.begin() + pn + lN); pn += (size_t)lN; }
        }
        const std::string m = std::to_string(mate + 1);
        write_file(prefix + "_" + m + ".scalcen", fn.data(), fn.size());
        write_file(prefix + "_" + m + ".scalcer", fr.data(), fr.size());
        write_file(prefix + "_" + m + ".scalceq", fq.data(), fq.size());
    }
}

void write_file(const std::string &path, const void *p, size_t bytes) {
    FILE *f = fopen(path.c_str(), "wb");
    if (!f) die("cannot write " + path + ": " + strerror(errno));
    if (bytes && fwrite(p, 1, bytes, f) != bytes) die("short write to " + path);
    fclose(f);
}

}  // namespace

int main(int argc, char **argv) {
    const char *in1 = nullptr, *in2 = nullptr, *cores = nullptr, *out = nullptr, *dump = nullptr, *container = nullptr, *assemble = nullptr;
    std::string library;
    int asm_L1 = 0, asm_L2 = 0, asm_paired = 0; long long asm_offset = 33;
    uint64_t bucket_set = 4ull << 30;   // -B, main.cpp default
    int use_names = 1, use_quals = 1, merged = 0, device = 0, sample_lines = 100000, device_parse = 0;
    int64_t batch_reads = 4 << 20;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        auto need = [&](const char *what) { if (i + 1 >= argc) die(std::string("missing value for ") + what); return argv[++i]; };
        if (a == "-r") in2 = need("-r");
        else if (a == "-P") cores = need("-P");
        else if (a == "-o") out = need("-o");
        else if (a == "-B") bucket_set = strtoull(need("-B"), nullptr, 10);
        else if (a == "-n") use_names = 0;
        else if (a == "--no-quals") use_quals = 0;
        else if (a == "--merged") merged = 1;
        else if (a == "--device-parse") device_parse = 1;   // the parse loop / output_name / output_quality on the GPU (scb_submit_fastq)
        else if (a == "--device") device = atoi(need("--device"));
        else if (a == "--batch") batch_reads = atoll(need("--batch"));
        else if (a == "--dump-soa") dump = need("--dump-soa");
        else if (a == "--container") container = need("--container");
        else if (a == "--library") library = need("--library");
        else if (a == "--assemble") assemble = need("--assemble");
        else if (a == "--L1") asm_L1 = atoi(need("--L1"));
        else if (a == "--L2") { asm_L2 = atoi(need("--L2")); asm_paired = 1; }
        else if (a == "--offset") asm_offset = atoll(need("--offset"));
        else if (a[0] == '-') die("unknown option " + a);
        else if (!in1) in1 = argv[i];
        else die("more than one input file (use -r for mate 2)");
    }
    if (assemble) {   // container assembly alone, from merged_<k>.tmp in a directory (no GPU): scb_boost --assemble DIR --container PREFIX -P cores.txt --L1 n [--L2 n] [-n] [--offset o]
        if (!container || !cores || asm_L1 <= 0) die("--assemble needs --container PREFIX, -P cores.txt and --L1");
        std::vector<std::string> cv;
        {
            LineReader cr(cores);
            std::string ln;
            while (cr.next(ln)) if (!ln.empty()) cv.push_back(ln);
        }
        std::vector<uint8_t> st[6];
        for (int k = 0; k < 4 + 2 * asm_paired; k++) st[k] = read_file(std::string(assemble) + "/merged_" + std::to_string(k) + ".tmp");
        write_containers(container, st, [&](int32_t c) { if (c < 0 || (size_t)c >= cv.size()) die("core index out of range"); return (int)cv[(size_t)c].size(); },
                         asm_L1, asm_L2, asm_paired != 0, use_names != 0, asm_offset, library);
        return 0;
    }
    if (container) merged = 1;
    if (!in1 || (!out && !dump && !container) || (!cores && !dump))
        die("usage: scb_boost in_1.fastq [-r in_2.fastq] -P cores.txt -o out_dir [-B bytes] [-n] [--no-quals] [--merged] [--batch reads] [--device d] [--device-parse] | --dump-soa dir");
    if (batch_reads < 1) die("--batch must be positive");

    const Sample s1 = sample_file(in1, sample_lines);
    Sample s2;
    if (in2) s2 = sample_file(in2, sample_lines);
    if (s1.L <= 0 || (in2 && s2.L <= 0)) die("empty input");
    const int L1 = s1.L, L2 = in2 ? s2.L : 0;

    scb_handle *h = nullptr;
    if (!dump) {
        scb_config cfg;
        memset(&cfg, 0, sizeof cfg);
        cfg.read_length[0] = L1; cfg.read_length[1] = L2;
        cfg.use_names = use_names; cfg.paired = in2 ? 1 : 0; cfg.use_quals = use_quals;
        cfg.device = device; cfg.bucket_set_bytes = bucket_set; cfg.emit_merged = merged;
        scb_check(scb_create_from_file(cores, &cfg, &h));
    }

    if (device_parse && !dump) {
        // the file goes to the device as text, in pieces of whole records; the library does the parse loop
        // (compress.cpp:614-671), output_name and output_quality. Everything after the submit is the same as below.
        auto slurp = [](const char *path) {
            FILE *f = fopen(path, "rb");
            if (!f) die(std::string("cannot open ") + path + ": " + strerror(errno));
            std::vector<uint8_t> v;
            std::vector<uint8_t> tmp(8 << 20);
            size_t r;
            while ((r = fread(tmp.data(), 1, tmp.size(), f)) > 0) v.insert(v.end(), tmp.begin(), tmp.begin() + r);
            fclose(f);
            return v;
        };
        const std::vector<uint8_t> t1 = slurp(in1), t2 = in2 ? slurp(in2) : std::vector<uint8_t>();
        // pieces of `batch_reads` records: advance 4 * batch_reads line ends in each text
        auto advance = [](const std::vector<uint8_t> &t, size_t from, int64_t records) {
            int64_t lines = 4 * records;
            size_t p = from;
            while (p < t.size() && lines > 0) { if (t[p] == '\n') lines--; p++; }
            return p;
        };
        const int32_t ph[2] = {s1.offset, in2 ? s2.offset : s1.offset};
        size_t p1 = 0, p2 = 0;
        int64_t n_total_dev = 0;
        while (p1 < t1.size()) {
            const size_t e1 = advance(t1, p1, batch_reads), e2 = in2 ? advance(t2, p2, batch_reads) : 0;
            int64_t got = 0;
            scb_check(scb_submit_fastq(h, t1.data() + p1, (int64_t)(e1 - p1), in2 ? t2.data() + p2 : nullptr, in2 ? (int64_t)(e2 - p2) : 0, 0, ph, &got));
            n_total_dev += got;
            p1 = e1; p2 = e2;
        }
        fprintf(stderr, "scb_boost: %lld records parsed on the device\n", (long long)n_total_dev);
    }
    LineReader r1(in1);
    LineReader *r2 = in2 ? new LineReader(in2) : nullptr;
    Soa soa, all;   // `all` only in --dump-soa mode
    std::string name, read, plus, qual, name2, read2, qual2;
    int64_t n_total = 0;
    auto submit = [&]() {
        if (!soa.n) return;
        if (dump) {
            all.seq1.insert(all.seq1.end(), soa.seq1.begin(), soa.seq1.end());
            all.qual1.insert(all.qual1.end(), soa.qual1.begin(), soa.qual1.end());
            all.seq2.insert(all.seq2.end(), soa.seq2.begin(), soa.seq2.end());
            all.qual2.insert(all.qual2.end(), soa.qual2.begin(), soa.qual2.end());
            const int64_t base = (int64_t)all.names.size();
            all.names.insert(all.names.end(), soa.names.begin(), soa.names.end());
            for (size_t i = 1; i < soa.name_off.size(); i++) all.name_off.push_back(base + soa.name_off[i]);
            all.n += soa.n;
        } else {
            scb_batch b;
            memset(&b, 0, sizeof b);
            b.n = soa.n; b.seq1 = soa.seq1.data(); b.qual1 = use_quals ? soa.qual1.data() : nullptr;
            b.names = use_names ? soa.names.data() : nullptr; b.name_off = use_names ? soa.name_off.data() : nullptr;
            b.seq2 = in2 ? soa.seq2.data() : nullptr; b.qual2 = (in2 && use_quals) ? soa.qual2.data() : nullptr;
            b.location = 0;
            scb_check(scb_submit(h, &b));   // returns after the copy: the vectors can be reused
        }
        soa.clear();
    };
    while (!(device_parse && !dump) && r1.next(name)) {
        if (!r1.next(read)) break;
        if (read.empty()) {   // compress.cpp:620-625 (the reference has not consumed the '+' / quality lines at that point either)
            fprintf(stderr, "Whooops... %s is empty, skipping it!\n", name.c_str());
            continue;
        }
        if ((int)read.size() != L1) die("read lengths in /1 do not match (" + std::to_string(L1) + " vs " + std::to_string(read.size()) + ")");   // compress.cpp:629-636
        if (!r1.next(plus) || !r1.next(qual)) die("truncated record in " + std::string(in1));
        if ((int)qual.size() != L1) die("quality line length differs from the read length");
        if (r2) {
            if (!r2->next(name2) || !r2->next(read2)) die("mate 2 file is shorter than mate 1");
            if ((int)read2.size() != L2) die("read lengths in /2 do not match");
            if (!r2->next(plus) || !r2->next(qual2)) die("truncated record in " + std::string(in2));
            if ((int)qual2.size() != L2) die("quality line length differs from the read length (mate 2)");
        }
        // output_name, names.cpp:48-62: text after '@' up to the first space (or end of line)
        if (use_names) {
            size_t e = 1;
            while (e < name.size() && name[e] != ' ') e++;
            if (name.empty()) e = 1;
            const size_t nl = name.empty() ? 0 : e - 1;
            if (nl > 255) die("read name longer than 255 bytes");
            soa.names.insert(soa.names.end(), name.begin() + (name.empty() ? 0 : 1), name.begin() + (name.empty() ? 0 : e));
            soa.name_off.push_back((int64_t)soa.names.size());
        }
        soa.seq1.insert(soa.seq1.end(), read.begin(), read.end());
        if (use_quals)   // output_quality, qualities.cpp:177-204
            for (int l = 0; l < L1; l++) soa.qual1.push_back((uint8_t)((read[l] == 'N' ? s1.offset : (unsigned char)qual[l]) - s1.offset));
        if (r2) {
            soa.seq2.insert(soa.seq2.end(), read2.begin(), read2.end());
            if (use_quals)
                for (int l = 0; l < L2; l++) soa.qual2.push_back((uint8_t)((read2[l] == 'N' ? s2.offset : (unsigned char)qual2[l]) - s2.offset));
        }
        n_total++;
        if (++soa.n == batch_reads) submit();
    }
    submit();
    delete r2;

    if (dump) {
        const std::string d = dump;
        write_file(d + "/seq1.bin", all.seq1.data(), all.seq1.size());
        write_file(d + "/qual1.bin", all.qual1.data(), all.qual1.size());
        write_file(d + "/names.bin", all.names.data(), all.names.size());
        write_file(d + "/name_off.bin", all.name_off.data(), all.name_off.size() * 8);
        write_file(d + "/seq2.bin", all.seq2.data(), all.seq2.size());
        write_file(d + "/qual2.bin", all.qual2.data(), all.qual2.size());
        char meta[256];
        snprintf(meta, sizeof meta, "n %lld\nL1 %d\nL2 %d\noffset1 %d\noffset2 %d\n", (long long)all.n, L1, L2, s1.offset, in2 ? s2.offset : 0);
        write_file(d + "/meta.txt", meta, strlen(meta));
        return 0;
    }

    // dump_trie, compress.cpp:524-552: files 0 names, 1 reads, 2 qualities, 3 meta, 4 reads of mate 2, 5 qualities of mate 2
    scb_result res;
    scb_check(scb_flush(h, &res));
    const int nf = 4 + 2 * (in2 ? 1 : 0);
    std::vector<uint8_t> buf;
    char path[4096];
    for (int c = 0; out && c < res.n_chunks; c++)
        for (int k = 0; k < nf; k++) {
            const int64_t bytes = res.chunk_off[k][c + 1] - res.chunk_off[k][c];
            buf.resize(bytes > 0 ? (size_t)bytes : 1);
            scb_check(scb_copy_stream(h, k, c, buf.data(), bytes));
            snprintf(path, sizeof path, "%s/t_%03d_%d.tmp", out, c, k);
            write_file(path, buf.data(), (size_t)bytes);
        }
    if (merged && out)   // what merge() leaves (compress.cpp:68-198): one set of streams in bucket order
        for (int k = 0; k < nf; k++) {
            const int64_t bytes = res.merged_size[k];
            buf.resize(bytes > 0 ? (size_t)bytes : 1);
            scb_check(scb_copy_stream(h, k, -1, buf.data(), bytes));
            snprintf(path, sizeof path, "%s/merged_%d.tmp", out, k);
            write_file(path, buf.data(), (size_t)bytes);
        }
    if (container) {
        std::vector<uint8_t> st[6];
        for (int k = 0; k < nf; k++) {
            st[k].resize(res.merged_size[k] > 0 ? (size_t)res.merged_size[k] : 0);
            scb_check(scb_copy_stream(h, k, -1, st[k].empty() ? (void *)&st[k] : (void *)st[k].data(), res.merged_size[k]));
        }
        write_containers(container, st, [&](int32_t c) { const char *cs = scb_core(h, c); if (!cs) die("core index out of range"); return (int)strlen(cs); },
                         L1, L2, in2 != nullptr, use_names != 0, s1.offset, library);
    }
    fprintf(stderr, "scb_boost: %lld reads, L %d%s, phred offset %d, %d flush chunk(s), %lld unbucketed, device %.3f ms\n", (long long)n_total, L1,
            in2 ? (" + " + std::to_string(L2)).c_str() : "", s1.offset, res.n_chunks, (long long)scb_unbucketed(h), res.device_ms);
    scb_destroy(h);
    return 0;
}
