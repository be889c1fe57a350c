/*
 * scalce_oracle.c - CPU restatement of SCALCE's boosting transform.
 *
 * TEST INFRASTRUCTURE ONLY. Nothing under scalce_b200/ (the product) may link, import or
 * execute this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg do,
 * and there only as the checker.
 *
 * Parity pin: this restatement is checked byte-for-byte against the UNMODIFIED reference CLI
 * (oracle/_ref/scalce, built by oracle/Makefile from /root/reference) run at -T 1, and against
 * the committed fixtures under tests/golden/ that the same CLI produced
 * (tests/make_golden.py). The reference has no tests or golden vectors of its own.
 *
 * Each function cites the reference lines it follows. Written from the behaviour, with index
 * arrays instead of pointer-linked nodes; not a copy.
 */
#define _POSIX_C_SOURCE 200809L
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define ORC_MAXBIN (1 << 30) /* const.h:94 */
#define ORC_BIN_NODE_BYTES 40 /* sizeof(bin_node) on LP64, reads.h:52-65; enters compress.cpp:702 */

/* ASCII -> 2 bit. const.cpp:47-49 + const.h:127: table of 58 entries indexed c-'A';
 * A,a,N and everything else -> 0, C/c -> 1, G/g -> 2, T/t -> 3. Outside ['A','A'+58) the
 * reference indexes out of bounds (undefined); we define those as 0. */
static inline int orc_getval(unsigned char c) {
    switch (c) {
    case 'C': case 'c': return 1;
    case 'G': case 'g': return 2;
    case 'T': case 't': return 3;
    default: return 0;
    }
}

typedef struct {
    int32_t child[4];  /* trie children, later the completed DFA (reads.cpp:298-315) */
    int32_t fail;
    int32_t nto;       /* next_to_output, -1 = null */
    int32_t level;
    int32_t id;        /* BFS order id, root 0 (reads.cpp:296) */
    int32_t output;    /* core index or -1 */
    uint64_t bin_size; /* lifetime count, never reset (reads.h:82, reads.cpp:246) */
    /* per-flush bin: indices into the chunk's record array, in arrival order */
    int32_t *bin; int32_t bin_n, bin_cap;
} onode;

typedef struct { uint8_t *p; size_t n, cap; } obuf;

typedef struct {
    int64_t read;   /* global input index */
    int16_t end;    /* rd.end, compress.cpp:682/685 */
} orec;

typedef struct orc {
    onode *nd; int32_t n_nodes, cap_nodes; /* node 0 = root */
    char **cores; int32_t n_cores;
    int L1, L2, use_names, paired, use_quals;
    uint64_t bucket_set_bytes;
    uint64_t total_size;      /* compress.cpp:596,702 */
    int64_t n_reads;
    int unbucketed;           /* reads.cpp:218,492 */
    /* current (unflushed) chunk */
    orec *rec; int64_t n_rec, cap_rec;
    /* retained inputs (concatenated copies so submit() may be called repeatedly) */
    obuf seq1, qual1, seq2, qual2, names; int64_t *name_off; int64_t cap_off;
    /* flushed chunks: 6 streams each (0 names,1 reads,2 quals,3 meta,4 reads2,5 quals2) */
    obuf (*chunk)[6]; int n_chunks, cap_chunks;
    int as_chunks;            /* flush chunks closed so far by orc_assign (no streams behind them) */
    /* merged result (compress.cpp:488-522) */
    obuf merged[6]; int merged_valid;
    /* per-read debug */
    int32_t *dbg_node_id, *dbg_core, *dbg_end, *dbg_chunk; int64_t cap_dbg;
} orc;

static void die(const char *m) { fprintf(stderr, "(ORACLE ERROR) %s\n", m); abort(); }
static void *xrealloc(void *p, size_t n) { void *q = realloc(p, n ? n : 1); if (!q) die("oom"); return q; }
static void ob_put(obuf *b, const void *src, size_t n) {
    if (b->n + n > b->cap) { size_t c = b->cap ? b->cap * 2 : 4096; while (c < b->n + n) c *= 2; b->p = xrealloc(b->p, c); b->cap = c; }
    memcpy(b->p + b->n, src, n); b->n += n;
}

/* reads.cpp:221-230 */
static int32_t node_new(orc *o, int level) {
    if (o->n_nodes == o->cap_nodes) { o->cap_nodes = o->cap_nodes ? o->cap_nodes * 2 : 1024; o->nd = xrealloc(o->nd, sizeof(onode) * (size_t)o->cap_nodes); }
    onode *t = &o->nd[o->n_nodes];
    memset(t, 0, sizeof *t);
    t->child[0] = t->child[1] = t->child[2] = t->child[3] = -1;
    t->fail = -1; t->nto = -1; t->output = -1; t->level = level; t->id = 0;
    return o->n_nodes++;
}

/* reads.cpp:253-267 (iterative; a later duplicate core overwrites output: "last index wins") */
static void pattern_insert(orc *o, const char *c, int id) {
    int32_t n = 0; int level = 0;
    for (; *c != 0 && *c != '\n'; c++) {
        int cx = orc_getval((unsigned char)*c);
        if (o->nd[n].child[cx] < 0) { int32_t k = node_new(o, level + 1); o->nd[n].child[cx] = k; }
        n = o->nd[n].child[cx]; level++;
    }
    o->nd[n].output = id;
}

/* reads.cpp:270-315 */
static void prepare_automaton(orc *o) {
    onode *nd = o->nd; int32_t *q = xrealloc(NULL, sizeof(int32_t) * (size_t)(o->n_nodes + 1));
    int64_t qs = 0, qe = 0;
    nd[0].fail = 0;
    for (int i = 0; i < 4; i++) if (nd[0].child[i] >= 0) { nd[nd[0].child[i]].fail = 0; q[qe++] = nd[0].child[i]; }
    int traversed = 0;
    while (qs < qe) {                              /* BFS #1: fail links, next_to_output, ids */
        int32_t cur = q[qs++];
        for (int i = 0; i < 4; i++) {
            int32_t t = nd[cur].child[i];
            if (t >= 0) {
                int32_t f = nd[cur].fail;
                while (f != 0 && nd[f].child[i] < 0) f = nd[f].fail;
                nd[t].fail = nd[f].child[i] >= 0 ? nd[f].child[i] : 0;
                /* root's own children: f==0 and child[i]==t itself only when cur==root, which is
                 * not the case here since root is never dequeued in BFS #1 */
                q[qe++] = t;
                nd[t].nto = nd[nd[t].fail].output >= 0 ? nd[t].fail : nd[nd[t].fail].nto;
            }
        }
        nd[cur].id = ++traversed;
    }
    qs = qe = 0; q[qe++] = 0;
    while (qs < qe) {                              /* BFS #2: complete the DFA; nto := self if output */
        int32_t cur = q[qs++];
        for (int i = 0; i < 4; i++) {
            if (nd[cur].child[i] >= 0) q[qe++] = nd[cur].child[i];
            int32_t c = cur;
            while (c != 0 && nd[c].child[i] < 0) c = nd[c].fail;
            nd[cur].child[i] = nd[c].child[i] >= 0 ? nd[c].child[i] : 0;
            c = cur;
            while (c >= 0 && nd[c].output == -1) c = nd[c].nto;
            nd[cur].nto = c;
        }
    }
    free(q);
}

orc *orc_create(const char *const *cores, int32_t n_cores, int L1, int L2, int use_names,
                int paired, int use_quals, uint64_t bucket_set_bytes) {
    orc *o = calloc(1, sizeof *o);
    o->L1 = L1; o->L2 = L2; o->use_names = use_names; o->paired = paired; o->use_quals = use_quals;
    o->bucket_set_bytes = bucket_set_bytes;
    node_new(o, 0);
    o->cores = xrealloc(NULL, sizeof(char *) * (size_t)(n_cores + 1)); o->n_cores = n_cores;
    for (int i = 0; i < n_cores; i++) { o->cores[i] = strdup(cores[i]); pattern_insert(o, cores[i], i); } /* reads.cpp:389-394 */
    prepare_automaton(o);
    return o;
}

/* reads.cpp:413-429. text has L chars (the '\n' terminator is the length bound here). */
static int aho_search(orc *o, const uint8_t *text, int L, int32_t *bucket) {
    onode *nd = o->nd; int32_t cur = 0, largest = -1; int bestpos = -1;
    for (int i = 0; i < L; i++) {
        cur = nd[cur].child[orc_getval(text[i])];
        int32_t x = nd[cur].nto;
        if (x >= 0) {
            if (largest < 0 || nd[largest].level < nd[x].level ||
                (nd[largest].level == nd[x].level && nd[largest].bin_size < nd[x].bin_size)) { bestpos = i; largest = x; }
        }
    }
    *bucket = largest >= 0 ? largest : 0;
    return bestpos;
}

/* reads.cpp:432-461: bases [n+l, L) then [0, n), 4 per byte MSB first, zero padded. */
static int output_read(const uint8_t *line, int L, uint8_t *dest, int n, int l) {
    int bc = 0, cc = 0; uint8_t ca = 0;
    for (int i = n + l; i < L; i++) { ca = (uint8_t)((ca << 2) | orc_getval(line[i])); if (++cc == 4) { dest[bc++] = ca; cc = 0; } }
    for (int i = 0; i < n; i++)     { ca = (uint8_t)((ca << 2) | orc_getval(line[i])); if (++cc == 4) { dest[bc++] = ca; cc = 0; } }
    if (cc) { while (cc != 4) { ca <<= 2; cc++; } dest[bc++] = ca; }
    return bc;
}

#define SZ_READ(l) (((l) / 4) + ((l) % 4 > 0)) /* const.h:63 */

static int name_len(orc *o, int64_t r) { return (int)(o->name_off[r + 1] - o->name_off[r]); }

/* ---- in-bucket sort: reads.cpp:547-634 -------------------------------------------------- */
typedef struct { orc *o; uint8_t *packed; int pk_stride; int limit; int32_t *nodes, *temp; } sortctx;
/* _POS (reads.cpp:557-558): digit i of the packed rotated read if i + end < L else 0 */
static inline int s_pos(sortctx *s, int32_t rec, int i) {
    if (i + s->o->rec[rec].end >= s->o->L1) return 0;
    return (s->packed[(size_t)rec * s->pk_stride + i / 4] >> ((3 - i % 4) * 2)) & 3;
}
static void radix_sort(sortctx *s, int pos, int64_t start, int64_t size) { /* reads.cpp:563-598 */
    while (1) {
        if (size <= 1 || pos >= s->limit) return;
        int64_t count[4] = {0, 0, 0, 0}, cum[5];
        for (int64_t i = start; i < start + size; i++) { count[s_pos(s, s->nodes[i], pos)]++; s->temp[i] = s->nodes[i]; }
        cum[0] = 0; for (int k = 1; k < 5; k++) cum[k] = cum[k - 1] + count[k - 1];
        int64_t w[4] = {cum[0], cum[1], cum[2], cum[3]};
        for (int64_t i = start; i < start + size; i++) { int c = s_pos(s, s->temp[i], pos); s->nodes[start + w[c]++] = s->temp[i]; }
        /* recurse into partitions; the last non-trivial one is iterated to bound stack depth */
        int last = -1; for (int k = 0; k < 4; k++) if (count[k] > 1) last = k;
        for (int k = 0; k < 4; k++) if (k != last && count[k] > 1) radix_sort(s, pos + 1, start + cum[k], count[k]);
        if (last < 0) return;
        start += cum[last]; size = count[last]; pos++;
    }
}

/* ---- flush: aho_output reads.cpp:466-499, bin_prepare 600-634, bin_dump 91-180 ----------- */
static void bin_dump(orc *o, int32_t t, obuf *f, sortctx *s, uint8_t *packed2buf) {
    onode *nd = &o->nd[t];
    int lenCore = nd->output >= 0 ? (int)strlen(o->cores[nd->output]) : 0;
    int sz_read = SZ_READ(o->L1 - lenCore);
    int sz_meta = o->L1 > 255 ? 2 : 1;
    int64_t tN = 0, tQ = 0, tR = 0, tR2 = 0, tQ2 = 0;
    for (int32_t k = 0; k < nd->bin_n; k++) {
        int32_t rc = s->nodes[k]; int64_t r = o->rec[rc].read;
        if (o->use_names) {
            int nl = name_len(o, r); uint8_t b = (uint8_t)nl;
            ob_put(&f[0], &b, 1); ob_put(&f[0], o->names.p + o->name_off[r], (size_t)nl); tN += nl + 1;
        }
        ob_put(&f[1], s->packed + (size_t)rc * s->pk_stride, (size_t)sz_read);
        ob_put(&f[1], &o->rec[rc].end, (size_t)sz_meta);      /* little-endian low bytes of int16 end */
        tR += sz_read + sz_meta;
        if (o->use_quals) { ob_put(&f[2], o->qual1.p + (size_t)r * o->L1, (size_t)o->L1); tQ += o->L1; }
        if (o->paired) {
            int n2 = output_read(o->seq2.p + (size_t)r * o->L2, o->L2, packed2buf, 0, 0);   /* compress.cpp:696 */
            ob_put(&f[4], packed2buf, (size_t)n2); tR2 += n2;
            if (o->use_quals) { ob_put(&f[5], o->qual2.p + (size_t)r * o->L2, (size_t)o->L2); tQ2 += o->L2; }
        }
    }
    int32_t a, b;
    if (nd->output == -1) { a = b = ORC_MAXBIN - 1; } else { a = nd->id; b = nd->output; }
    ob_put(&f[3], &a, 4); ob_put(&f[3], &b, 4);
    ob_put(&f[3], &tN, 8); ob_put(&f[3], &tR, 8); ob_put(&f[3], &tQ, 8);
    if (o->paired) { ob_put(&f[3], &tR2, 8); ob_put(&f[3], &tQ2, 8); }
    nd->bin_n = 0;                                             /* bin_free, reads.cpp:84-85 */
}

static void bin_prepare_and_dump(orc *o, int32_t t, obuf *f, sortctx *s, uint8_t *p2) {
    onode *nd = &o->nd[t];
    s->nodes = xrealloc(s->nodes, sizeof(int32_t) * (size_t)nd->bin_n);
    s->temp = xrealloc(s->temp, sizeof(int32_t) * (size_t)nd->bin_n);
    memcpy(s->nodes, nd->bin, sizeof(int32_t) * (size_t)nd->bin_n);
    s->limit = o->L1 - nd->level;                              /* reads.cpp:625 */
    radix_sort(s, 0, 0, nd->bin_n);
    bin_dump(o, t, f, s, p2);
}

static void flush_chunk(orc *o) {
    if (o->n_chunks == o->cap_chunks) { o->cap_chunks = o->cap_chunks ? o->cap_chunks * 2 : 8; o->chunk = xrealloc(o->chunk, sizeof(obuf[6]) * (size_t)o->cap_chunks); }
    obuf *f = o->chunk[o->n_chunks++]; memset(f, 0, sizeof(obuf[6]));
    /* packed rotated reads of this chunk (the payload the reference keeps in its pool) */
    sortctx s; memset(&s, 0, sizeof s); s.o = o; s.pk_stride = SZ_READ(o->L1) + 1;
    s.packed = xrealloc(NULL, (size_t)s.pk_stride * (size_t)(o->n_rec + 1));
    uint8_t *p2 = xrealloc(NULL, (size_t)SZ_READ(o->L2 > 0 ? o->L2 : 1) + 8);
    for (int64_t k = 0; k < o->n_rec; k++) {
        int64_t r = o->rec[k].read; int lvl = o->dbg_core[r] >= 0 ? (int)strlen(o->cores[o->dbg_core[r]]) : 0;
        int end = o->rec[k].end;
        /* compress.cpp:679-686: output_read(read, dest, n-level+1, level) with n = end-1; or (0,0) */
        output_read(o->seq1.p + (size_t)r * o->L1, o->L1, s.packed + (size_t)k * s.pk_stride, end ? end - lvl : 0, end ? lvl : 0);
    }
    /* BFS over the completed DFA from the root's four children (reads.cpp:466-490). If a base has
     * no core starting with it, root->child[i] is the root itself and the root bucket is emitted
     * at that point of the traversal instead of last - reproduced as is. */
    char *visited = calloc((size_t)o->n_nodes + 1, 1); int32_t *q = xrealloc(NULL, sizeof(int32_t) * ((size_t)o->n_nodes + 5));
    int64_t qs = 0, qe = 0; onode *nd = o->nd;
    visited[nd[0].id] = 1;
    for (int i = 0; i < 4; i++) { q[qe++] = nd[0].child[i]; visited[nd[nd[0].child[i]].id] = 1; }
    /* NB: visited is indexed by id; ids are unique per node (root 0) so this equals a per-node flag */
    while (qs < qe) {
        int32_t cur = q[qs++];
        if (nd[cur].bin_n) bin_prepare_and_dump(o, cur, f, &s, p2);
        for (int i = 0; i < 4; i++) { int32_t c = nd[cur].child[i]; if (!visited[nd[c].id]) { q[qe++] = c; visited[nd[c].id] = 1; } }
    }
    if (nd[0].bin_n) { o->unbucketed += nd[0].bin_n; bin_prepare_and_dump(o, 0, f, &s, p2); }
    free(visited); free(q); free(s.packed); free(s.nodes); free(s.temp); free(p2);
    o->n_rec = 0;
}

/* Per-read driver, compress.cpp:673-715 at -T 1. Inputs are the SoA the host parser builds:
 * seq ASCII [n][L]; qual = bytes already produced by output_quality (qualities.cpp:177-204);
 * names = chars after '@' up to the first space (names.cpp:48-62), offsets name_off[n+1]. */
void orc_submit(orc *o, int64_t n, const uint8_t *seq1, const uint8_t *qual1, const uint8_t *names,
                const int64_t *name_off, const uint8_t *seq2, const uint8_t *qual2) {
    int64_t base = o->n_reads;
    ob_put(&o->seq1, seq1, (size_t)n * o->L1);
    if (o->use_quals) ob_put(&o->qual1, qual1, (size_t)n * o->L1);
    if (o->paired) { ob_put(&o->seq2, seq2, (size_t)n * o->L2); if (o->use_quals) ob_put(&o->qual2, qual2, (size_t)n * o->L2); }
    if (base + n + 1 > o->cap_off) { o->cap_off = (base + n + 1) * 2; o->name_off = xrealloc(o->name_off, sizeof(int64_t) * (size_t)o->cap_off); }
    if (base == 0) o->name_off[0] = 0;
    if (o->use_names) {
        ob_put(&o->names, names + name_off[0], (size_t)(name_off[n] - name_off[0]));
        for (int64_t i = 0; i < n; i++) o->name_off[base + i + 1] = o->name_off[base] + (name_off[i + 1] - name_off[0]);
    } else for (int64_t i = 0; i < n; i++) o->name_off[base + i + 1] = 0;
    if (base + n > o->cap_dbg) {
        o->cap_dbg = (base + n) * 2;
        o->dbg_node_id = xrealloc(o->dbg_node_id, 4 * (size_t)o->cap_dbg); o->dbg_core = xrealloc(o->dbg_core, 4 * (size_t)o->cap_dbg);
        o->dbg_end = xrealloc(o->dbg_end, 4 * (size_t)o->cap_dbg); o->dbg_chunk = xrealloc(o->dbg_chunk, 4 * (size_t)o->cap_dbg);
    }
    for (int64_t i = 0; i < n; i++) {
        int64_t r = base + i; int32_t bucket;
        int bp = aho_search(o, o->seq1.p + (size_t)r * o->L1, o->L1, &bucket);     /* compress.cpp:673 */
        onode *b = &o->nd[bucket];
        int32_t sz = o->use_names ? name_len(o, r) + 1 : 1;                       /* output_name, names.cpp:48-62 */
        int end;
        if (bp != -1) { sz += SZ_READ(o->L1 - b->level); end = bp + 1; } else { sz += SZ_READ(o->L1); end = 0; }
        if (o->use_quals) sz += o->L1;
        if (o->paired) { sz += SZ_READ(o->L2); if (o->use_quals) sz += o->L2; }
        /* aho_trie_bucket, reads.cpp:233-250 */
        if (o->n_rec == o->cap_rec) { o->cap_rec = o->cap_rec ? o->cap_rec * 2 : 4096; o->rec = xrealloc(o->rec, sizeof(orec) * (size_t)o->cap_rec); }
        o->rec[o->n_rec].read = r; o->rec[o->n_rec].end = (int16_t)end;
        if (b->bin_n == b->bin_cap) { b->bin_cap = b->bin_cap ? b->bin_cap * 2 : 4; b->bin = xrealloc(b->bin, 4 * (size_t)b->bin_cap); }
        b->bin[b->bin_n++] = (int32_t)o->n_rec; o->n_rec++;
        b->bin_size++;
        o->dbg_node_id[r] = bucket == 0 ? ORC_MAXBIN - 1 : b->id; o->dbg_core[r] = b->output; o->dbg_end[r] = end; o->dbg_chunk[r] = o->n_chunks;
        o->total_size += (uint64_t)sz + ORC_BIN_NODE_BYTES;                       /* compress.cpp:702 */
        if (o->total_size >= o->bucket_set_bytes) { flush_chunk(o); o->total_size = 0; } /* compress.cpp:708-713 */
    }
    o->n_reads += n;
}

/* The same per-read driver reduced to what DECIDES bucket, end marker and flush chunk (aho_search reads.cpp:413-429, the
 * rd.sz accounting and flush trigger compress.cpp:675-715, bin_size++ reads.cpp:246): no payload is kept and nothing is
 * emitted, so inputs of any size can be streamed through in pieces (tests/test_gpu_fullsize.py: all 50 M reads of the bench
 * workload). name_off[n+1] gives the name lengths (ignored without names). Use a fresh oracle: do not mix with orc_submit. */
void orc_assign(orc *o, int64_t n, const uint8_t *seq1, const int64_t *name_off, int32_t *node_id, int32_t *core, int32_t *end_out,
                int32_t *chunk) {
    for (int64_t i = 0; i < n; i++) {
        int32_t bucket;
        int bp = aho_search(o, seq1 + (size_t)i * o->L1, o->L1, &bucket);
        onode *b = &o->nd[bucket];
        int32_t sz = o->use_names ? (int32_t)(name_off[i + 1] - name_off[i]) + 1 : 1;
        int end;
        if (bp != -1) { sz += SZ_READ(o->L1 - b->level); end = bp + 1; } else { sz += SZ_READ(o->L1); end = 0; }
        if (o->use_quals) sz += o->L1;
        if (o->paired) { sz += SZ_READ(o->L2); if (o->use_quals) sz += o->L2; }
        b->bin_size++;
        if (bucket == 0) o->unbucketed++;
        node_id[i] = bucket == 0 ? ORC_MAXBIN - 1 : b->id; core[i] = b->output; end_out[i] = end; chunk[i] = o->as_chunks;
        o->total_size += (uint64_t)sz + ORC_BIN_NODE_BYTES;
        if (o->total_size >= o->bucket_set_bytes) { o->as_chunks++; o->total_size = 0; }
    }
    o->n_reads += n;
}

void orc_finish(orc *o) { if (o->total_size) { flush_chunk(o); o->total_size = 0; } } /* compress.cpp:799-801 */

/* Temp-file merge, compress.cpp:68-198 + 488-522. For one chunk the reference skips the merge
 * (compress.cpp:807-808) and the single chunk's files are the result. */
void orc_merge(orc *o) {
    for (int k = 0; k < 6; k++) o->merged[k].n = 0;
    o->merged_valid = 1;
    int flc = o->n_chunks;
    if (flc == 1) { for (int k = 0; k < 6; k++) ob_put(&o->merged[k], o->chunk[0][k].p, o->chunk[0][k].n); return; }
    if (flc == 0) return;
    int nlen = 3 + 2 * o->paired; size_t rsz = 8 + 8 * (size_t)nlen;
    int nf = 4 + 2 * o->paired; size_t mm = 0;
    uint8_t *metadata = calloc(1, 1); size_t meta_cap = 1;
    for (int idx = 0; idx < nf; idx++) {
        if (idx == 3) continue;
        size_t *mpos = calloc((size_t)flc, sizeof(size_t)), *dpos = calloc((size_t)flc, sizeof(size_t));
        int32_t *bins = malloc(4 * (size_t)flc), *cores = malloc(4 * (size_t)flc);
        int minV = ORC_MAXBIN, minC = ORC_MAXBIN, minI = 0;
        for (int i = 0; i < flc; i++) {
            obuf *m = &o->chunk[i][3];
            memcpy(&bins[i], m->p + mpos[i], 4); memcpy(&cores[i], m->p + mpos[i] + 4, 4); mpos[i] += 8;
            if (bins[i] < minV) { minV = bins[i]; minC = cores[i]; minI = i; }
        }
        int binex = flc; size_t metadata_pos = 0; int64_t totalLen = 0; int END = idx - (idx > 3);
        while (binex) {
            obuf *m = &o->chunk[minI][3]; int64_t len;
            memcpy(&len, m->p + mpos[minI] + 8 * (size_t)END, 8);     /* the END-th int64 of the record */
            mpos[minI] += 8 * (size_t)nlen;
            ob_put(&o->merged[idx], o->chunk[minI][idx].p + dpos[minI], (size_t)len); dpos[minI] += (size_t)len;
            totalLen += len;
            int prevMinV = minV, prevMinC = minC; minV = ORC_MAXBIN;
            if (mpos[minI] + 4 > m->n) { bins[minI] = ORC_MAXBIN; binex--; }
            else { memcpy(&bins[minI], m->p + mpos[minI], 4); memcpy(&cores[minI], m->p + mpos[minI] + 4, 4); mpos[minI] += 8; }
            for (int i = 0; i < flc; i++) if (bins[i] < minV) { minV = bins[i]; minC = cores[i]; minI = i; }
            if (minV != prevMinV) {
                if (metadata_pos + rsz > meta_cap) { size_t c = meta_cap * 2 + rsz; metadata = xrealloc(metadata, c); memset(metadata + meta_cap, 0, c - meta_cap); meta_cap = c; }
                memcpy(metadata + metadata_pos, &prevMinV, 4); memcpy(metadata + metadata_pos + 4, &prevMinC, 4);
                memcpy(metadata + metadata_pos + 8 + 8 * (size_t)END, &totalLen, 8);
                metadata_pos += rsz; totalLen = 0;
            }
        }
        mm = metadata_pos;
        free(mpos); free(dpos); free(bins); free(cores);
    }
    ob_put(&o->merged[3], metadata, mm); free(metadata);
}

/* ---- accessors -------------------------------------------------------------------------- */
int orc_n_chunks(orc *o) { return o->n_chunks; }
int64_t orc_stream_size(orc *o, int chunk, int k) { return chunk < 0 ? (int64_t)o->merged[k].n : (int64_t)o->chunk[chunk][k].n; }
const uint8_t *orc_stream_data(orc *o, int chunk, int k) { return chunk < 0 ? o->merged[k].p : o->chunk[chunk][k].p; }
int orc_unbucketed(orc *o) { return o->unbucketed; }
int32_t orc_n_nodes(orc *o) { return o->n_nodes - 1; }
void orc_debug(orc *o, int32_t *node_id, int32_t *core, int32_t *end, int32_t *chunk) {
    size_t b = 4 * (size_t)o->n_reads;
    if (node_id) memcpy(node_id, o->dbg_node_id, b);
    if (core) memcpy(core, o->dbg_core, b);
    if (end) memcpy(end, o->dbg_end, b);
    if (chunk) memcpy(chunk, o->dbg_chunk, b);
}
/* lifetime count of a core (bin_size of its node), by core index; -1 -> root */
uint64_t orc_lifetime_count(orc *o, int32_t core) {
    if (core < 0) return o->nd[0].bin_size;
    for (int32_t i = 1; i < o->n_nodes; i++) if (o->nd[i].output == core) return o->nd[i].bin_size;
    return 0;
}

/* BFS node id (reads.cpp:296) of the node a core ends at - the "id" written to meta. */
int32_t orc_core_node_id(orc *o, int32_t core) {
    int32_t n = 0;
    for (const char *c = o->cores[core]; *c != 0 && *c != '\n'; c++) {
        /* the DFA is complete by now; walking a core from the root follows its own trie path */
        n = o->nd[n].child[orc_getval((unsigned char)*c)];
    }
    return o->nd[n].id;
}

/* Helper for design experiments and scan-kernel unit parity (derived from reads.cpp:413-429, not a
 * reference function): the count-independent part of aho_search - the maximum level reached and
 * the distinct nodes of that level in order of first occurrence, with that first position.
 * Writes up to cap entries; returns the number of distinct candidates. */
int orc_candidates(orc *o, const uint8_t *text, int L, int cap, int32_t *cand_core, int32_t *cand_pos, int32_t *level) {
    onode *nd = o->nd; int32_t cur = 0; int best = 0, n = 0; int32_t tmp_node[4096];
    for (int i = 0; i < L; i++) {
        cur = nd[cur].child[orc_getval(text[i])];
        int32_t x = nd[cur].nto; if (x < 0) continue;
        if (nd[x].level > best) { best = nd[x].level; n = 0; }
        if (nd[x].level == best) {
            int dup = 0; for (int k = 0; k < n && k < 4096; k++) if (tmp_node[k] == x) { dup = 1; break; }
            if (!dup) { if (n < 4096) tmp_node[n] = x; if (n < cap) { cand_core[n] = nd[x].output; cand_pos[n] = i; } n++; }
        }
    }
    *level = best; return n;
}

/* ---- decompress-side inverse, decompress.cpp:331-352 (mate 1) -------------------------------------
 * stream: packed reads + end markers in bucket order WITHOUT the inline bucket headers; seg_core / seg_reads: what the
 * headers (decompress.cpp:262-272) carry. Rebuilds every read around its core, restores 'N' where the quality byte is 0
 * and adds the phred offset back. Returns the number of reads written (rows of L1 bytes, no newline). */
int64_t orc_inverse(orc *o, const uint8_t *stream, const int32_t *seg_core, const int64_t *seg_reads, int64_t nseg,
                    const uint8_t *quals, int phred, uint8_t *seq_out, uint8_t *qual_out) {
    static const char alphabet[] = "ACGT";
    const int L = o->L1, sz_meta = L > 255 ? 2 : 1;
    int64_t K = 0; size_t pos = 0;
    for (int64_t s = 0; s < nseg; s++) {
        int32_t core = seg_core[s];
        int corlen = (core == ORC_MAXBIN - 1) ? 0 : (int)strlen(o->cores[core]);      /* decompress.cpp:268 */
        for (int64_t r = 0; r < seg_reads[s]; r++, K++) {
            const uint8_t *p = stream + pos; pos += (size_t)SZ_READ(L - corlen);          /* :332 */
            int64_t end = 0; memcpy(&end, stream + pos, (size_t)sz_meta); pos += (size_t)sz_meta;   /* :335 */
            uint8_t *l = seq_out + (size_t)K * L; int lc = 0;
            if (end) {
                for (int i = L - (int)end; i < L - corlen; i++) l[lc++] = (uint8_t)alphabet[p[i >> 2] >> ((~i & 3) << 1) & 3];   /* :337-339 */
                for (int i = 0; i < corlen; i++) l[lc++] = (uint8_t)o->cores[core][i];                                         /* :340-341 */
            }
            for (int i = 0; i < L - (int)end; i++) l[lc++] = (uint8_t)alphabet[p[i >> 2] >> ((~i & 3) << 1) & 3];              /* :344-345 */
            if (quals)
                for (int i = 0; i < L; i++) {                                                                                   /* :348-352 */
                    uint8_t q = quals[(size_t)K * L + i];
                    if (!q) l[i] = 'N';
                    if (qual_out) qual_out[(size_t)K * L + i] = (uint8_t)(q + phred);
                }
        }
    }
    return K;
}

/* ---- the host front end: parse loop compress.cpp:614-671, output_name names.cpp:48-62, output_quality
 * qualities.cpp:177-204 at lossy percentage 0 (values[c] = c) with its input-order statistics. text: plain FASTQ, whole
 * records. prev[2] = the two payload symbols before this text (500, 500 at the start of a job, qualities.cpp:179);
 * freq3 [80*80] / freq4 [80*80*80] are updated in place (NULL: statistics off, as with -A). names: characters after '@' up to
 * the first space; name_off[n+1]. Returns the number of records, -1 on a malformed record. */
int64_t orc_parse_fastq(const uint8_t *text, int64_t bytes, int L, int phred, uint8_t *seq, uint8_t *qual, uint8_t *names,
                        int64_t *name_off, uint64_t *freq3, uint64_t *freq4, uint32_t *prev) {
    int64_t pos = 0, n = 0, nb = 0;
    if (name_off) name_off[0] = 0;
    while (pos < bytes) {
        int64_t ls[4], le[4];
        for (int k = 0; k < 4; k++) {                      /* f_gets x4, compress.cpp:615-642 */
            if (pos > bytes) return -1;
            ls[k] = pos;
            while (pos < bytes && text[pos] != '\n') pos++;
            le[k] = pos;
            pos++;
            if (k < 3 && le[k] >= bytes) return -1;
        }
        if (text[ls[0]] != '@' || le[1] - ls[1] != L || le[3] - ls[3] != L) return -1;
        int64_t q = ls[0] + 1;
        while (q < le[0] && text[q] != ' ') q++;           /* names.cpp:55 */
        if (q - (ls[0] + 1) > 255) return -1;
        if (names) { memcpy(names + nb, text + ls[0] + 1, (size_t)(q - (ls[0] + 1))); }
        nb += q - (ls[0] + 1);
        if (name_off) name_off[n + 1] = nb;
        memcpy(seq + (size_t)n * L, text + ls[1], (size_t)L);
        if (qual)
            for (int i = 0; i < L; i++) {                   /* qualities.cpp:181-202 */
                uint8_t d = (uint8_t)((text[ls[1] + i] == 'N' ? phred : text[ls[3] + i]) - phred);
                qual[(size_t)n * L + i] = d;
                if (freq3 && freq4) {
                    if (prev[1] < 256) {
                        freq3[prev[1] * 80 + d]++;
                        if (prev[0] < 256) freq4[(prev[0] * 80 + prev[1]) * 80 + d]++;
                    } else {
                        for (int e = 0; e < 80 * 80 * 80; e++) freq4[e] = 1;
                    }
                    prev[0] = prev[1]; prev[1] = d;
                }
            }
        n++;
    }
    return n;
}

void orc_destroy(orc *o) {
    if (!o) return;
    for (int32_t i = 0; i < o->n_nodes; i++) free(o->nd[i].bin);
    for (int i = 0; i < o->n_cores; i++) free(o->cores[i]);
    for (int c = 0; c < o->n_chunks; c++) for (int k = 0; k < 6; k++) free(o->chunk[c][k].p);
    for (int k = 0; k < 6; k++) free(o->merged[k].p);
    free(o->nd); free(o->cores); free(o->rec); free(o->chunk); free(o->name_off);
    free(o->seq1.p); free(o->qual1.p); free(o->seq2.p); free(o->qual2.p); free(o->names.p);
    free(o->dbg_node_id); free(o->dbg_core); free(o->dbg_end); free(o->dbg_chunk); free(o);
}
