// sim_resolve.c - CPU model of the tie-break's fixed-point SCHEDULES (not of the kernels): how many rounds do the
// geometric-block schedule of one GPU and the joint rounds of the sharded run need, and what do estimate-driven
// pre-rounds change? Used to choose what to measure on the GPU; nothing here is part of the product or of a test.
//
// Rule (reads.cpp:420-425, 246): read i goes to the FIRST of its candidates with the largest population so far.
// Model of the headline workload (50M x 150 bp, 2048 cores: 1024 x 8 bp, 512 x 9, 256 x 10, 128 x 11, 128 x 12): per
// read the number of hits of each core length is Poisson (positions x cores / 4^len), the read's candidates are that
// many distinct uniformly random cores of the longest length that hit.
// A round = every subtile (1776 per rank, as 148 CTAs x 12 warps) is swept sequentially from start counts derived
// from the previous round's per-subtile histograms (Jacobi across subtiles) - DESIGN.md section 5.
//
//   gcc -O2 -fopenmp -o /tmp/sim_resolve tools/sim_resolve.c -lm
//   /tmp/sim_resolve single 50000000            rounds of the one-GPU geometric schedule
//   /tmp/sim_resolve shard 8 50000000 0         joint rounds at 8 ranks, no pre-rounds
//   /tmp/sim_resolve shard 8 50000000 3         ... with 3 pre-rounds on ranks >= 1
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define NB 2048
#define MAXC 6
// bookkeeping of the incremental rounds (resolve_dense.cuh "fragile reads"): a subtile whose start counts moved by Dmax since its last
// FULL sweep and whose decisions differ from that sweep in E reads must be swept in full again when 2 * (Dmax + E) > T; otherwise a
// replay of its fragile list suffices. The model's sweeps are always exact - this only COUNTS which subtiles the kernel would re-sweep.
static uint32_t *g_S0 = NULL; static uint8_t *g_sel0 = NULL; static int g_round_in_block = 0; static int64_t g_redo = 0, g_redo_reads = 0; static int g_T = 1024;
static double g_model_us = 0;
static int SUB = 1776;   // subtiles per block (148 CTAs x 12 warps); SIM_SUB overrides

static uint64_t rs;
static inline uint64_t rnd(uint64_t *s) { *s ^= *s << 13; *s ^= *s >> 7; *s ^= *s << 17; return *s; }
static inline double urnd(uint64_t *s) { return (rnd(s) >> 11) * (1.0 / 9007199254740992.0); }
static int poisson(uint64_t *s, double lam) { double L = exp(-lam), p = 1; int k = 0; do { k++; p *= urnd(s); } while (p > L); return k - 1; }

typedef struct { int64_t n; uint8_t *nc; uint16_t *c; uint8_t *sel; } Reads;

static void gen(Reads *R, int64_t n, uint64_t seed) {
    R->n = n; R->nc = malloc(n); R->c = malloc((size_t)n * MAXC * 2); R->sel = malloc(n);
    memset(R->sel, 0xff, n);
    const int L = 150, len[5] = {8, 9, 10, 11, 12}, cnt[5] = {1024, 512, 256, 128, 128}, first[5] = {0, 1024, 1536, 1792, 1920};
    double lam[5];
    for (int k = 0; k < 5; k++) lam[k] = (double)(L - len[k] + 1) * cnt[k] / pow(4.0, len[k]);
#pragma omp parallel
    {
        uint64_t s = seed * 0x9E3779B97F4A7C15ull + 12345;
#pragma omp for schedule(static)
        for (int64_t i = 0; i < n; i++) {
            if ((i & 0xffff) == 0) s = (seed + (uint64_t)i) * 0x9E3779B97F4A7C15ull + 777;
            int h[5], top = -1;
            for (int k = 0; k < 5; k++) { h[k] = poisson(&s, lam[k]); if (h[k] > 0) top = k; }
            int m = 0;
            if (top >= 0) {
                int want = h[top] > MAXC ? MAXC : h[top];
                while (m < want) {
                    uint16_t b = (uint16_t)(first[top] + rnd(&s) % cnt[top]);
                    int dup = 0;
                    for (int q = 0; q < m; q++) dup |= R->c[i * MAXC + q] == b;
                    if (!dup) R->c[i * MAXC + m++] = b; else want--;     // a repeated core is one candidate
                }
            }
            R->nc[i] = (uint8_t)m;
        }
    }
}

// SIM_CHEAP_FIRST=1: the guess round decides every read INDEPENDENTLY from its subtile's (extrapolated) start counts - a
// data-parallel pass instead of a sequential sweep; hist[] receives the subtile's histogram under those decisions
static int64_t sweep_static(Reads *R, int64_t lo, int64_t hi, const uint32_t *cnt, uint32_t *hist) {
    int64_t ch = 0;
    for (int64_t i = lo; i < hi; i++) {
        const int nc = R->nc[i];
        if (!nc) continue;
        const uint16_t *c = R->c + i * MAXC;
        int best = 0; uint32_t bc = cnt[c[0]];
        for (int k = 1; k < nc; k++) if (cnt[c[k]] > bc) { bc = cnt[c[k]]; best = k; }
        hist[c[best]]++;
        if (R->sel[i] != best) { R->sel[i] = (uint8_t)best; ch++; }
    }
    return ch;
}

// one sweep of reads [lo, hi) from counts cnt[] (modified); returns the number of decisions that changed
static int64_t sweep(Reads *R, int64_t lo, int64_t hi, uint32_t *cnt) {
    int64_t ch = 0;
    for (int64_t i = lo; i < hi; i++) {
        const int nc = R->nc[i];
        if (!nc) continue;
        const uint16_t *c = R->c + i * MAXC;
        int best = 0; uint32_t bc = cnt[c[0]];
        for (int k = 1; k < nc; k++) if (cnt[c[k]] > bc) { bc = cnt[c[k]]; best = k; }
        cnt[c[best]]++;
        if (R->sel[i] != best) { R->sel[i] = (uint8_t)best; ch++; }
    }
    return ch;
}

// ---- model of the incremental rounds: fragile flags, replay, and the policy for subtiles whose margin bound fails -------------
// g_frag[i] = 1 if read i's decision margin at its subtile's last FULL sweep was <= T (it is on the fragile list).
// SIM_DEFER=1: a subtile whose bound fails is NOT re-swept at once (today: it is, and the whole round waits for it); it keeps
// replaying - which may leave non-fragile reads with stale decisions - and is only swept in full in a later round, when the replay
// rounds have gone quiet. Termination then needs a quiet round with no stale subtile left.
static uint8_t *g_frag = NULL, *g_stale = NULL; static int g_defer = 0, g_verify = 0; static int64_t g_sweeps = 0, g_replays = 0;
static int64_t sweep_margin(Reads *R, int64_t lo, int64_t hi, uint32_t *cnt, int T) {
    int64_t ch = 0;
    for (int64_t i = lo; i < hi; i++) {
        const int nc = R->nc[i];
        g_frag[i] = 0;
        if (!nc) continue;
        const uint16_t *c = R->c + i * MAXC;
        int best = 0; uint32_t bc = cnt[c[0]];
        for (int k = 1; k < nc; k++) if (cnt[c[k]] > bc) { bc = cnt[c[k]]; best = k; }
        int64_t margin = 1 << 30;
        for (int k = 0; k < nc; k++) if (k != best) { int64_t m = (int64_t)bc - cnt[c[k]] - (k < best ? 1 : 0); if (m < margin) margin = m; }
        g_frag[i] = nc >= 2 && margin <= T;
        cnt[c[best]]++;
        if (R->sel[i] != best) { R->sel[i] = (uint8_t)best; ch++; }
    }
    return ch;
}
static int64_t replay(Reads *R, int64_t lo, int64_t hi, uint32_t *cnt) {   // fragile reads re-decided exactly, the others keep their decision
    int64_t ch = 0;
    for (int64_t i = lo; i < hi; i++) {
        const int nc = R->nc[i];
        if (!nc) continue;
        const uint16_t *c = R->c + i * MAXC;
        int best = R->sel[i];
        if (g_frag[i]) {
            best = 0; uint32_t bc = cnt[c[0]];
            for (int k = 1; k < nc; k++) if (cnt[c[k]] > bc) { bc = cnt[c[k]]; best = k; }
            if (R->sel[i] != best) { R->sel[i] = (uint8_t)best; ch++; }
        }
        cnt[c[best]]++;
    }
    return ch;
}

// One round over block [n0, n1) of R split into SUB subtiles. H[t][b]: per-subtile histograms of the previous round
// (updated). first: start counts extrapolated from base (g0 = reads before n0 in the job). Returns changed decisions.
static int64_t round_block(Reads *R, int64_t n0, int64_t n1, const uint32_t *base, uint32_t *H, int first, int64_t g0, uint32_t *tot) {
    const int64_t len = n1 - n0;
    int64_t ts = (len + SUB - 1) / SUB; ts = (ts + 31) / 32 * 32; if (ts < 32) ts = 32;
    const int ns = (int)((len + ts - 1) / ts);
    uint32_t *start = malloc((size_t)ns * NB * 4);
    if (first) {
        for (int t = 0; t < ns; t++)
            for (int b = 0; b < NB; b++)
                start[(size_t)t * NB + b] = base[b] + (g0 + n0 > 0 ? (uint32_t)((uint64_t)base[b] * (uint64_t)(t * ts) / (uint64_t)(g0 + n0)) : 0);
    } else {
        for (int b = 0; b < NB; b++) { uint32_t run = base[b]; for (int t = 0; t < ns; t++) { start[(size_t)t * NB + b] = run; run += H[(size_t)t * NB + b]; } }
    }
    int64_t changed = 0;
    const int cheap_first = getenv("SIM_CHEAP_FIRST") != NULL;
#pragma omp parallel for schedule(dynamic, 8) reduction(+ : changed)
    for (int t = 0; t < ns; t++) {
        uint32_t cnt[NB];
        memcpy(cnt, start + (size_t)t * NB, sizeof cnt);
        const int64_t lo = n0 + t * ts, hi = lo + ts < n1 ? lo + ts : n1;
        if (first && cheap_first) {
            uint32_t hh[NB];
            memset(hh, 0, sizeof hh);
            changed += sweep_static(R, lo, hi, cnt, hh);
            memcpy(H + (size_t)t * NB, hh, sizeof hh);
            continue;
        }
        if (g_frag) {
            int full = g_round_in_block < 2 || (g_verify && g_stale[t]);
            int64_t ch_t = 0;
            if (!full) {
                uint32_t dmax = 0;
                for (int b = 0; b < NB; b++) { int32_t d = (int32_t)(start[(size_t)t * NB + b] - g_S0[(size_t)t * NB + b]); uint32_t a = d < 0 ? -d : d; if (a > dmax) dmax = a; }
                uint32_t save[NB]; memcpy(save, cnt, sizeof save);
                ch_t = replay(R, lo, hi, cnt);
                int64_t E = 0;
                for (int64_t i = lo; i < hi; i++) E += R->sel[i] != g_sel0[i];
                if (2 * ((int64_t)dmax + E) > g_T) {
                    if (g_defer) g_stale[t] = 1;                              // keep the replay's result, sweep later
                    else { memcpy(cnt, save, sizeof save); full = 1; }        // today's kernel: sweep in full in the same round
                } else if (g_defer && !getenv("SIM_STICKY")) g_stale[t] = 0;  // the bound holds again: the replay was exact (as in the kernel)
            }
            if (full) {
                ch_t += sweep_margin(R, lo, hi, cnt, g_T);
                memcpy(g_S0 + (size_t)t * NB, start + (size_t)t * NB, NB * 4); memcpy(g_sel0 + lo, R->sel + lo, hi - lo);
                g_stale[t] = 0;
#pragma omp atomic
                g_sweeps++;
            } else {
#pragma omp atomic
                g_replays++;
            }
            changed += ch_t;
            for (int b = 0; b < NB; b++) H[(size_t)t * NB + b] = cnt[b] - start[(size_t)t * NB + b];
            continue;
        }
        changed += sweep(R, lo, hi, cnt);
        for (int b = 0; b < NB; b++) H[(size_t)t * NB + b] = cnt[b] - start[(size_t)t * NB + b];
        if (g_S0) {
            int full = g_round_in_block < 2;                       // guess round and the round after it always sweep in full
            if (!full) {
                uint32_t dmax = 0;
                for (int b = 0; b < NB; b++) { int32_t d = (int32_t)(start[(size_t)t * NB + b] - g_S0[(size_t)t * NB + b]); uint32_t a = d < 0 ? -d : d; if (a > dmax) dmax = a; }
                int64_t E = 0;
                for (int64_t i = lo; i < hi; i++) E += R->sel[i] != g_sel0[i];
                full = 2 * ((int64_t)dmax + E) > g_T;
                if (full) {
#pragma omp atomic
                    g_redo++;
#pragma omp atomic
                    g_redo_reads += hi - lo;
                }
            }
            if (full) { memcpy(g_S0 + (size_t)t * NB, start + (size_t)t * NB, NB * 4); memcpy(g_sel0 + lo, R->sel + lo, hi - lo); }
        }
    }
    if (tot) { memset(tot, 0, NB * 4); for (int t = 0; t < ns; t++) for (int b = 0; b < NB; b++) tot[b] += H[(size_t)t * NB + b]; }
    free(start);
    return changed;
}

// geometric schedule of one GPU (api.cu dense_blocks): x4 until 1M reads precede, then x2. Returns total rounds.
static int single(Reads *R, uint32_t *base) {
    uint32_t *H = calloc((size_t)SUB * NB, 4), tot[NB];
    int64_t n0 = 0; int rounds = 0, nblk = 0;
    while (n0 < R->n) {
        const char *ge = getenv("SIM_GROWTH_LATE"), *gs = getenv("SIM_GROWTH_EARLY"), *sm = getenv("SIM_SMALL");   // api.cu: 2, 4, 1M
        const int64_t g_late = ge ? atoi(ge) : 2, g_early = gs ? atoi(gs) : 4, small = sm ? atoll(sm) : (1 << 20);
        const int64_t first_len = getenv("SIM_FIRST") ? atoll(getenv("SIM_FIRST")) : 4096;   // api.cu: 4096
        int64_t lenb = n0 < small ? (g_early - 1) * n0 : (g_late - 1) * n0; if (lenb < first_len) lenb = first_len;
        int64_t n1 = n0 + lenb < R->n ? n0 + lenb : R->n;
        int first = 1, r = 0;
        char trace[768]; int tl = 0;
        if (getenv("SIM_REDO")) { if (!g_S0) { g_S0 = calloc((size_t)SUB * NB, 4); g_sel0 = malloc(R->n); if (getenv("SIM_T")) g_T = atoi(getenv("SIM_T")); } }
        g_round_in_block = 0;
        if (getenv("SIM_POLICY")) {   // model the replay / re-sweep machinery: SIM_POLICY=now (today's kernel) or defer
            if (!g_frag) { g_frag = calloc(R->n, 1); g_stale = calloc(SUB + 1, 1); if (!g_S0) { g_S0 = calloc((size_t)SUB * NB, 4); g_sel0 = malloc(R->n); } if (getenv("SIM_T")) g_T = atoi(getenv("SIM_T")); }
            g_defer = !strcmp(getenv("SIM_POLICY"), "defer");
            memset(g_stale, 0, SUB + 1);
            const int64_t per = (n1 - n0 + SUB - 1) / SUB;
            double us = 0;
            for (;;) {
                g_sweeps = g_replays = 0;
                int64_t ch = round_block(R, n0, n1, base, H, first, 0, tot); first = 0; r++;
                // time model: 35 us fixed; a round lasts as long as its slowest subtile: 25 ns per read swept, 100 ns per fragile record replayed
                // (replay cost from the block's actual fragile fraction, 100 ns per record: 32 records per ~3 us replay iteration)
                int64_t nf = 0; for (int64_t i = n0; i < n1; i++) nf += g_frag[i];
                const double frag = (double)nf / (double)(n1 - n0);
                us += 35.0 + (g_sweeps ? per * 0.025 : per * frag * 0.1);
                if (tl < 700) tl += snprintf(trace + tl, sizeof trace - tl, " %ld(%ld)", (long)ch, (long)g_sweeps);
                g_round_in_block++;
                g_verify = 0;
                if (!ch) {
                    int any = 0; for (int t = 0; t < SUB; t++) any |= g_stale[t];
                    if (!any) break;
                    g_verify = 1;                                             // quiet, but stale subtiles remain: sweep them in the next round
                }
            }
            g_model_us += us;
            if (getenv("SIM_TRACE")) printf("    changed(subtiles swept in full) per round:%s   ~%.0f us\n", trace, us);
            for (int b = 0; b < NB; b++) base[b] += tot[b];
            printf("  block %2d [%9ld, %9ld): %d rounds\n", nblk, (long)n0, (long)n1, r);
            rounds += r; nblk++; n0 = n1;
            continue;
        }
        for (;;) {
            g_redo = 0; g_redo_reads = 0;
            int64_t ch = round_block(R, n0, n1, base, H, first, 0, tot); first = 0; r++;
            if (tl < 700) tl += g_S0 ? snprintf(trace + tl, sizeof trace - tl, " %ld(%ld)", (long)ch, (long)g_redo) : snprintf(trace + tl, sizeof trace - tl, " %ld", (long)ch);
            g_round_in_block++;
            if (!ch) break;
        }
        if (getenv("SIM_TRACE")) printf("    changed per round:%s\n", trace);
        for (int b = 0; b < NB; b++) base[b] += tot[b];
        printf("  block %2d [%9ld, %9ld): %d rounds\n", nblk, (long)n0, (long)n1, r);
        rounds += r; nblk++; n0 = n1;
    }
    free(H);
    return rounds;
}

// Overlapped schedule: the next block starts (guess round, then sweeps) as soon as the block before it changes fewer than
// `thr` decisions in a round - its tail rounds and the next block's first rounds share the same global rounds. A block
// retires when it changed nothing in a round in which every earlier block was already final. Returns global rounds.
static int overlapped(Reads *R, uint32_t *base, int64_t thr) {
    enum { MAXB = 64 };
    int64_t b0[MAXB], b1[MAXB]; int nblk = 0;
    for (int64_t n0 = 0; n0 < R->n;) {
        int64_t lenb = n0 < (1 << 20) ? 3 * n0 : n0; if (lenb < 4096) lenb = 4096;
        int64_t n1 = n0 + lenb < R->n ? n0 + lenb : R->n;
        b0[nblk] = n0; b1[nblk] = n1; nblk++; n0 = n1;
    }
    uint32_t *H[MAXB] = {0}; uint32_t (*tot)[NB] = calloc(nblk, sizeof *tot);
    int started[MAXB] = {0}, rounds_of[MAXB] = {0};
    int64_t last_ch[MAXB];
    int lo = 0, hi = 0, rounds = 0;                   // active blocks [lo, hi]
    started[0] = 1; H[0] = calloc((size_t)SUB * NB, 4); last_ch[0] = -1;
    while (lo < nblk) {
        // populations before each active block: final base + current totals of the active blocks before it
        uint32_t run[NB]; memcpy(run, base, sizeof run);
        int64_t ch_now[MAXB];
        for (int k = lo; k <= hi; k++) {
            uint32_t bk[NB]; memcpy(bk, run, sizeof bk);
            uint32_t nt[NB];
            ch_now[k] = round_block(R, b0[k], b1[k], bk, H[k], rounds_of[k] == 0, 0, nt);
            for (int b = 0; b < NB; b++) run[b] += tot[k][b];          // Jacobi: later blocks see LAST round's totals of earlier ones
            memcpy(tot[k], nt, sizeof nt);
            rounds_of[k]++;
        }
        rounds++;
        // retire from the front
        while (lo <= hi && rounds_of[lo] >= 2 && ch_now[lo] == 0 && last_ch[lo] == 0) {   // two quiet rounds in a row: inputs were final
            for (int b = 0; b < NB; b++) base[b] += tot[lo][b];
            free(H[lo]); lo++;
        }
        for (int k = lo; k <= hi; k++) last_ch[k] = ch_now[k];
        if (lo > hi && lo < nblk) { hi = lo; started[hi] = 1; H[hi] = calloc((size_t)SUB * NB, 4); last_ch[hi] = -1; }
        else if (hi + 1 < nblk && hi - lo < 2 && rounds_of[hi] >= 2 && ch_now[hi] < thr) { hi++; started[hi] = 1; H[hi] = calloc((size_t)SUB * NB, 4); last_ch[hi] = -1; }
    }
    free(tot);
    return rounds;
}

int main(int argc, char **argv) {
    if (getenv("SIM_SUB")) SUB = atoi(getenv("SIM_SUB"));
    if (argc >= 4 && !strcmp(argv[1], "overlap")) {
        Reads R; gen(&R, atoll(argv[2]), 1);
        uint32_t base[NB] = {0};
        int r = overlapped(&R, base, atoll(argv[3]));
        // check against the sequential answer
        uint32_t cnt[NB] = {0}; Reads S = R; S.sel = malloc(R.n); memset(S.sel, 0xff, R.n); sweep(&S, 0, R.n, cnt);
        int64_t bad = 0; for (int64_t i = 0; i < R.n; i++) bad += R.nc[i] && R.sel[i] != S.sel[i];
        printf("overlapped schedule (next block starts below %ld changes), %ld reads: %d global rounds; decisions differing from the sequential answer: %ld\n",
               (long)atoll(argv[3]), (long)R.n, r, (long)bad);
        return 0;
    }
    if (argc < 3) { fprintf(stderr, "usage: sim_resolve single N | shard G N_PER_RANK PREROUNDS\n"); return 2; }
    if (!strcmp(argv[1], "single")) {
        Reads R; gen(&R, atoll(argv[2]), 1);
        uint32_t base[NB] = {0};
        int r = single(&R, base);
        printf("single GPU, %ld reads: %d rounds in total\n", (long)R.n, r);
        if (getenv("SIM_POLICY")) {
            uint32_t cnt[NB] = {0}; Reads S = R; S.sel = malloc(R.n); memset(S.sel, 0xff, R.n); sweep(&S, 0, R.n, cnt);
            int64_t bad = 0; for (int64_t i = 0; i < R.n; i++) bad += R.nc[i] && R.sel[i] != S.sel[i];
            printf("policy %s, T = %d: modelled time %.2f ms; decisions differing from the sequential answer: %ld\n", getenv("SIM_POLICY"), g_T, g_model_us * 1e-3, (long)bad);
        }
        return 0;
    }
    const int G = atoi(argv[2]); const int64_t n = atoll(argv[3]); const int pre = argc > 4 ? atoi(argv[4]) : 0;
    Reads *R = malloc(sizeof(Reads) * G);
    for (int g = 0; g < G; g++) gen(&R[g], n, 100 + g);
    uint32_t (*hist)[NB] = calloc(G, sizeof *hist);        // per-rank bucket histogram of the current assignment
    uint32_t **H = malloc(sizeof(void *) * G);
    for (int g = 0; g < G; g++) H[g] = calloc((size_t)SUB * NB, 4);
    // rank 0: exact (what its solo resolve produces)
    { uint32_t cnt[NB] = {0}; sweep(&R[0], 0, n, cnt); memcpy(hist[0], cnt, sizeof cnt); }
    // pre-rounds on ranks >= 1: first from nothing, then from the rank's own histogram scaled to the reads before it
    for (int k = 0; k < pre; k++)
        for (int g = 1; g < G; g++) {
            uint32_t base[NB];
            for (int b = 0; b < NB; b++) base[b] = k == 0 ? 0 : (uint32_t)((uint64_t)hist[g][b] * (uint64_t)g);
            int64_t ch = round_block(&R[g], 0, n, base, H[g], k == 0, (int64_t)g * n, hist[g]);
            if (g == 1 || g == G - 1) printf("  pre-round %d rank %d: %ld changed\n", k, g, (long)ch);
        }
    int rounds = 0, first = pre == 0;
    const int gs = getenv("SIM_GS") != NULL;
    for (;;) {
        uint32_t (*nh)[NB] = calloc(G, sizeof *nh);
        int64_t changed = 0;
        for (int g = 1; g < G; g++) {
            uint32_t base[NB];
            for (int b = 0; b < NB; b++) {
                if (first) base[b] = (uint32_t)((uint64_t)hist[0][b] * (uint64_t)g);           // rank 0's histogram scaled (exact for rank 1)
                else { uint32_t s = 0; for (int q = 0; q < g; q++) s += (gs && q >= 1) ? nh[q][b] : hist[q][b]; base[b] = s; }   // SIM_GS: Gauss-Seidel across ranks
            }
            changed += round_block(&R[g], 0, n, base, H[g], first, (int64_t)g * n, nh[g]);
        }
        for (int g = 1; g < G; g++) memcpy(hist[g], nh[g], sizeof hist[g]);     // Jacobi across ranks: all see last round's histograms
        free(nh);
        rounds++;
        printf("  joint round %2d: %ld decisions changed\n", rounds, (long)changed);
        if (!first && changed == 0) break;
        first = 0;
        if (rounds > 200) break;
    }
    printf("%d ranks x %ld reads, %d pre-rounds: %d joint rounds\n", G, (long)n, pre, rounds);
    return 0;
}
