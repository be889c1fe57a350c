// parse.cuh - SURVEY.md 8(f2): the host front end of the transform on the device.
// FASTQ text (4 lines per record, fixed read length) -> the structure of arrays scb_submit takes:
//   parse loop            compress.cpp:614-671  (line boundaries, '@' name line, read / quality lines of L characters)
//   output_name           names.cpp:48-62       (characters after '@' up to the first space or the end of the line)
//   output_quality        qualities.cpp:177-204 at lossy 0: payload = quality - offset, 0 under an upper-case 'N';
//                         and its INPUT-ORDER context statistics: ac_freq3[prev1][cur], ac_freq4[prev0][prev1][cur] over the
//                         running stream of payload symbols (prev carried across reads, files and calls)
// Kernels: newline count per tile -> exclusive scan -> line ends; one thread per record for the name length and the
// checks; one warp per record for the copies; a histogram pass with per-CTA shared-memory privatisation.
#pragma once
#include "common.cuh"
#include "prims.cuh"

namespace scb {

constexpr int kParseTile = 4096;          // bytes per thread block of the newline passes (256 threads x 16 bytes)
constexpr int kAcDepth = 80;              // AC_DEPTH, arithmetic.h:47

struct NlCount {   // newlines in tile t (for the exclusive scan)
    const uint32_t *c;
    __device__ __forceinline__ uint64_t operator()(int64_t t) const { return (uint64_t)c[t]; }
};

__global__ void __launch_bounds__(256) parse_count_nl_k(const uint8_t *__restrict__ text, int64_t bytes, uint32_t *__restrict__ tile_cnt) {
    __shared__ uint32_t sm[8];
    const int64_t b0 = (int64_t)blockIdx.x * kParseTile + (int64_t)threadIdx.x * 16;
    uint32_t c = 0;
    for (int k = 0; k < 16; k++) c += (b0 + k < bytes && text[b0 + k] == '\n') ? 1u : 0u;
    c = __reduce_add_sync(0xffffffffu, c);
    if (lane_id() == 0) sm[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) { uint32_t t = 0; for (int w = 0; w < 8; w++) t += sm[w]; tile_cnt[blockIdx.x] = t; }
}

// line_end[k] = byte position of the k-th newline
__global__ void __launch_bounds__(256) parse_line_ends_k(const uint8_t *__restrict__ text, int64_t bytes, const uint64_t *__restrict__ tile_first,
                                                         int64_t *__restrict__ line_end, int64_t cap) {
    __shared__ uint32_t sm[9];
    const int64_t b0 = (int64_t)blockIdx.x * kParseTile + (int64_t)threadIdx.x * 16;
    uint32_t m = 0;
    for (int k = 0; k < 16; k++) m |= (b0 + k < bytes && text[b0 + k] == '\n') ? (1u << k) : 0u;
    const uint32_t c = __popc(m);
    const uint32_t ex = block_excl_scan<uint32_t, 256>(c, sm, (uint32_t *)nullptr);
    int64_t k = (int64_t)tile_first[blockIdx.x] + ex;
    while (m) {
        const int j = __ffs(m) - 1;
        m &= m - 1;
        if (k < cap) line_end[k] = b0 + j;
        k++;
    }
}

// per record: the four lines, the name length, the checks. err: bit 0 name line does not start with '@', bit 1 read line
// length != L, bit 2 quality line length != L, bit 3 name longer than 255, bit 4 quality symbol outside [offset, offset + 80)
struct ParseRec {
    const uint8_t *text; int64_t bytes; const int64_t *line_end; int64_t n; int L;
    uint32_t *name_len; uint32_t *err;
};
__device__ __forceinline__ int64_t parse_line_start(const ParseRec &p, int64_t line) { return line == 0 ? 0 : p.line_end[line - 1] + 1; }
__global__ void __launch_bounds__(256) parse_records_k(ParseRec p) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= p.n) return;
    const int64_t s0 = parse_line_start(p, 4 * r), e0 = p.line_end[4 * r];
    const int64_t s1 = e0 + 1, e1 = p.line_end[4 * r + 1];
    const int64_t s3 = p.line_end[4 * r + 2] + 1, e3 = p.line_end[4 * r + 3];
    uint32_t err = 0;
    if (p.text[s0] != '@') err |= 1u;
    if (e1 - s1 != p.L) err |= 2u;
    if (e3 - s3 != p.L) err |= 4u;
    int64_t q = s0 + 1;
    while (q < e0 && p.text[q] != ' ') q++;              // names.cpp:55: up to the first space or the newline
    const int64_t nl = q - (s0 + 1);
    if (nl > 255) err |= 8u;
    p.name_len[r] = (uint32_t)(nl > 255 ? 255 : nl);
    if (err) atomicOr(p.err, err);
}

struct NameLenU { const uint32_t *v; __device__ __forceinline__ uint64_t operator()(int64_t i) const { return (uint64_t)v[i]; } };

// one warp per record: read row, quality payload row, name bytes
struct ParseCopy {
    const uint8_t *text; const int64_t *line_end; int64_t n; int L; int phred;
    const uint64_t *name_off;          // [n + 1]
    uint8_t *seq, *qual, *names;       // qual / names may be null
    uint32_t *err;
};
__global__ void __launch_bounds__(256) parse_copy_k(ParseCopy p) {
    const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (r >= p.n) return;
    const int l = lane_id();
    const int64_t s0 = (r == 0 ? 0 : p.line_end[4 * r - 1] + 1), s1 = p.line_end[4 * r] + 1, s3 = p.line_end[4 * r + 2] + 1;
    const uint8_t *rd = p.text + s1, *ql = p.text + s3;
    uint32_t bad = 0;
    for (int i = l; i < p.L; i += 32) {
        const uint8_t b = rd[i];
        p.seq[r * (int64_t)p.L + i] = b;
        if (p.qual) {
            const int v = (int)ql[i] - p.phred;                         // qualities.cpp:183 at lossy 0 (values[c] = c)
            if (v < 0 || v >= kAcDepth) bad = 16u;
            p.qual[r * (int64_t)p.L + i] = b == 'N' ? (uint8_t)0 : (uint8_t)v;
        }
    }
    if (p.names) {
        const uint64_t o = p.name_off[r];
        const int nl = (int)(p.name_off[r + 1] - o);
        for (int i = l; i < nl; i += 32) p.names[o + i] = p.text[s0 + 1 + i];
    }
    if (bad) atomicOr(p.err, bad);
}

// ---- ac_freq3 / ac_freq4 over the stream of payload symbols, qualities.cpp:186-199 -------------------------------------------
// sym[t], t in [0, total): the rows of this call one after the other. The two symbols before sym[0] are prev0 / prev1 (>= 256:
// there is none yet, as the reference's initial 500). Each thread walks kStatRun consecutive symbols; freq3 (6400 bins) is
// privatised per CTA in shared memory, freq4 (512000 bins) goes through a direct-mapped shared-memory cache of (bin, count)
// slots - quality streams concentrate on few contexts, and one global atomic per symbol on the same few addresses would
// serialise - that is flushed with global atomics at the end; a slot conflict goes to global memory directly.
constexpr int kStatRun = 64;
constexpr int kStatSlots = 2048;    // 25.6 KB (freq3) + 16 KB stay under the 48 KB of static shared memory
__global__ void __launch_bounds__(256) parse_qstats_k(const uint8_t *__restrict__ sym, int64_t total, uint32_t prev0, uint32_t prev1,
                                                      unsigned long long *__restrict__ freq3, unsigned long long *__restrict__ freq4) {
    __shared__ uint32_t s3[kAcDepth * kAcDepth];
    __shared__ uint32_t s_tag[kStatSlots], s_cnt[kStatSlots];
    for (int i = threadIdx.x; i < kAcDepth * kAcDepth; i += blockDim.x) s3[i] = 0;
    for (int i = threadIdx.x; i < kStatSlots; i += blockDim.x) { s_tag[i] = 0xffffffffu; s_cnt[i] = 0; }
    __syncthreads();
    const int64_t nruns = (total + kStatRun - 1) / kStatRun;
    for (int64_t run = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; run < nruns; run += (int64_t)gridDim.x * blockDim.x) {
        const int64_t t0 = run * kStatRun, t1 = t0 + kStatRun < total ? t0 + kStatRun : total;
        uint32_t p0 = t0 >= 2 ? sym[t0 - 2] : (t0 == 1 ? prev1 : prev0);
        uint32_t p1 = t0 >= 1 ? sym[t0 - 1] : prev1;
        uint32_t last_bin = 0xffffffffu, last_cnt = 0;
        for (int64_t t = t0; t < t1; t++) {
            const uint32_t c = sym[t];
            if (p1 < 256) {
                atomicAdd(&s3[p1 * kAcDepth + c], 1u);
                if (p0 < 256) {
                    const uint32_t bin = (p0 * kAcDepth + p1) * kAcDepth + c;
                    if (bin == last_bin) last_cnt++;
                    else {
                        if (last_cnt) {
                            const uint32_t slot = (last_bin * 2654435761u) >> 21;          // 11 bits
                            const uint32_t old = atomicCAS(&s_tag[slot], 0xffffffffu, last_bin);
                            if (old == 0xffffffffu || old == last_bin) atomicAdd(&s_cnt[slot], last_cnt);
                            else atomicAdd(&freq4[last_bin], (unsigned long long)last_cnt);
                        }
                        last_bin = bin; last_cnt = 1;
                    }
                }
            }
            p0 = p1; p1 = c;
        }
        if (last_cnt) {
            const uint32_t slot = (last_bin * 2654435761u) >> 21;
            const uint32_t old = atomicCAS(&s_tag[slot], 0xffffffffu, last_bin);
            if (old == 0xffffffffu || old == last_bin) atomicAdd(&s_cnt[slot], last_cnt);
            else atomicAdd(&freq4[last_bin], (unsigned long long)last_cnt);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kAcDepth * kAcDepth; i += blockDim.x) if (s3[i]) atomicAdd(&freq3[i], (unsigned long long)s3[i]);
    for (int i = threadIdx.x; i < kStatSlots; i += blockDim.x) if (s_cnt[i]) atomicAdd(&freq4[s_tag[i]], (unsigned long long)s_cnt[i]);
}

}  // namespace scb
