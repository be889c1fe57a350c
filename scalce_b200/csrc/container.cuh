// container.cuh - the two neighbours of the transform that SURVEY.md 8(f) ranks next, on the device:
//   (f3) bucket-record assembly of the .scalcer body from the merged meta records + stream 1
//        (combine_and_compress_with_split, compress.cpp:345-384: per non-empty bucket `int32 core, int64 n_reads`,
//        then that bucket's packed reads + end markers), so that the host writer gets ONE contiguous buffer;
//   (f4) the decompress-side inverse (decompress.cpp:331-352): un-rotate + unpack every read around its core,
//        restore 'N' where the quality is 0, add the phred offset back - bucket-ordered streams -> FASTQ rows.
// Both work from a per-bucket segment table (core index, reads) - what the meta records / the container's inline
// headers carry - and are pure streaming kernels.
#pragma once
#include "common.cuh"
#include "prims.cuh"

namespace scb {

// one meta record (reads.cpp:160-176): int32 id, int32 core, int64 tN, tR, tQ [, tR2, tQ2]
struct MetaRecs {
    const uint8_t *p; int rsz;
    __device__ __forceinline__ int32_t core(int64_t s) const { int32_t v; memcpy(&v, p + s * rsz + 4, 4); return v; }
    __device__ __forceinline__ int64_t tR(int64_t s) const { int64_t v; memcpy(&v, p + s * rsz + 16, 8); return v; }
};

// per segment: core index, core length, record size, reads (tR / record size)
__global__ void ct_segments_k(MetaRecs m, int64_t nseg, const int32_t *__restrict__ core_len, int L, int sz_meta, int32_t *__restrict__ seg_core,
                              int64_t *__restrict__ seg_reads, int64_t *__restrict__ seg_bytes) {
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nseg) return;
    const int32_t c = m.core(s);
    const int lv = c == SCB_ROOT_ID_DEV ? 0 : core_len[c];
    const int64_t rec = sz_read(L - lv) + sz_meta, bytes = m.tR(s);
    seg_core[s] = c;
    seg_reads[s] = bytes / rec;          // compress.cpp:371-376
    seg_bytes[s] = bytes;
}

struct SegBytes { const int64_t *v; __device__ __forceinline__ uint64_t operator()(int64_t i) const { return (uint64_t)v[i]; } };

// last s in [0, n) with start[s] <= x (start ascending, start[0] = 0)
__device__ __forceinline__ int64_t ct_find(const uint64_t *__restrict__ start, int64_t n, uint64_t x) {
    int64_t lo = 0, hi = n - 1;
    while (lo < hi) {
        const int64_t mid = (lo + hi + 1) >> 1;
        if (start[mid] <= x) lo = mid; else hi = mid - 1;
    }
    return lo;
}

// (f3) body[out] : for segment s, 12 header bytes at in_start[s] + 12 s, then the segment's bytes of stream 1.
// One thread per 16 output bytes; the segment is found once and advanced linearly.
__global__ void __launch_bounds__(256) ct_assemble_k(const uint8_t *__restrict__ stream1, const uint64_t *__restrict__ in_start /* [nseg + 1] */, int64_t nseg,
                                                     const int32_t *__restrict__ seg_core, const int64_t *__restrict__ seg_reads,
                                                     uint8_t *__restrict__ body, int64_t body_bytes) {
    const int64_t o0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 16;
    if (o0 >= body_bytes) return;
    // output start of segment s = in_start[s] + 12 s: find the last s with that <= o0 (monotone in s)
    int64_t lo = 0, hi = nseg - 1;
    while (lo < hi) {
        const int64_t mid = (lo + hi + 1) >> 1;
        if ((int64_t)in_start[mid] + 12 * mid <= o0) lo = mid; else hi = mid - 1;
    }
    int64_t s = lo;
    int64_t seg_out = (int64_t)in_start[s] + 12 * s, seg_end = (int64_t)in_start[s + 1] + 12 * (s + 1);
    uint8_t hdr[12];
    { const int32_t c = seg_core[s]; const int64_t r = seg_reads[s]; memcpy(hdr, &c, 4); memcpy(hdr + 4, &r, 8); }
    __align__(16) uint8_t out[16];
    int nb = 0;
    for (int k = 0; k < 16; k++) {
        const int64_t o = o0 + k;
        if (o >= body_bytes) break;
        while (o >= seg_end) {           // empty segments cannot occur (records exist for non-empty buckets only), but be safe
            s++;
            seg_out = seg_end; seg_end = (int64_t)in_start[s + 1] + 12 * (s + 1);
            const int32_t c = seg_core[s]; const int64_t r = seg_reads[s]; memcpy(hdr, &c, 4); memcpy(hdr + 4, &r, 8);
        }
        const int64_t within = o - seg_out;
        out[k] = within < 12 ? hdr[within] : stream1[(int64_t)in_start[s] + within - 12];
        nb++;
    }
    if (nb == 16) *(uint4 *)(body + o0) = *(const uint4 *)out;      // body is 16-byte aligned
    else for (int k = 0; k < nb; k++) body[o0 + k] = out[k];
}

// (f4) one thread per (read, 16-base chunk of the output row). Rotated read r = s[end..L) ++ s[0..end-lv) (reads.cpp:432-461):
//   out[i] = r[L - end + i]        for i <  end - lv       (decompress.cpp:337-339)
//          = core[i - (end - lv)]   for end - lv <= i < end (340-341)
//          = r[i - end]             for i >= end            (344-345); end = 0: out = r (no core, root bucket)
// then 'N' where the quality byte is 0 and quality + phred offset (348-352).
struct InvParams {
    const uint8_t *stream1;          // packed reads + end markers, bucket order, no inline headers
    const uint64_t *in_start;        // [nseg + 1] byte start of each segment in stream1
    const uint64_t *read_start;      // [nseg + 1] first read of each segment
    const int32_t *seg_core;
    const int32_t *core_len; const uint64_t *core_off; const uint8_t *core_chars;   // the core set by core index
    int64_t nseg, n;
    int L, sz_meta;
    const uint8_t *quals;            // [n][L] bucket order, or null
    int phred;
    uint8_t *seq_out, *qual_out;     // [n][L]
};

__global__ void __launch_bounds__(256) ct_inverse_k(InvParams p) {
    const int chunks = (p.L + 15) >> 4;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t j = t / chunks;
    if (j >= p.n) return;
    const int i0 = (int)(t - j * chunks) * 16;
    const int64_t s = ct_find(p.read_start, p.nseg, (uint64_t)j);
    const int32_t c = p.seg_core[s];
    const int lv = c == SCB_ROOT_ID_DEV ? 0 : p.core_len[c];
    const int nbytes = sz_read(p.L - lv);
    const uint8_t *rec = p.stream1 + p.in_start[s] + (uint64_t)(j - (int64_t)p.read_start[s]) * (uint64_t)(nbytes + p.sz_meta);
    int end = rec[nbytes];
    if (p.sz_meta == 2) end |= (int)rec[nbytes + 1] << 8;
    const uint8_t *core = c == SCB_ROOT_ID_DEV ? nullptr : p.core_chars + p.core_off[c];
    const uint8_t *q = p.quals ? p.quals + j * (int64_t)p.L : nullptr;
    uint8_t *so = p.seq_out + j * (int64_t)p.L, *qo = p.qual_out ? p.qual_out + j * (int64_t)p.L : nullptr;
    for (int k = 0; k < 16; k++) {
        const int i = i0 + k;
        if (i >= p.L) break;
        uint8_t ch;
        if (end != 0 && i >= end - lv && i < end) ch = core[i - (end - lv)];
        else {
            const int src = end == 0 ? i : (i < end - lv ? p.L - end + i : i - end);
            ch = (uint8_t)"ACGT"[(rec[src >> 2] >> ((~src & 3) << 1)) & 3];
        }
        if (q) {
            const uint8_t qv = q[i];
            if (qv == 0) ch = 'N';
            if (qo) qo[i] = (uint8_t)(qv + p.phred);
        }
        so[i] = ch;
    }
}

// mate 2: stored unrotated, no end marker (compress.cpp:696, decompress.cpp:344-345 with end = 0)
__global__ void __launch_bounds__(256) ct_inverse2_k(const uint8_t *__restrict__ stream4, int64_t n, int L2, const uint8_t *__restrict__ quals, int phred,
                                                     uint8_t *__restrict__ seq_out, uint8_t *__restrict__ qual_out) {
    const int chunks = (L2 + 15) >> 4;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t j = t / chunks;
    if (j >= n) return;
    const int i0 = (int)(t - j * chunks) * 16;
    const uint8_t *rec = stream4 + j * (int64_t)sz_read(L2);
    for (int k = 0; k < 16; k++) {
        const int i = i0 + k;
        if (i >= L2) break;
        uint8_t ch = (uint8_t)"ACGT"[(rec[i >> 2] >> ((~i & 3) << 1)) & 3];
        if (quals) {
            const uint8_t qv = quals[j * (int64_t)L2 + i];
            if (qv == 0) ch = 'N';
            if (qual_out) qual_out[j * (int64_t)L2 + i] = (uint8_t)(qv + phred);
        }
        seq_out[j * (int64_t)L2 + i] = ch;
    }
}

}  // namespace scb
