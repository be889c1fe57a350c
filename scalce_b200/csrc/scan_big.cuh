// scan_big.cuh - core scan for automata that do not fit shared memory (core sets of production size: the reference
// sizes patterns[] for 5-10 M cores, reads.cpp:336, 385; 1 M cores of 10-14 bases are 2.9 M states, 46 MB of transitions).
//
// The completed DFA (reads.cpp:298-315) stays in global memory and is served by L2 (measured: 88-90 % L2 sector hit rate,
// 0.1 % in L1 - on random reads the walk sits at depth >= 9 almost always, so every base is one random 4-byte read of a
// table far larger than L1). The kernel is therefore a latency machine: what counts is how many independent lookups
// are in flight per SM and how few instructions surround each of them.
//   pack_reads16_k  ASCII rows -> 2-bit packed rows (SWAR, scan_smem.cuh's pack4), one thread per 16-base word
//   scan_big_k      one lane per read, 32 warps per SM, nothing staged: the lane reads its own packed words (10 per
//                   150 bp read) and walks. Every table entry carries, next to the next state, the LEVEL of the longest
//                   core ending there (entry = state | level << 26), so the walk keeps only hits of the running maximum
//                   level in its per-lane queue (a higher level restarts the queue) - aho_search (reads.cpp:413-429)
//                   never looks at anything else. Distinct cores among the kept hits are the read's candidates.
// The first version of this kernel queued EVERY hit and sorted levels out afterwards: with 17 % of the positions of a
// random read ending some core of a 1 M set that was 25 hits per read, 4100 instructions per read, half the lanes idle
// (ncu: 16.4 active threads per instruction), a slow path for 8 % of the reads and a second launch because the candidate
// space was reserved by hit count (profiles/r02_ncu_big_kernels.txt). L2 eviction-policy hints on the table loads
// changed nothing (profiles/r02_scan_big_hints.txt).
#pragma once
#include "common.cuh"
#include "pipeline.cuh"
#include "scan_smem.cuh"

namespace scb {

constexpr uint32_t kBigStateBits = 26;         // entry = next state | level << 26
constexpr uint32_t kBigStateMask = (1u << kBigStateBits) - 1u;

// ASCII -> packed rows; seq 4-byte aligned, n * L bytes readable
__global__ void __launch_bounds__(256) pack_reads16_k(const uint8_t *__restrict__ seq, int64_t n, int L, int PW, uint32_t *__restrict__ packed) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * PW) return;
    const int64_t i = t / PW;
    const int k = (int)(t - i * PW);
    const int64_t b = i * (int64_t)L + 16 * k, total = n * (int64_t)L;
    const int nv = L - 16 * k < 16 ? L - 16 * k : 16;
    uint32_t wv;
    if (b + 20 <= total) {
        const uint32_t *a = (const uint32_t *)(seq + (b & ~(int64_t)3));
        const uint32_t sh = (uint32_t)(b & 3) * 8u;
        const uint32_t x0 = a[0], x1 = a[1], x2 = a[2], x3 = a[3], x4 = a[4];
        const uint32_t y0 = __funnelshift_r(x0, x1, sh), y1 = __funnelshift_r(x1, x2, sh), y2 = __funnelshift_r(x2, x3, sh), y3 = __funnelshift_r(x3, x4, sh);
        uint32_t bad = 0;
        wv = (pack4(y0, bad) << 24) | (pack4(y1, bad) << 16) | (pack4(y2, bad) << 8) | pack4(y3, bad);
        if (bad) wv = (pack4_masked(y0) << 24) | (pack4_masked(y1) << 16) | (pack4_masked(y2) << 8) | pack4_masked(y3);
    } else {
        wv = 0;
        for (int j = 0; j < 16; j++) wv = (wv << 2) | (j < nv ? base_code(seq[b + j]) : 0u);
    }
    if (nv < 16) wv &= ~(0xffffffffu >> (2 * nv));
    packed[t] = wv;
}

struct ScanBigParams {
    const uint32_t *packed; int64_t n; int L; int PW;
    const uint32_t *trans;        // [ns * 4] next state | level of the longest core ending there << 26
    const uint32_t *hit_info;     // [n_hit]  rank | level << 24, indexed by state - H0 (states with an output come last)
    uint32_t H0;
    uint8_t *lvl; uint16_t *ncand; uint64_t *cand_off; uint32_t *cand_rank; uint16_t *cand_pos;
    unsigned long long *cand_total; uint64_t cand_cap;
};

constexpr int kBigThreads = 512;
__host__ __device__ inline size_t scan_big_smem_bytes(int threads, int Q) { return (size_t)threads * Q * 6; }

// kBigQ = queued max-level hits per read: 32 -> 2 CTAs (32 warps) per SM, 16 -> 4 CTAs (64 warps) per SM
template <int kBigQ>
__global__ void __launch_bounds__(kBigThreads, kBigQ <= 16 ? 4 : 2) scan_big_k(ScanBigParams p) {
    extern __shared__ __align__(16) uint8_t sm[];
    // per lane: kBigQ states (u32) and kBigQ positions (u16); the lanes of a warp are interleaved slot-wise, so slot a of
    // the 32 lanes falls into 32 different banks
    const int w = threadIdx.x >> 5, l = lane_id();
    const int nw = blockDim.x >> 5;
    uint32_t *qs = (uint32_t *)sm + (size_t)w * (32 * kBigQ) + l;                                  // slot a at qs[a * 32]
    uint16_t *qp = (uint16_t *)((uint32_t *)sm + (size_t)nw * (32 * kBigQ)) + (size_t)w * (32 * kBigQ) + l;   // slot a at qp[a * 32]
    const uint32_t *__restrict__ tr = p.trans;
    const int L = p.L, PW = p.PW;
    uint64_t c_cur = 0, c_end = 0;                    // this warp's slice of the candidate arrays
    const int64_t n_tiles = cdiv(p.n, 32);
    const int64_t stride = (int64_t)gridDim.x * nw;
    for (int64_t tile = (int64_t)blockIdx.x * nw + w; tile < n_tiles; tile += stride) {
        const int64_t i = tile * 32 + l;
        const bool live = i < p.n;
        uint32_t best = 0, qn = 0;
        if (live) {
            const uint32_t *row = p.packed + i * (int64_t)PW;
            uint32_t st = 0;
            for (int k = 0; k < PW; k++) {
                const uint32_t wv = __ldg(row + k);
                const int nbase = L - 16 * k < 16 ? L - 16 * k : 16;
#pragma unroll
                for (int j = 0; j < 16; j++) {
                    if (j >= nbase) break;
                    const uint32_t c = (wv >> (30 - 2 * j)) & 3u;
                    const uint32_t e = __ldg(tr + ((size_t)st * 4u + c));
                    st = e & kBigStateMask;
                    const uint32_t lv = e >> kBigStateBits;
                    if (lv >= best && lv != 0) {                  // a hit of the running maximum level (or a new maximum)
                        if (lv > best) { best = lv; qn = 0; }
                        if (qn < (uint32_t)kBigQ) { qs[qn * 32] = st; qp[qn * 32] = (uint16_t)(16 * k + j); }
                        qn++;
                    }
                }
            }
        }
        // candidate space by the number of max-level hits (an upper bound of the distinct ones)
        const uint32_t v = live ? qn : 0u;
        const uint32_t inc = warp_incl_scan(v);
        const uint32_t wtot = __shfl_sync(0xffffffffu, inc, 31);
        if (c_cur + wtot > c_end) {
            unsigned long long take = wtot > (uint32_t)kCandChunk ? wtot : (uint32_t)kCandChunk, got = 0;
            if (l == 0) got = atomicAdd(p.cand_total, take);
            got = __shfl_sync(0xffffffffu, got, 0);
            c_cur = got; c_end = got + take;
        }
        const uint64_t o = c_cur + (inc - v);
        c_cur += wtot;
        if (live) {
            const bool room = o + (uint64_t)v <= p.cand_cap;
            uint32_t cnt = 0;
            if (qn <= (uint32_t)kBigQ) {
                // distinct cores among the queued hits, in order of first occurrence; the slot is overwritten by the rank
                for (uint32_t a = 0; a < qn; a++) {
                    const uint32_t r = __ldg(p.hit_info + (qs[a * 32] - p.H0)) & 0x00ffffffu;
                    bool dup = false;
                    for (uint32_t b2 = 0; b2 < a && !dup; b2++) dup = qs[b2 * 32] == r;
                    qs[a * 32] = dup ? 0xffffffffu : r;
                    if (!dup) {
                        if (room) { p.cand_rank[o + cnt] = r; p.cand_pos[o + cnt] = qp[a * 32]; }
                        cnt++;
                    }
                }
            } else {
                // more max-level hits than the queue holds (degenerate sets): walk again keeping only hits of the (now known)
                // level, deduped against the candidates written so far. Without room the host reruns with the exact size.
                uint32_t st = 0;
                const uint32_t *row = p.packed + i * (int64_t)PW;
                for (int q = 0; q < L; q++) {
                    const uint32_t e = __ldg(tr + ((size_t)st * 4u + pk_code(row, q)));
                    st = e & kBigStateMask;
                    if ((e >> kBigStateBits) == best) {
                        const uint32_t r = __ldg(p.hit_info + (st - p.H0)) & 0x00ffffffu;
                        bool dup = false;
                        if (room) for (uint32_t b2 = 0; b2 < cnt && !dup; b2++) dup = p.cand_rank[o + b2] == r;
                        if (!dup) {
                            if (room) { p.cand_rank[o + cnt] = r; p.cand_pos[o + cnt] = (uint16_t)q; }
                            cnt++;
                        }
                    }
                }
            }
            p.lvl[i] = (uint8_t)best;
            p.ncand[i] = (uint16_t)cnt;
            p.cand_off[i] = o;
        }
        __syncwarp();
    }
}

}  // namespace scb
