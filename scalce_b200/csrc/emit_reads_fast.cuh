// emit_reads_fast.cuh - stream 1 (rotated 2-bit reads + end marker) with fewer instructions per read
// (default for reads of more than 32 bases since the end of round 1: 2.3 -> 1.9 ms at 50M x 150 bp; SCB_EMIT_READS_V2=0
// selects emit_reads_st_k).
//
// ncu on emit_reads_st_k (gpurun_out/r01_top.raw.csv): issue bound - 83 % issue active, 44 warp instructions per read.
// SASS: the row staging moves 4 bytes per ~21 instructions (13 items per read), and the record loop spends ~100
// instructions per 4 output bytes, most of them on per-byte conditions. Here:
//   * rows are staged with 8-byte loads (PW even: rows are 8-byte aligned), 5 items per read at 150 bp;
//   * a record is assembled by emit_record (emit_record.h: branch-free words, four unconditional byte stores per full
//     word; checked base by base on the CPU), ~25 instructions per 4 output bytes;
//   * grid-stride over tiles, so the same kernel serves the co-resident mode (emit_coresident.cuh).
// Shared-memory layout and the final flush are those of emit_reads_st_k; the bytes written are identical.
#pragma once
#include "emit2.cuh"
#include "emit_record.h"

namespace scb {

__global__ void __launch_bounds__(256) emit_reads_fast_k(EmitMParams e, int RPB, uint32_t inv_half, int recmax, int64_t n_blk) {
    extern __shared__ __align__(16) uint8_t sbd[];
    // layout: record bytes [RPB*recmax + 48] | rows [RPB][PW + pad] u32 (odd pitch)
    const int PW = e.PW, PWs = (PW + kEmitRowPad) | 1;
    uint32_t *s_rows = (uint32_t *)(sbd + (((size_t)RPB * recmax + 48 + 15) & ~(size_t)15));
    const bool wide = (PW & 1) == 0 && (((uintptr_t)e.packed) & 7) == 0;   // rows 8-byte aligned
    const uint32_t half = (uint32_t)(wide ? PW / 2 : PW);                  // load items per row
    for (int64_t blk = blockIdx.x; blk < n_blk; blk += gridDim.x) {
        const int64_t p0 = blk * RPB, p1 = (p0 + RPB < e.n) ? p0 + RPB : e.n;
        const int np = (int)(p1 - p0);
        const uint64_t g0 = e.offR[p0];
        const int len = (int)(e.offR[p1] - g0);
        uint8_t *sb = sbd + (int)(g0 & 15);
        {   // rows: fetched once per read; the pad words are zeroed by the read's own thread below
            const uint32_t items = (uint32_t)np * half;
            for (uint32_t t = threadIdx.x; t < items; t += 256) {
                const uint32_t pl = half == 1 ? t : __umulhi(t, inv_half), k = t - pl * half;   // ceil(2^32 / 1) does not fit 32 bits
                const uint32_t *src = e.packed + (int64_t)e.perm[p0 + pl] * PW;
                uint32_t *dst = s_rows + (size_t)pl * PWs;
                if (wide) {
                    const int64_t v = ldg_g64((const int64_t *)src + k);
                    dst[2 * k] = (uint32_t)v; dst[2 * k + 1] = (uint32_t)((uint64_t)v >> 32);
                } else {
                    dst[k] = ldg_g64(src + k);
                }
            }
            for (int pl = threadIdx.x; pl < np; pl += 256)
                for (int k = PW; k < PWs; k++) s_rows[(size_t)pl * PWs + k] = 0u;
        }
        __syncthreads();
        for (int pl = threadIdx.x; pl < np; pl += 256) {
            const uint64_t m = e.ms[p0 + pl];
            emit_record(s_rows + (size_t)pl * PWs, e.L1, meta_lvl(m), meta_end(m), e.sz_meta, sb + (int)(e.offR[p0 + pl] - g0));
        }
        __syncthreads();
        flush_staged(e.oR, g0, len, sbd);
        __syncthreads();   // staging buffers are reused by the next tile
    }
}

}  // namespace scb
