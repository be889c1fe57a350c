// scan_smem.cuh - core scan with the automaton resident in shared memory: parameter block, shared-memory layout and the
// device helpers (tile staging, SWAR 2-bit packing, the DFA step). The kernel is scan_smem2_k (scan_smem2.cuh); its
// predecessor scan_smem_k, whose pipeline is described here, was removed after it lost its A/B run (8.07 vs 7.63 ms).
//
// Persistent kernel, one CTA per SM; every WARP runs its own pipeline over tiles of 32 reads and never
// waits for another warp (no block barriers after the table is loaded), so the stalls of one warp's phase
// are covered by the other warps' phases. Per warp tile:
//   A  pack   the tile's 32 ASCII rows arrive as one bulk copy (cp.async.bulk + the warp's mbarrier; 16-byte cp.async
//             per lane with -DSCB_SCAN_BULK=0) - the warp's NEXT tile streams in while this one is walked; the lanes turn them into 2-bit packed words, 16 bases per word, SWAR on four
//             bytes at a time (bytes that are not ACGT/acgt become A, const.cpp:47-49). The words go to
//             global memory (coalesced: a tile's packed rows are one contiguous run) and to a shared-memory
//             copy with an odd row pitch (conflict-free for phase B).
//   B  walk   one lane per read: one 32-bit shared load per 16 bases, then per base
//             shift+mask -> address -> LDS.U16 of the u16 transition table (entries = next state * 4) ->
//             compare. States are renumbered so that "some core ends here" is state >= H0; a hit sets a
//             bit in the word's position mask and appends the state to a per-read u16 queue - three
//             predicated instructions, no divergence.
//   C  pick   per read, one pass over its hits: bucket rank, maximum core level, ordered distinct
//             candidates of that level (everything aho_search, reads.cpp:413-429, needs except the
//             running populations).
//   D  emit   candidate space comes from a per-warp bump allocator refilled from one global counter in
//             chunks (no block-wide scan); lists, counts, offsets, level out.
// Reads with more hits than the queue holds (very dense core sets) take a slower exact path.
#pragma once
#include "common.cuh"
#include "pipeline.cuh"

namespace scb {

constexpr int kHitQ = 32;      // queued hit states per read (u16 each)
constexpr int kHitQGuard = 16; // a 16-base word is only walked on the fast path if it cannot overflow the queue

struct ScanSmemParams {
    const uint8_t *seq; int64_t n; int L;
    const uint16_t *trans; const uint32_t *hit_rank; const uint8_t *rank_level;
    int ns, n_hit, nb, H0, R;
    uint8_t *lvl; uint16_t *ncand; uint64_t *cand_off; uint32_t *cand_rank; uint16_t *cand_pos;
    unsigned long long *cand_total; uint64_t cand_cap;   // bump counter over the candidate arrays (holes allowed)
    int64_t n_tiles;              // tiles of 32 reads
    uint32_t *packed; int PW;     // 2-bit packed copy of every read, PW = ceil(L/16) words
    uint32_t inv_pw;              // ceil(2^32 / PW)
    int pitch;                    // shared-memory row pitch of the packed tile in words (odd)
};

// shared-memory footprint (host + device agree through these); W = warps per CTA
constexpr int kCandChunk = 4096;   // candidate slots a warp takes from the global counter at a time
__host__ __device__ inline size_t scan_smem_table_bytes(int ns, int n_hit, int nb) {
    return (((size_t)ns * 8 + (size_t)n_hit * 4 + (size_t)nb) + 15) & ~(size_t)15;
}
__host__ __device__ inline int scan_smem_pitch(int PW) { return PW | 1; }
__host__ __device__ inline size_t scan_smem_warp_bytes(int L, int PW) {
    const size_t tile = (size_t)32 * L + 32;                        // 32*L is a multiple of 16
    const size_t pk = (size_t)32 * scan_smem_pitch(PW) * 4, q = (size_t)32 * kHitQ * 2, hm = (((size_t)32 * PW * 2) + 15) & ~(size_t)15;
    return ((tile + pk + q + hm) + 15 + 16) & ~(size_t)15;          // + the warp's mbarrier (last 16 bytes)
}
__host__ __device__ inline size_t scan_smem_total(int ns, int n_hit, int nb, int W, int L, int PW) {
    return scan_smem_table_bytes(ns, n_hit, nb) + (size_t)W * scan_smem_warp_bytes(L, PW);
}

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc) {
    uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// one warp stages its tile (32 rows, or fewer at the end of the input) into its own buffer
__device__ __forceinline__ void stage_warp_tile(const ScanSmemParams &p, int64_t tile, uint8_t *buf) {
    const int64_t row0 = tile * 32;
    int64_t rows = p.n - row0;
    if (rows > 32) rows = 32;
    if (rows <= 0) return;
    const int bytes = (int)rows * p.L;
    const uint8_t *src = p.seq + row0 * p.L;
    const int n16 = bytes >> 4;
    for (int k = lane_id(); k < n16; k += 32) cp_async16(buf + (k << 4), src + ((int64_t)k << 4));
    for (int k = (n16 << 4) + lane_id(); k < bytes; k += 32) buf[k] = src[k];
}

// ---- the same staging as ONE bulk copy per warp tile (TMA engine, SASS UBLKCP): a tile's rows are one contiguous,
// 16-byte aligned run of global memory. Lane 0 arms the warp's mbarrier with the byte count and issues the copy; the
// warp waits on the barrier's phase before packing. The (< 16) bytes a ragged last tile leaves over are copied by hand.
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(bar);
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(a), "r"(parity) : "memory");
}
// returns true if a copy was issued (the caller then owes one mbar_wait)
__device__ __forceinline__ bool stage_warp_tile_bulk(const ScanSmemParams &p, int64_t tile, uint8_t *buf, uint64_t *bar) {
    const int64_t row0 = tile * 32;
    int64_t rows = p.n - row0;
    if (rows > 32) rows = 32;
    if (rows <= 0) return false;
    const int bytes = (int)rows * p.L;
    const uint8_t *src = p.seq + row0 * p.L;
    const int b16 = bytes & ~15;
    for (int k = b16 + lane_id(); k < bytes; k += 32) buf[k] = src[k];
    if (b16 == 0) return false;
    if (lane_id() == 0) {
        const uint32_t d = (uint32_t)__cvta_generic_to_shared(buf), b = (uint32_t)__cvta_generic_to_shared(bar);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(b), "r"((uint32_t)b16) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                     ::"r"(d), "l"(src), "r"((uint32_t)b16), "r"(b) : "memory");
    }
    return true;
}

__device__ __forceinline__ void sts_u16(uint32_t saddr, uint32_t v) {
    asm volatile("st.shared.u16 [%0], %1;\n" ::"r"(saddr), "h"((uint16_t)v) : "memory");
}

// four ASCII bases (one per byte, first base in the low byte) -> 8 bits, first base in the top two bits.
// *bad gets a non-zero bit for any byte that is not one of ACGTacgt.
__device__ __forceinline__ uint32_t pack4(uint32_t w, uint32_t &bad) {
    const uint32_t x = w | 0x20202020u;
    const uint32_t cw = (x >> 1) & 0x03030303u;                    // a 0, c 1, t 2, g 3
    const uint32_t t2 = (cw >> 1) & ~cw & 0x01010101u;             // 1 in the bytes that hold t
    const uint32_t r = 0x61616161u + (cw << 1) + ((t2 << 4) - t2); // the letter each code stands for
    bad |= r ^ x;
    const uint32_t code = cw ^ ((cw >> 1) & 0x01010101u);          // a 0, c 1, g 2, t 3
    return (code * 0x40100401u) >> 24;
}

// same, exact for any byte: codes of bytes that are not ACGTacgt are forced to 0 (A)
__device__ __forceinline__ uint32_t pack4_masked(uint32_t w) {
    const uint32_t x = w | 0x20202020u;
    const uint32_t cw = (x >> 1) & 0x03030303u;
    const uint32_t t2 = (cw >> 1) & ~cw & 0x01010101u;
    const uint32_t r = 0x61616161u + (cw << 1) + ((t2 << 4) - t2);
    const uint32_t df = r ^ x;
    const uint32_t nz = (df | ((df & 0x7f7f7f7fu) + 0x7f7f7f7fu)) & 0x80808080u;   // 0x80 in the bytes where df != 0
    uint32_t code = cw ^ ((cw >> 1) & 0x01010101u);
    code &= ~((nz >> 7) | (nz >> 6));
    return (code * 0x40100401u) >> 24;
}

// the DFA over 2-bit codes, for the rare paths (queue overflow)
struct SmemDfa {
    const uint16_t *trans; const uint32_t *hit_rank; const uint8_t *rank_level; int H0;
};
__device__ __forceinline__ uint32_t pk_code(const uint32_t *row, int q) { return (row[q >> 4] >> (30 - 2 * (q & 15))) & 3u; }
__device__ __forceinline__ uint32_t dfa_step(const SmemDfa &d, uint32_t st, uint32_t c) { return (uint32_t)d.trans[(st << 2) | c] >> 2; }
__device__ __forceinline__ bool seen_before_smem(const uint32_t *row, int p, uint32_t r, const SmemDfa &d) {
    uint32_t st = 0;
    for (int q = 0; q < p; q++) {
        st = dfa_step(d, st, pk_code(row, q));
        if (st >= (uint32_t)d.H0 && d.hit_rank[st - d.H0] == r) return true;
    }
    return false;
}

}  // namespace scb
