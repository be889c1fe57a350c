// common.cuh - shared device/host helpers for the sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <string>

namespace scb {

// ---- error plumbing ---------------------------------------------------------------------------
struct CudaError {
    std::string msg;
};
#define SCB_CUDA(expr)                                                                              \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess)                                                                      \
            throw ::scb::CudaError{std::string(#expr) + ": " + cudaGetErrorString(_e) + " @" + __FILE__ + ":" + \
                                   std::to_string(__LINE__)};                                       \
    } while (0)

// every kernel launch goes through this so scb_kernel_launches() is an honest count
extern std::atomic<long long> g_launches;
#define SCB_LAUNCH(kernel, grid, block, smem, stream, ...)                                          \
    do {                                                                                            \
        auto _kfn = kernel;                                                                         \
        _kfn<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);                                   \
        ::scb::g_launches++;                                                                        \
        SCB_CUDA(cudaGetLastError());                                                               \
    } while (0)

// ---- base coding: const.cpp:47-49 / const.h:127 --------------------------------------------------
// A,a,N,n and anything else -> 0; C,c -> 1; G,g -> 2; T,t -> 3.
__host__ __device__ __forceinline__ uint32_t base_code(uint32_t c) {
    c |= 0x20u;
    return (c == 'c') ? 1u : (c == 'g') ? 2u : (c == 't') ? 3u : 0u;
}
// SZ_READ, const.h:63
__host__ __device__ __forceinline__ int sz_read(int l) { return (l >> 2) + ((l & 3) != 0); }

__host__ __device__ __forceinline__ int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }
__device__ __forceinline__ uint32_t lanemask_lt() {
    uint32_t m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

// Loads for random gathers of short rows: ask L2 to fetch 64-byte granules from HBM instead of its
// default 128 (SASS LDG.E.LTC64B). Halves the over-fetch when only 2-150 bytes around an address are used.
__device__ __forceinline__ uint4 ldg_g64(const uint4 *p) {
    uint4 v;
    asm volatile("ld.global.L2::64B.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ uint32_t ldg_g64(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.global.L2::64B.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ uint32_t ldg_g64(const uint16_t *p) {
    uint32_t v;
    asm volatile("ld.global.L2::64B.u16 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ uint32_t ldg_g64(const uint8_t *p) {
    uint32_t v;
    asm volatile("ld.global.L2::64B.u8 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ int64_t ldg_g64(const int64_t *p) {
    int64_t v;
    asm volatile("ld.global.L2::64B.s64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
}

constexpr int kSMs = 148;  // B200
#define SCB_ROOT_ID_DEV ((1 << 30) - 1)  // MAXBIN-1, const.h:94

}  // namespace scb
