// Shared helpers for the mvx_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/mvx_b200.h"

namespace mvx {

void set_error(const char *fmt, ...);
void count_launch(int n = 1);

#define MVX_CUDA_CHECK(expr)                                                                         \
    do {                                                                                             \
        cudaError_t _e = (expr);                                                                     \
        if (_e != cudaSuccess) {                                                                     \
            mvx::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e));    \
            return MVX_ECUDA;                                                                        \
        }                                                                                            \
    } while (0)

#define MVX_LAUNCH_CHECK()                                                                           \
    do {                                                                                             \
        mvx::count_launch();                                                                         \
        cudaError_t _e = cudaGetLastError();                                                         \
        if (_e != cudaSuccess) {                                                                     \
            mvx::set_error("%s:%d: kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(_e)); \
            return MVX_ECUDA;                                                                        \
        }                                                                                            \
    } while (0)

#define MVX_REQUIRE(cond, code, msg)                                                                 \
    do {                                                                                             \
        if (!(cond)) {                                                                               \
            mvx::set_error("%s:%d: %s", __FILE__, __LINE__, msg);                                    \
            return code;                                                                             \
        }                                                                                            \
    } while (0)

constexpr int kSMs = 148;  // B200

static inline int64_t round_up(int64_t a, int64_t b) { return (a + b - 1) / b * b; }
static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// streaming (evict-first) 16-byte store: the dense grid is written once and not re-read by this path
__device__ __forceinline__ void st_cs_f4(float4 *p, float4 v) {
    asm volatile("st.global.cs.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}
__device__ __forceinline__ float4 ld_nc_f4(const float4 *p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}

// block-wide exclusive scan of one int per thread (blockDim.x <= 1024, multiple of 32). Returns the
// exclusive prefix; *total receives the block sum (valid in all threads).
__device__ __forceinline__ int block_exclusive_scan(int v, int *total) {
    __shared__ int warp_sums[32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    int inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += t;
    }
    __syncthreads();  // protect warp_sums reuse across calls
    if (lane == 31) warp_sums[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        int w = lane < nw ? warp_sums[lane] : 0;
        int winc = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, winc, d);
            if (lane >= d) winc += t;
        }
        warp_sums[lane] = winc;  // inclusive over warps
    }
    __syncthreads();
    const int base = wid == 0 ? 0 : warp_sums[wid - 1];
    *total = warp_sums[nw - 1];
    return base + inc - v;
}

// the same for one 64-bit value per thread (two packed 32-bit counters scan in one pass: they do not carry into each other
// as long as each total stays below 2^32)
__device__ __forceinline__ unsigned long long block_exclusive_scan_u64(unsigned long long v, unsigned long long *total) {
    __shared__ unsigned long long warp_sums64[32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    unsigned long long inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned long long t = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += t;
    }
    __syncthreads();  // protect warp_sums64 reuse across calls
    if (lane == 31) warp_sums64[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        unsigned long long winc = lane < nw ? warp_sums64[lane] : 0ull;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned long long t = __shfl_up_sync(0xffffffffu, winc, d);
            if (lane >= d) winc += t;
        }
        warp_sums64[lane] = winc;  // inclusive over warps
    }
    __syncthreads();
    const unsigned long long base = wid == 0 ? 0ull : warp_sums64[wid - 1];
    *total = warp_sums64[nw - 1];
    return base + inc - v;
}

}  // namespace mvx
