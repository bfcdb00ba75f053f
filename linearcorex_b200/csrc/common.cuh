// Shared helpers for the sm_100a Linear CorEx kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

namespace lcx {

constexpr int kWarp = 32;

// ---- error plumbing -------------------------------------------------------------------------
// No exceptions cross the C ABI: every entry point returns an int and records a message that
// lcx_last_error() hands back.
extern thread_local char g_err[512];

inline int fail(int code, const char* what, const char* detail) {
    snprintf(g_err, sizeof(g_err), "%s: %s", what, detail ? detail : "");
    return code;
}

#define LCX_CUDA(expr)                                                                   \
    do {                                                                                 \
        cudaError_t _e = (expr);                                                         \
        if (_e != cudaSuccess) return ::lcx::fail(-2, #expr, cudaGetErrorString(_e));    \
    } while (0)

#define LCX_REQUIRE(cond, msg)                                     \
    do {                                                           \
        if (!(cond)) return ::lcx::fail(-1, "bad argument", msg);  \
    } while (0)

#define LCX_TRY(expr)              \
    do {                           \
        int _rc = (expr);          \
        if (_rc < 0) return _rc;   \
    } while (0)

// cudaFuncSetAttribute applies to the current device's context only: one "done" flag per device ordinal, so a second
// session on another GPU of the same process configures its own copy of the kernel.
struct PerDeviceOnce {
    bool done[64];
    bool first_time() {
        int d = 0;
        if (cudaGetDevice(&d) != cudaSuccess) return true;
        d &= 63;
        if (done[d]) return false;
        done[d] = true;
        return true;
    }
};

inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }
inline long long round_up(long long a, long long b) { return (a + b - 1) / b * b; }

// ---- device utilities -----------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// |v| folded into a running maximum.  A NaN or an infinity turns the maximum into +inf and it stays there (fmax keeps
// +inf), so non-finite input poisons the scale derived from the maximum -- and through it every output -- the way it
// poisons numpy's float64 products in the reference, instead of being skipped by fmax's NaN rule.
__device__ __forceinline__ double amax_acc(double mx, double v) {
    const double a = fabs(v);
    return (a <= 1.7976931348623157e308) ? fmax(mx, a) : __longlong_as_double(0x7ff0000000000000LL);
}

__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Deterministic block-wide sum for blockDim.x == 256 (fixed shuffle tree + fixed warp order).
// `scratch` must hold 8 doubles.  Result valid on every thread.
__device__ __forceinline__ double block_sum_256(double v, double* scratch) {
    v = warp_sum(v);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) scratch[w] = v;
    __syncthreads();
    double t = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += scratch[i];
    return t;
}

__device__ __forceinline__ double block_max_256(double v, double* scratch) {
    v = warp_max(v);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) scratch[w] = v;
    __syncthreads();
    double t = scratch[0];
#pragma unroll
    for (int i = 1; i < 8; ++i) t = fmax(t, scratch[i]);
    return t;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}

// 16-byte async copy global -> shared with zero fill of the bytes beyond `src_bytes`.
__device__ __forceinline__ void cp_async_16(uint32_t dst, const void* src, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(src), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// D(8x8) += A(8x4) * B(4x8), all fp64.  SASS: DMMA.8x8x4 (the native fp64 tensor shape on sm_100a).
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

}  // namespace lcx
