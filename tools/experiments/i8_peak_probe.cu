// Measures the dense tcgen05 kind::i8 rate of this GPU: one CTA per SM issues back-to-back 128 x 256 x 32 int8 MMAs on
// operands that stay in shared memory (no TMA traffic, no epilogue), so the only limits are the tensor pipe and the power
// cap.  Prints a burst figure (one ~50 ms launch from idle clocks) and a sustained one (back-to-back launches for ~4 s).
// Context for the roofline denominator of bench.py, which uses 2 x the driver-measured bf16 rate for the int8 kernels.
// Not part of the product.  Build + run on a GPU box:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o /tmp/i8_peak tools/experiments/i8_peak_probe.cu && /tmp/i8_peak
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 2; } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)(layout & 7) << 61;
    return d;
}
__host__ __device__ constexpr uint32_t make_idesc_i8(int M, int N) {
    return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

constexpr int kN = 256;

__global__ void __launch_bounds__(128, 1) peak_kernel(int iters, int* sink, int random_data) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* sa = smem;              // 128 rows x 64 B (SW64 K-major)
    uint8_t* sb = smem + 8192;       // 256 rows x 64 B
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_smem;
    const int tid = threadIdx.x, warp = tid >> 5;
    // operand bytes: tiny constants (few toggling bits = little power) or full-range pseudo-random int8 (what real digit
    // planes look like: the power cap then sets the clock)
    for (int i = tid; i < 8192 + kN * 64; i += 128) {
        uint32_t h = (uint32_t)i * 2654435761u + (uint32_t)blockIdx.x * 40503u;
        h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
        smem[i] = random_data ? (uint8_t)(h & 0xff) : (uint8_t)((i * 7 + blockIdx.x) & 3);
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_smem;
    if (tid == 0) {
        const uint32_t idesc = make_idesc_i8(128, kN);
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {   // two accumulators x two K steps: independent instructions back to back
                const uint64_t da = make_desc(smem_u32(sa) + (u & 1) * 32, 16, 512, 4);
                const uint64_t db = make_desc(smem_u32(sb) + (u & 1) * 32, 16, 512, 4);
                asm volatile(
                    "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                    "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_base + (uint32_t)((u >> 1) * kN)),
                    "l"(da), "l"(db), "r"(idesc), "r"(it > 0 ? 1u : 0u));
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    uint32_t done = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred q;\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2;\n\tselp.u32 %0, 1, 0, q;\n\t}\n"
                     : "=r"(done)
                     : "r"(smem_u32(&bar)), "r"(0u)
                     : "memory");
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (tid == 0 && sink != nullptr && iters < 0) sink[0] = 1;
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
}

int main() {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    const int smem = 200 * 1024;  // one CTA per SM
    CK(cudaFuncSetAttribute(peak_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const int iters = 100000;                                      // 4 MMAs each
    const double ops_per_launch = 2.0 * 128 * kN * 32 * 4.0 * iters * sms;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    peak_kernel<<<sms, 128, smem>>>(1000, nullptr, 0);            // module load
    CK(cudaDeviceSynchronize());
    for (int rnd = 0; rnd < 2; ++rnd) {
        float ms = 0;
        CK(cudaEventRecord(e0));
        peak_kernel<<<sms, 128, smem>>>(iters, nullptr, rnd);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&ms, e0, e1));
        printf("%s, %d SMs, %s operands: burst %.1f TOP/s (one launch of %.1f ms); ", prop.name, sms,
               rnd ? "random int8" : "low-entropy", ops_per_launch / ms / 1e9, ms);
        const int reps = (int)(4000.0 / ms) + 1;
        CK(cudaEventRecord(e0));
        for (int r = 0; r < reps; ++r) peak_kernel<<<sms, 128, smem>>>(iters, nullptr, rnd);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&ms, e0, e1));
        printf("sustained %.1f TOP/s (%d launches, %.2f s); ", ops_per_launch * reps / ms / 1e9, reps, ms / 1e3);
        const int tail = reps / 4 > 0 ? reps / 4 : 1;
        CK(cudaEventRecord(e0));
        for (int r = 0; r < tail; ++r) peak_kernel<<<sms, 128, smem>>>(iters, nullptr, rnd);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&ms, e0, e1));
        printf("last quarter again: %.1f TOP/s\n", ops_per_launch * tail / ms / 1e9);
    }
    return 0;
}
