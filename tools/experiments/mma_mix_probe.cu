// What does the MMA issue sequence of oz_gemm_kernel cost when operands are already in shared memory?
// One CTA per SM issues the production K-block sequence (oz::issue_kblock from csrc/ozaki_i8.cuh: 8 wide tcgen05.mma
// kind::i8 per 32-deep K step for S = 6, bn = 64) back to back on resident operands -- no TMA, no epilogue -- and reports
// cycles per 64-deep K block.  21 plane products of 128 x 64 x 64 need 1344 tensor cycles at the full kind::i8 rate; the
// real kernel takes ~1650.  The difference between this probe and 1344 is what the instruction mix itself costs (an MMA
// re-reads its 4 KB A slab from shared memory whatever its N); the difference between the real kernel and this probe is
// what operand delivery / pipeline stalls cost.  Also sweeps single-instruction widths N (cycles per instruction at
// M = 128, K = 32, A and B from shared memory) and candidate mixes for equal-width factor tiles (bn = 56 for m = 100).
// Not part of the product.  Build + run on a GPU box:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 --expt-relaxed-constexpr -I linearcorex_b200/csrc \
//        -o tools/experiments/bin/mma_mix_probe tools/experiments/mma_mix_probe.cu -lcuda && tools/experiments/bin/mma_mix_probe
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

namespace lcx { thread_local char g_err[512] = ""; }
#include "ozaki_i8.cuh"

using namespace lcx;
using namespace lcx::oz;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 2; } } while (0)

// MODE 0: production issue_kblock<S, KMAJOR, BN> (widths of 8 mod 16 use its spill form); 2: single width N = BN, 16 per "block"
template <int MODE, int S, bool KMAJOR, int BN>
__global__ void __launch_bounds__(128, 1) mix_kernel(int iters, long long* cycles_out) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_smem;
    const int tid = threadIdx.x, warp = tid >> 5;
    constexpr int STAGE = 7 * (kBM * kBK + 64 * kBK);  // room for up to 7 planes either side (+ spill rows)
    for (int i = tid; i < 2 * STAGE; i += 128) smem[i] = (uint8_t)((i * 7 + blockIdx.x) & 3);  // low entropy: no power throttling
    if (tid == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 0) tmem_alloc(&tmem_base_smem, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;
    long long t0 = 0, t1 = 0;
    if (tid == 0) {
        t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            const uint32_t sa = smem_u32(smem + (it & 1) * STAGE);
            const uint32_t sb = sa + 7 * kBM * kBK;
            if (MODE == 0) issue_kblock<S, KMAJOR, BN>(sa, sb, tmem_base, it == 0);
            else {
                const uint32_t idesc = make_idesc_i8(kBM, BN, KMAJOR ? 0 : 1, 0);
#pragma unroll
                for (int u = 0; u < 16; ++u) {
                    const int kk = u & 1, pl = (u >> 1) % 6;
                    const uint64_t da = KMAJOR ? make_smem_desc(sa + pl * 8192 + kk * 32, 16, 512, 4)
                                               : make_smem_desc(sa + pl * 8192 + kk * 4096, 8192, 1024, 2);
                    const uint64_t db = make_smem_desc(sb + kk * 32, 16, 512, 4);
                    umma_i8(tmem_base + (uint32_t)((u & 1) * 256), da, db, idesc, (it > 0 || u > 1) ? 1u : 0u);
                }
            }
        }
        umma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    tc_fence_after();
    if (tid == 0) {
        t1 = clock64();
        cycles_out[blockIdx.x] = t1 - t0;
    }
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 512);
}

// The production K block issued back to back (as in mix_kernel MODE 0) WHILE a second thread streams `fill_bytes` per K block
// from an L2-resident buffer into other shared memory with bulk copies (the TMA fill of the real kernel, minus any
// dependency between the two).  If MMA operand fetch and the fill share one shared-memory port, the K block slows down
// from its resident-operand time towards (operand bytes read + bytes filled) / port width.
template <int S, bool KMAJOR, int BN>
__global__ void __launch_bounds__(128, 1) mix_fill_kernel(int iters, int fill_bytes, const uint8_t* __restrict__ src,
                                                         long long* cycles_out, long long* fill_cycles_out) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(8) uint64_t bar;
    __shared__ __align__(8) uint64_t fbar[2];
    __shared__ uint32_t tmem_base_smem;
    const int tid = threadIdx.x, warp = tid >> 5;
    constexpr int STAGE = 6 * (kBM * kBK + 64 * kBK);          // 72 KB
    uint8_t* fill_dst = smem + 2 * STAGE;                       // 2 x 36 KB landing buffers
    for (int i = tid; i < 2 * STAGE; i += 128) smem[i] = (uint8_t)((i * 7 + blockIdx.x) & 3);
    if (tid == 0) {
        mbar_init(&bar, 1);
        mbar_init(&fbar[0], 1);
        mbar_init(&fbar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 0) tmem_alloc(&tmem_base_smem, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;
    if (tid == 0) {
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            const uint32_t sa = smem_u32(smem + (it & 1) * STAGE);
            const uint32_t sb = sa + 6 * kBM * kBK;
            issue_kblock<S, KMAJOR, BN>(sa, sb, tmem_base, it == 0);
        }
        umma_commit(&bar);
        mbar_wait(&bar, 0);
        cycles_out[blockIdx.x] = clock64() - t0;
    } else if (tid == 32 && fill_bytes > 0) {
        // each iteration = fill_bytes in two halves, each half one mbarrier phase; a half is re-issued once it has landed
        const int half = fill_bytes / 2;
        const uint8_t* mine = src + (size_t)blockIdx.x * (256 << 10);
        const long long t0 = clock64();
        for (int it = 0; it < 2 * iters; ++it) {
            const int b = it & 1;
            if (it >= 2) mbar_wait(&fbar[b], ((it >> 1) - 1) & 1);
            mbar_expect_tx(&fbar[b], half);
            for (int off = 0; off < half; off += 9216) {
                const int n = (half - off) < 9216 ? (half - off) : 9216;
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                 smem_u32(fill_dst + b * 36864 + off)),
                             "l"(mine + ((size_t)(it * 36864 + off) & ((256 << 10) - 1) & ~(size_t)1023)), "r"(n), "r"(smem_u32(&fbar[b]))
                             : "memory");
            }
        }
        mbar_wait(&fbar[0], ((2 * iters - 2) >> 1) & 1);
        mbar_wait(&fbar[1], ((2 * iters - 1) >> 1) & 1);
        fill_cycles_out[blockIdx.x] = clock64() - t0;
    }
    __syncthreads();
    tc_fence_after();
    if (warp == 0) tmem_dealloc(tmem_base, 512);
}

template <int S, bool KMAJOR, int BN>
static int run_fill(const char* label, int fill_bytes, int sms, long long* d_cycles, const uint8_t* src) {
    const int smem = 220 * 1024;
    auto kern = mix_fill_kernel<S, KMAJOR, BN>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const int iters = 20000;
    kern<<<sms, 128, smem>>>(200, fill_bytes, src, d_cycles, d_cycles + 512);
    CK(cudaDeviceSynchronize());
    kern<<<sms, 128, smem>>>(iters, fill_bytes, src, d_cycles, d_cycles + 512);
    CK(cudaDeviceSynchronize());
    static long long h[1024];
    CK(cudaMemcpy(h, d_cycles, 1024 * sizeof(long long), cudaMemcpyDeviceToHost));
    double a = 0, b = 0;
    for (int i = 0; i < sms; ++i) { a += (double)h[i]; b += (double)h[512 + i]; }
    printf("%-58s MMA %8.1f cycles per block;  fill of %3d KB per block: %8.1f cycles per block (%5.1f B/clk)\n", label,
           a / sms / iters, fill_bytes >> 10, fill_bytes ? b / sms / iters : 0.0, fill_bytes ? fill_bytes / (b / sms / iters) : 0.0);
    return 0;
}

template <int MODE, int S, bool KMAJOR, int BN>
static int run(const char* label, double ideal_cycles, int sms, long long* d_cycles) {
    const int smem = 210 * 1024;
    auto kern = mix_kernel<MODE, S, KMAJOR, BN>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const int iters = 20000;
    kern<<<sms, 128, smem>>>(200, d_cycles);
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0));
    kern<<<sms, 128, smem>>>(iters, d_cycles);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    static long long h[1024];
    CK(cudaMemcpy(h, d_cycles, sms * sizeof(long long), cudaMemcpyDeviceToHost));
    double sum = 0;
    for (int i = 0; i < sms; ++i) sum += (double)h[i];
    const double cyc = sum / sms / iters;
    printf("%-58s %9.1f cycles per block  (ideal %7.1f, %5.1f %% of the tensor rate)  %.2f ms\n", label, cyc, ideal_cycles,
           100.0 * ideal_cycles / cyc, ms);
    return 0;
}

int main() {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    long long* d_cycles;
    CK(cudaMalloc(&d_cycles, 1024 * sizeof(long long)));
    printf("# %s, %d SMs; one CTA per SM, operands resident in shared memory, low-entropy data (no power throttling)\n", prop.name, sms);
    printf("# single widths: 16 instructions of M = 128, K = 32 per block (ideal = 16 * N / 2 cycles)\n");
    if (run<2, 6, true, 32>("K-major A, N = 32", 16 * 16, sms, d_cycles)) return 2;
    if (run<2, 6, true, 64>("K-major A, N = 64", 16 * 32, sms, d_cycles)) return 2;
    if (run<2, 6, true, 96>("K-major A, N = 96", 16 * 48, sms, d_cycles)) return 2;
    if (run<2, 6, true, 128>("K-major A, N = 128", 16 * 64, sms, d_cycles)) return 2;
    if (run<2, 6, true, 192>("K-major A, N = 192", 16 * 96, sms, d_cycles)) return 2;
    if (run<2, 6, true, 224>("K-major A, N = 224", 16 * 112, sms, d_cycles)) return 2;
    if (run<2, 6, true, 256>("K-major A, N = 256", 16 * 128, sms, d_cycles)) return 2;
    if (run<2, 6, false, 64>("MN-major A, N = 64", 16 * 32, sms, d_cycles)) return 2;
    if (run<2, 6, false, 128>("MN-major A, N = 128", 16 * 64, sms, d_cycles)) return 2;
    if (run<2, 6, false, 256>("MN-major A, N = 256", 16 * 128, sms, d_cycles)) return 2;
    printf("# production K block (64 deep): S(S+1)/2 plane products of 128 x bn x 64\n");
    if (run<0, 6, true, 64>("S=6 bn=64 K-major A  (first contraction)", 21 * 64.0, sms, d_cycles)) return 2;
    if (run<0, 6, false, 64>("S=6 bn=64 MN-major A (second contraction)", 21 * 64.0, sms, d_cycles)) return 2;
    if (run<0, 6, true, 48>("S=6 bn=48 K-major A  (tail tile of m = 100)", 21 * 48.0, sms, d_cycles)) return 2;
    if (run<0, 6, false, 48>("S=6 bn=48 MN-major A", 21 * 48.0, sms, d_cycles)) return 2;
    if (run<0, 6, true, 56>("S=6 bn=56 K-major A, equal tiles + spill", 21 * 56.0, sms, d_cycles)) return 2;
    if (run<0, 6, false, 56>("S=6 bn=56 MN-major A, equal tiles + spill", 21 * 56.0, sms, d_cycles)) return 2;
    if (run<0, 5, true, 64>("S=5 bn=64 K-major A", 15 * 64.0, sms, d_cycles)) return 2;
    if (run<0, 3, true, 128>("S=3 bn=128 K-major A", 6 * 128.0, sms, d_cycles)) return 2;
    if (run<0, 7, true, 64>("S=7 bn=64 K-major A", 28 * 64.0, sms, d_cycles)) return 2;
    printf("# production K block with a concurrent bulk-copy fill of shared memory from an L2-resident buffer\n");
    uint8_t* src;
    CK(cudaMalloc(&src, (size_t)sms * (256 << 10) + (64 << 10)));
    CK(cudaMemset(src, 1, (size_t)sms * (256 << 10)));
    if (run_fill<6, true, 64>("S=6 bn=64 K-major A, no fill", 0, sms, d_cycles, src)) return 2;
    if (run_fill<6, true, 64>("S=6 bn=64 K-major A + 36 KB fill", 36864, sms, d_cycles, src)) return 2;
    if (run_fill<6, true, 64>("S=6 bn=64 K-major A + 72 KB fill (production ratio)", 73728, sms, d_cycles, src)) return 2;
    if (run_fill<6, false, 64>("S=6 bn=64 MN-major A + 72 KB fill", 73728, sms, d_cycles, src)) return 2;
    if (run_fill<6, true, 56>("S=6 bn=56 K-major A + 72 KB fill", 73728, sms, d_cycles, src)) return 2;
    if (run_fill<5, true, 64>("S=5 bn=64 K-major A + 60 KB fill (production ratio)", 61440, sms, d_cycles, src)) return 2;
    if (run_fill<7, true, 64>("S=7 bn=64 K-major A + 72 KB fill (production: 84 KB)", 73728, sms, d_cycles, src)) return 2;
    return 0;
}
