// Feasibility probe for the next lever of oz_gemm_kernel: the M-side int8 operand staged through TMEM
// (tcgen05.cp smem -> TMEM, then tcgen05.mma with A in TMEM) instead of being re-read from shared memory by every MMA.
// One CTA computes D[128 x 64] = A[128 x 32] * B[64 x 32]^T (int8 -> int32) three ways and compares with the host:
//   mode 0: A and B from shared memory (the form the product kernel uses; validates the hand-made SW64 image)
//   mode 1: tcgen05.cp.128x256b of the same A image into TMEM, then the A-from-TMEM MMA
// Not part of the product; nothing links it.  Build + run (GPU box):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o /tmp/tmem_a_probe tools/tmem_a_probe.cu && /tmp/tmem_a_probe
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
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

// status[0]: 0 ok, 1 = barrier timeout
__global__ void __launch_bounds__(128, 1) probe_kernel(const int8_t* __restrict__ a_img, const int8_t* __restrict__ b_img,
                                                       int32_t* __restrict__ d_out, int mode, int* status) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* sa = smem;            // 128 rows x 64 B (SW64 image, K = 64 of which the probe uses the first 32)
    uint8_t* sb = smem + 8192;     // 64 rows x 64 B
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_smem;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 8192; i += 128) sa[i] = (uint8_t)a_img[i];
    for (int i = tid; i < 4096; i += 128) sb[i] = (uint8_t)b_img[i];
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // generic-proxy writes to smem must be visible to the async proxy (tensor core / tcgen05.cp reads)
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "r"(128));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_smem;
    const uint32_t d_tmem = tmem_base;        // 64 columns of accumulators
    const uint32_t a_tmem = tmem_base + 64;   // 8 columns: 128 lanes x 32 B of A
    if (tid == 0) {
        const uint64_t da = make_desc(smem_u32(sa), 16, 512, 4);
        const uint64_t db = make_desc(smem_u32(sb), 16, 512, 4);
        const uint32_t idesc = make_idesc_i8(128, 64);
        if (mode == 0) {
            asm volatile(
                "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
                "l"(da), "l"(db), "r"(idesc), "r"(0u));
        } else {
            asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(a_tmem), "l"(da));
            asm volatile(
                "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d_tmem),
                "r"(a_tmem), "l"(db), "r"(idesc), "r"(0u));
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    // bounded wait
    bool ok = false;
    for (int it = 0; it < 2000000 && !ok; ++it) {
        uint32_t p;
        asm volatile("{\n\t.reg .pred q;\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2;\n\tselp.u32 %0, 1, 0, q;\n\t}\n"
                     : "=r"(p)
                     : "r"(smem_u32(&bar)), "r"(0u)
                     : "memory");
        ok = p != 0;
    }
    if (!ok && tid == 0) status[0] = 1;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (ok) {
        const uint32_t lane_addr = d_tmem + ((uint32_t)(warp * 32) << 16);
        for (int c0 = 0; c0 < 64; c0 += 16) {
            uint32_t r[16];
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                  "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                : "r"(lane_addr + (uint32_t)c0));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            for (int j = 0; j < 16; ++j) d_out[tid * 64 + c0 + j] = (int32_t)r[j];
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128));
    }
}

// SW64 K-major image: row r at 64 B pitch, 16-byte chunk c stored at chunk c ^ ((r >> 1) & 3)
static void make_image(const int8_t* src, int rows, int k_valid, int8_t* img) {
    memset(img, 0, (size_t)rows * 64);
    for (int r = 0; r < rows; ++r)
        for (int k = 0; k < k_valid; ++k) {
            const int c = k >> 4, b = k & 15;
            img[r * 64 + ((c ^ ((r >> 1) & 3)) << 4) + b] = src[r * k_valid + k];
        }
}

int main() {
    const int M = 128, N = 64, K = 32;
    static int8_t a[M * K], b[N * K], a_img[M * 64], b_img[N * 64];
    static int32_t want[M * N], got[M * N];
    srand(7);
    for (int i = 0; i < M * K; ++i) a[i] = (int8_t)(rand() % 255 - 127);
    for (int i = 0; i < N * K; ++i) b[i] = (int8_t)(rand() % 255 - 127);
    for (int i = 0; i < M; ++i)
        for (int j = 0; j < N; ++j) {
            int32_t s = 0;
            for (int k = 0; k < K; ++k) s += (int32_t)a[i * K + k] * (int32_t)b[j * K + k];
            want[i * N + j] = s;
        }
    make_image(a, M, K, a_img);
    make_image(b, N, K, b_img);
    int8_t *da, *db;
    int32_t* dd;
    int* dstat;
    CK(cudaMalloc(&da, sizeof(a_img)));
    CK(cudaMalloc(&db, sizeof(b_img)));
    CK(cudaMalloc(&dd, sizeof(got)));
    CK(cudaMalloc(&dstat, sizeof(int)));
    CK(cudaMemcpy(da, a_img, sizeof(a_img), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(db, b_img, sizeof(b_img), cudaMemcpyHostToDevice));
    CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384));
    int rc = 0;
    for (int mode = 0; mode < 2; ++mode) {
        CK(cudaMemset(dd, 0xff, sizeof(got)));
        CK(cudaMemset(dstat, 0, sizeof(int)));
        probe_kernel<<<1, 128, 16384>>>(da, db, dd, mode, dstat);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("mode %d: kernel failed: %s\n", mode, cudaGetErrorString(e)); return 3; }
        int st = 0;
        CK(cudaMemcpy(&st, dstat, sizeof(int), cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(got, dd, sizeof(got), cudaMemcpyDeviceToHost));
        long long bad = 0;
        for (int i = 0; i < M * N; ++i) bad += got[i] != want[i];
        printf("mode %d (%s): status %d, %lld of %d outputs differ; D[0][0..3] = %d %d %d %d (want %d %d %d %d), D[5][7] = %d (want %d)\n",
               mode, mode == 0 ? "A from shared memory" : "A through tcgen05.cp into TMEM", st, bad, M * N, got[0], got[1], got[2],
               got[3], want[0], want[1], want[2], want[3], got[5 * N + 7], want[5 * N + 7]);
        rc |= (bad != 0 || st != 0) << mode;
    }
    printf("rc = %d\n", rc);
    return 0;
}
