// Operand-delivery probe for oz_gemm_kernel.  First seven cases measured at the very end of round 1
// (profiles/r01_dsmem_feed_probe.txt); the HBM-resident and multicast cases below were added afterwards and are NOT yet run.
//
// Question.  oz_gemm_kernel is bound by operand delivery into the SMs: every SM of a CTA pair receives the whole X~ tile
// (48 KB per 64-deep K block; TMA multicast halves the L2 reads, not the bytes delivered to each SM) plus its 24 KB of
// factor planes, against ~42 B/clk per SM of crossbar -> SM delivery.  If bytes that arrive from the PEER SM
// through distributed shared memory (cp.async.bulk.shared::cluster.shared::cta) do not count against that ceiling, each CTA
// could fetch half of the X~ planes from L2 for itself only and forward them to its peer: 48 KB from L2 + 24 KB from the
// peer per K block instead of 72 KB from L2 -- the tensor pipe would go from 72 % to > 90 % active.
//
// Probe.  Clusters of two CTAs on all 148 SMs, no math: per iteration each CTA (a) bulk-loads `l2_bytes` from an
// L2-resident region into a shared-memory stage and (b) bulk-copies `peer_bytes` from its own shared memory into the
// peer's stage, both in 8 KB pieces (one digit plane of a 128 x 64 tile), 3 stages deep, with credits so a stage is only
// overwritten once its receiver has waited on it.  Reported: bytes per clock per SM for L2 only, peer only, and both.
//   additive  (both ~ L2-only + peer-only)  -> build the forwarding variant of oz_gemm_kernel;
//   not additive (both ~ L2-only)            -> delivery into an SM is one port whatever the source; drop the idea.
//
// Round-1 result (B200, L2-resident source, one issuing thread per CTA): L2 only 72 KB / 48 KB / 24 KB per iteration = 1396 / 1137 /
// 852 cycles (52.8 / 43.2 / 28.8 B/clk per SM; ~570 cycles of per-iteration issue + wait overhead in the single thread, ~90 B/clk
// marginal); peer only 24 KB = 1957 cycles (12.6 B/clk); L2 48 KB + peer 24 KB = 2941 cycles -- WORSE than 72 KB from L2 alone.
// So (1) DSMEM bulk copies are slow (<= 12.6 B/clk per SM) and slow the L2 stream down as well: forwarding is dead;
// (2) L2-resident data reaches every SM at >= 52.8 B/clk, i.e. the real kernel's 1 644 cycles per 72 KB K block (44.8 B/clk)
// is NOT the fabric's L2 -> SM ceiling.  What the real kernel does differently: its X~ planes come from HBM (6 GB per launch, L2 hit
// rate 40-50 %) and are multicast to the CTA pair.  `feed_probe_mc` and the HBM-resident cases reproduce exactly that traffic
// without any math; if they also run at ~1 650 cycles the bound is this data path, otherwise it is inside the SM.
//
// build + run (one B200):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/dsmem_feed_probe tools/experiments/dsmem_feed_probe.cu && /tmp/dsmem_feed_probe
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x)                                                                         \
    do {                                                                              \
        cudaError_t e_ = (x);                                                         \
        if (e_ != cudaSuccess) {                                                      \
            fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); \
            exit(1);                                                                  \
        }                                                                             \
    } while (0)

constexpr int kStages = 3;
constexpr int kPiece = 8192;           // one int8 digit plane of a 128 x 64 tile
constexpr int kMaxL2 = 49152;          // per-stage capacity for bytes from L2 (+ kMaxPeer more when nothing comes from the peer)
constexpr int kMaxPeer = 24576;        // per-stage capacity for bytes from the peer
constexpr int kSmem = kStages * (kMaxL2 + kMaxPeer) + 1024;

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t peer_addr(uint32_t local, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local), "r"(rank));
    return r;
}
__device__ __forceinline__ void bar_init(uint64_t* b, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(count));
}
__device__ __forceinline__ void bar_expect(uint64_t* b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bar_wait(uint64_t* b, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "W_LOOP:\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra W_DONE;\n\t"
        "bra W_LOOP;\n\t"
        "W_DONE:\n\t"
        "}\n" ::"r"(s32(b)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bar_arrive_remote(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(s32(bar))
                 : "memory");
}
__device__ __forceinline__ void bulk_s2peer(uint32_t dst_cluster, uint32_t src_cta, uint32_t bytes, uint32_t bar_cluster) {
    asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_cluster),
                 "r"(src_cta), "r"(bytes), "r"(bar_cluster)
                 : "memory");
}

// One elected thread per CTA drives everything (as the TMA producer warp of the real kernel does).
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
feed_probe(const uint8_t* __restrict__ src, long long region_bytes, int l2_bytes, int peer_bytes, int iters,
           unsigned long long* __restrict__ cycles) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(8) uint64_t full_l2[kStages], full_peer[kStages], empty_peer[kStages];
    uint32_t rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    const uint32_t peer = rank ^ 1u;
    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) {
            bar_init(&full_l2[s], 1);
            bar_init(&full_peer[s], 1);
            bar_init(&empty_peer[s], 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (threadIdx.x == 0) {
        const uint8_t* my_src = src + (long long)blockIdx.x * region_bytes;
        long long off = 0;
        const unsigned long long t0 = clock64();
        for (int i = 0; i < iters + kStages - 1; ++i) {
            if (i < iters) {
                const int s = i % kStages;
                const uint32_t ph = (uint32_t)(i / kStages) & 1u;
                uint8_t* st_l2 = smem + s * (kMaxL2 + kMaxPeer);
                uint8_t* st_peer = st_l2 + kMaxL2;
                if (peer_bytes > 0) {
                    // the peer has waited on what I sent into its stage s one round ago (first round: passes at once)
                    bar_wait(&empty_peer[s], ph ^ 1u);
                    const uint32_t dst = peer_addr(s32(st_peer), peer), bar = peer_addr(s32(&full_peer[s]), peer);
                    for (int b = 0; b < peer_bytes; b += kPiece)  // source: my own stage (contents are irrelevant here)
                        bulk_s2peer(dst + b, s32(st_l2) + b, min(kPiece, peer_bytes - b), bar);
                    bar_expect(&full_peer[s], (uint32_t)peer_bytes);  // what the peer sends me this iteration
                }
                if (l2_bytes > 0) {
                    bar_expect(&full_l2[s], (uint32_t)l2_bytes);
                    for (int b = 0; b < l2_bytes; b += kPiece) {
                        bulk_g2s(s32(st_l2) + b, my_src + off, min(kPiece, l2_bytes - b), &full_l2[s]);
                        off += kPiece;
                        if (off + kPiece > region_bytes) off = 0;
                    }
                }
            }
            const int j = i - (kStages - 1);  // consume with a lag of kStages - 1 iterations
            if (j >= 0) {
                const int s = j % kStages;
                const uint32_t ph = (uint32_t)(j / kStages) & 1u;
                if (l2_bytes > 0) bar_wait(&full_l2[s], ph);
                if (peer_bytes > 0) {
                    bar_wait(&full_peer[s], ph);
                    bar_arrive_remote(peer_addr(s32(&empty_peer[s]), peer));  // credit: the peer may overwrite my stage s
                }
            }
        }
        cycles[blockIdx.x] = clock64() - t0;
    }
    __syncthreads();
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__device__ __forceinline__ void bar_arrive_local(uint64_t* b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(b)) : "memory");
}
__device__ __forceinline__ void bulk_g2s_mc(uint32_t dst, const void* src, uint32_t bytes, uint64_t* bar, uint16_t mask) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(dst),
        "l"(src), "r"(bytes), "r"(s32(bar)), "h"(mask)
        : "memory");
}

// The real kernel's traffic without its math: per iteration 48 KB shared by the pair (six 8 KB pieces, piece p fetched by cluster
// rank p % 2 and multicast into both CTAs) + 24 KB of the CTA's own; a stage is refilled once BOTH CTAs have waited on it.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
feed_probe_mc(const uint8_t* __restrict__ src, long long region_bytes, int iters, unsigned long long* __restrict__ cycles) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(8) uint64_t full[kStages], empty[kStages];
    uint32_t rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    const uint32_t peer = rank ^ 1u;
    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) {
            bar_init(&full[s], 1);
            bar_init(&empty[s], 2);  // this CTA and its peer have both consumed the stage
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (threadIdx.x == 0) {
        const uint8_t* shared_src = src + (long long)(blockIdx.x & ~1u) * region_bytes;  // the pair's common operand
        const uint8_t* own_src = src + (long long)(blockIdx.x | 1u) * region_bytes + (long long)rank * (region_bytes / 2);
        long long off_sh = 0, off_own = 0;
        const unsigned long long t0 = clock64();
        for (int i = 0; i < iters + kStages - 1; ++i) {
            if (i < iters) {
                const int s = i % kStages;
                const uint32_t ph = (uint32_t)(i / kStages) & 1u;
                uint8_t* st = smem + s * (kMaxL2 + kMaxPeer);
                bar_wait(&empty[s], ph ^ 1u);
                bar_expect(&full[s], (uint32_t)(kMaxL2 + kMaxPeer));
                for (int p = 0; p < kMaxL2 / kPiece; ++p)
                    if ((uint32_t)(p & 1) == rank) bulk_g2s_mc(s32(st) + p * kPiece, shared_src + off_sh + (long long)p * kPiece, kPiece, &full[s], 3);
                off_sh += kMaxL2;
                if (off_sh + kMaxL2 > region_bytes) off_sh = 0;
                for (int p = 0; p < kMaxPeer / kPiece; ++p)
                    bulk_g2s(s32(st) + kMaxL2 + p * kPiece, own_src + off_own + (long long)p * kPiece, kPiece, &full[s]);
                off_own += kMaxPeer;
                if (off_own + kMaxPeer > region_bytes / 2) off_own = 0;
            }
            const int j = i - (kStages - 1);
            if (j >= 0) {
                const int s = j % kStages;
                bar_wait(&full[s], (uint32_t)(j / kStages) & 1u);
                bar_arrive_local(&empty[s]);
                bar_arrive_remote(peer_addr(s32(&empty[s]), peer));
            }
        }
        cycles[blockIdx.x] = clock64() - t0;
    }
    __syncthreads();
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

static double run_mc(const uint8_t* src, long long region, int iters, unsigned long long* d_cyc, int ctas, float* ms_out) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    for (int rep = 0; rep < 2; ++rep) {
        CK(cudaEventRecord(e0));
        feed_probe_mc<<<ctas, 128, kSmem>>>(src, region, iters, d_cyc);
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
    }
    CK(cudaEventElapsedTime(ms_out, e0, e1));
    unsigned long long* h = (unsigned long long*)malloc(sizeof(unsigned long long) * ctas);
    CK(cudaMemcpy(h, d_cyc, sizeof(unsigned long long) * ctas, cudaMemcpyDeviceToHost));
    double worst = 0;
    for (int i = 0; i < ctas; ++i) worst = h[i] > worst ? (double)h[i] : worst;
    free(h);
    return worst;
}

static double run(const uint8_t* src, long long region, int l2_bytes, int peer_bytes, int iters, unsigned long long* d_cyc,
                  int ctas, float* ms_out) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    for (int rep = 0; rep < 2; ++rep) {  // first launch warms L2 and the instruction cache
        CK(cudaEventRecord(e0));
        feed_probe<<<ctas, 128, kSmem>>>(src, region, l2_bytes, peer_bytes, iters, d_cyc);
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
    }
    CK(cudaEventElapsedTime(ms_out, e0, e1));
    unsigned long long* h = (unsigned long long*)malloc(sizeof(unsigned long long) * ctas);
    CK(cudaMemcpy(h, d_cyc, sizeof(unsigned long long) * ctas, cudaMemcpyDeviceToHost));
    double worst = 0;
    for (int i = 0; i < ctas; ++i) worst = h[i] > worst ? (double)h[i] : worst;
    free(h);
    return worst;
}

int main() {
    const int ctas = 148, iters = 4000;
    const long long region = 432 * 1024;  // per CTA; 148 regions = 62 MB: resident in the 126 MB L2 after the first launch
    uint8_t* src;
    unsigned long long* d_cyc;
    CK(cudaMalloc(&src, region * ctas));
    CK(cudaMemset(src, 1, region * ctas));
    CK(cudaMalloc(&d_cyc, sizeof(unsigned long long) * ctas));
    CK(cudaFuncSetAttribute(feed_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
    struct Case { const char* name; int l2, peer; };
    const Case cases[] = {
        {"L2 only, 72 KB per iteration (today's K block)", kMaxL2 + kMaxPeer, 0},  // spans both regions of a stage
        {"L2 only, 48 KB", 49152, 0},
        {"L2 only, 24 KB", 24576, 0},
        {"peer only, 24 KB", 0, 24576},
        {"peer only, 8 KB", 0, 8192},
        {"L2 48 KB + peer 24 KB (the forwarding scheme)", 49152, 24576},
        {"L2 24 KB + peer 24 KB", 24576, 24576},
    };
    printf("%-52s %12s %12s %12s %10s\n", "case (per CTA and iteration)", "cycles/iter", "L2 B/clk/SM", "peer B/clk/SM", "ms");
    for (const Case& c : cases) {
        float ms = 0.f;
        const double cyc = run(src, region, c.l2, c.peer, iters, d_cyc, ctas, &ms);
        const double per = cyc / iters;
        printf("%-52s %12.0f %12.1f %12.1f %10.3f\n", c.name, per, c.l2 / per, c.peer / per, ms);
    }
    // ---- not yet run: the real kernel's traffic pattern (multicast pair) and an HBM-resident source ----
    CK(cudaFuncSetAttribute(feed_probe_mc, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
    const long long big_region = 40LL << 20;  // per CTA; 148 regions = 5.9 GB: every load misses the 126 MB L2
    uint8_t* big;
    CK(cudaMalloc(&big, big_region * ctas));
    CK(cudaMemset(big, 1, big_region * ctas));
    const int it_big = 500;  // 500 x 72 KB = 36 MB per CTA: one pass over its region, no reuse
    {
        float ms = 0.f;
        double per = run_mc(src, region, iters, d_cyc, ctas, &ms) / iters;
        printf("%-52s %12.0f %12.1f %12s %10.3f\n", "pair multicast 48 KB + own 24 KB, L2-resident", per, 73728 / per, "-", ms);
        per = run(big, big_region, kMaxL2 + kMaxPeer, 0, it_big, d_cyc, ctas, &ms) / it_big;
        printf("%-52s %12.0f %12.1f %12s %10.3f\n", "72 KB per iteration, HBM-resident source", per, 73728 / per, "-", ms);
        per = run_mc(big, big_region, it_big, d_cyc, ctas, &ms) / it_big;
        printf("%-52s %12.0f %12.1f %12s %10.3f\n", "pair multicast 48 KB + own 24 KB, HBM-resident", per, 73728 / per, "-", ms);
    }
    printf("(a 64-deep K block of oz_gemm_kernel<6> needs 1344 tensor cycles at full rate; the kernel takes 1644 at config 3)\n");
    return 0;
}
