// EXPERIMENT RECORD -- not compiled into liblcx_b200.so, nothing includes this file.
//
// CTA-pair (tcgen05 cta_group::2) variants of oz_gemm_kernel, written to lower the operand bytes delivered per MAC
// (both contractions sit at the crossbar-to-SM delivery ceiling).  Everything below was validated on a B200
// through the product's own tests with the variant switched in (bit-exact digit-plane test, the 12-shape pass-pair sweep,
// 96 parity / property tests) -- the arithmetic is right -- and measured at config 3 (100 000 x 10 000 x 100, S = 6,
// one pass pair incl. digit slicing, tools/gpu_probe.py):
//
//   single-CTA kernel (the product): wide N-concatenated MMAs, X~ tile multicast to a 64+48 pair      3.66-3.72 ms
//   pair, 21 narrow MMAs (N = 64) per K step, M-side operand from shared memory                        4.27 ms
//   pair, 21 narrow MMAs, X~ planes staged into TMEM by tcgen05.cp.128x256b + A-from-TMEM MMAs         4.76 ms
//   pair, 12 MMAs of two planes each (N = 128, zero plane completes the pairs), operands from smem     4.26 ms
//
// What the numbers say: a cta_group::2 kind::i8 MMA costs about 37 + 0.41 N cycles (63 at N = 64, 89 at N = 128; the
// work is N/2 cycles), so only N = 256 instructions approach full rate -- and a fixed plane count per instruction is the
// only N-concatenation that composes with the pair's split of N (the half boundary moves with the plane count
// otherwise), which at four planes wastes half of the slots.  tcgen05.cp is not free either: ~40 cycles per 4 KB slab,
// serialised with the MMAs on the same pipe, for a reuse of only 3.5 MMAs per slab.  So the 15 % fewer delivered bytes do
// not pay at S = 6, m = 100; the single-CTA wide-instruction kernel stays.  (tcgen05.cp + A-from-TMEM itself works with
// the same SW64 K-major descriptor the MMA takes: tools/experiments/tmem_a_probe.cu.)
//
// The fragments assume the context of linearcorex_b200/csrc/ozaki_i8.cuh (namespace lcx::oz, its PTX wrappers, GemmParams).
#if 0
// ---- CTA-pair (cta_group::2) forms: one MMA spans two SMs (M = 256), each holding half of the N-side operand ----
__device__ __forceinline__ void tmem_alloc2(uint32_t* dst_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols));
}
// D[tmem, both CTAs] (+)= A[smem, 128 rows per CTA] * B[smem, half of the N rows per CTA]^T; issued by the leader CTA only
__device__ __forceinline__ void umma2_i8(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate));
}
__device__ __forceinline__ void umma2_commit_mc(uint64_t* bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                     smem_u32(bar)),
                 "h"(mask)
                 : "memory");
}
// arrive on the same-named barrier of another CTA of the cluster (release at cluster scope)
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t cta_rank) {
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(bar)), "r"(cta_rank));
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP_C:\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_C;\n\t"
        "bra WAIT_LOOP_C;\n\t"
        "DONE_C:\n\t"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
}

// ---- first attempt: X~ planes through tensor memory ----
// D[tmem] (+)= A[tmem] * B[smem]^T: the M-side operand comes from tensor memory (staged there by tcgen05.cp)
__device__ __forceinline__ void umma2_i8_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::i8 [%0], [%1], %2, %3, p;\n\t"
        "}\n" ::"r"(d_tmem),
        "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate));
}
// 128 rows x 32 bytes of a K-major shared-memory tile -> 128 lanes x 8 columns of tensor memory, in both CTAs of the pair
__device__ __forceinline__ void tmem_cp2_128x256b(uint32_t taddr, uint64_t s_desc) {
    asm volatile("tcgen05.cp.cta_group::2.128x256b [%0], %1;" ::"r"(taddr), "l"(s_desc));
}
template <int S, int BN, bool A_TMEM>
__device__ __forceinline__ void issue_kblock_pair(uint32_t sa, uint32_t sb, uint32_t d_tmem, uint32_t a_tmem, bool first_block) {
    constexpr int A_BYTES = kBM * kBK;
    constexpr int BH_BYTES = (BN / 2) * kBK;  // this CTA's half of one factor plane
    constexpr uint32_t idesc = make_idesc_i8(2 * kBM, BN, 0, 0);
#pragma unroll
    for (int kk = 0; kk < kBK / 32; ++kk) {
        const uint32_t abuf = a_tmem + (uint32_t)(kk * S * 8);  // K steps alternate between the two staging buffers
        if (A_TMEM) {
#pragma unroll
            for (int ka = 0; ka < S; ++ka)
                tmem_cp2_128x256b(abuf + (uint32_t)(ka * 8), make_smem_desc(sa + ka * A_BYTES + kk * 32, 16, 512, 4));
        }
#pragma unroll
        for (int ka = 0; ka < S; ++ka) {
            const uint64_t da = make_smem_desc(sa + ka * A_BYTES + kk * 32, 16, 512, 4);
#pragma unroll
            for (int l = 0; l < S - ka; ++l) {
                const uint64_t db = make_smem_desc(sb + l * BH_BYTES + kk * 32, 16, 512, 4);
                const uint32_t acc = (!first_block || kk > 0 || ka > 0) ? 1u : 0u;
                if (A_TMEM) umma2_i8_ts(d_tmem + (uint32_t)((ka + l) * BN), abuf + (uint32_t)(ka * 8), db, idesc, acc);
                else umma2_i8_ss(d_tmem + (uint32_t)((ka + l) * BN), da, db, idesc, acc);
            }
        }
    }
}


// ---- CTA-pair variant (S = 6) ----------------------------------------------------------------------------------------
// Both contractions run at the crossbar-to-SM delivery ceiling, so what counts is bytes delivered per MAC.
// Two SMs form one tcgen05 CTA pair (cta_group::2, M = 256): each CTA holds its own 128 rows of the X~ tile and only
// HALF of the factor planes (the tensor cores of both SMs read both halves), so a pair takes 2 x 48 KB + 24 KB per 64-deep
// K block for 256 x 64 outputs instead of 2 x (48 + 24) KB for 128 x (64 + 48): -15 % at m = 100, and the 64- and 48-wide
// factor tiles no longer run in lock-step.
// Instruction shape: a kind::i8 MMA narrower than N = 128 runs at about half rate (measured: ~63 cycles for N = 64), and
// the single-CTA kernel's N-concatenation over S-k planes does not compose with the pair's N split (the half boundary
// would move with the plane count).  A FIXED count of two planes per instruction does: digit k of X~ meets the plane pairs
// (l, l+1) with k+l even, and a zero plane in front of plane 0 completes the pairs for odd k -- 12 instructions of
// N = 2 bn per K step for the 21 products (7/8 useful).  TMEM holds three blocks of 2 bn columns, block p = groups
// (2p, 2p+1) laid out [g0 half0 | g1 half0 | g0 half1 | g1 half1] (half h = the factor rows CTA h supplies).
// Only the leader CTA issues MMAs and commits; the peer's TMA completion reaches it through a remote mbarrier arrive.
template <bool KMAJOR, int BN>
__device__ __forceinline__ void issue_kblock_pair(uint32_t sa, uint32_t sbz, uint32_t tmem_base, bool first_block) {
    constexpr int S = 6;
    constexpr int A_BYTES = kBM * kBK;
    constexpr int BH_BYTES = (BN / 2) * kBK;  // this CTA's half of one factor plane; the zero plane sits at sbz
    constexpr uint32_t idesc = make_idesc_i8(2 * kBM, 2 * BN, KMAJOR ? 0 : 1, 0);
#pragma unroll
    for (int kk = 0; kk < kBK / 32; ++kk) {
#pragma unroll
        for (int ka = 0; ka < S; ++ka) {
            const uint64_t da = KMAJOR ? make_smem_desc(sa + ka * A_BYTES + kk * 32, 16, 512, 4)
                                       : make_smem_desc(sa + ka * A_BYTES + kk * 4096, 8192, 1024, 2);
#pragma unroll
            for (int l = -(ka & 1); l + 1 <= S - 1 - ka; l += 2) {   // planes (l, l+1); l = -1 is the zero plane
                const uint64_t db = make_smem_desc(sbz + (l + 1) * BH_BYTES + kk * 32, 16, 512, 4);
                const uint32_t acc = (!first_block || kk > 0 || ka > 0) ? 1u : 0u;
                umma2_i8(tmem_base + (uint32_t)(((ka + l) / 2) * 2 * BN), da, db, idesc, acc);
            }
        }
    }
}

template <bool KMAJOR>
__global__ void __launch_bounds__(kThreads, 1)
oz_gemm_pair_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapBh,
                    const __grid_constant__ CUtensorMap mapBth, const GemmParams p) {
    constexpr int S = 6;
    constexpr int kBN = 64;
    constexpr int kStages = 3;
    constexpr int A_BYTES = kBM * kBK;            // 8 KB per plane
    constexpr int BH_BYTES = (kBN / 2) * kBK;     // 2 KB per plane: half of a full-width factor tile
    constexpr int STAGE_BYTES = S * A_BYTES + (S + 1) * BH_BYTES + 1024;  // + zero plane; keeps stages 1 KB aligned
    constexpr uint32_t TMEM_COLS = 512;           // 3 blocks of 2 x 64 columns

    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(8) uint64_t full_bar[kStages];       // this CTA's TMA landed
    __shared__ __align__(8) uint64_t peer_full_bar[kStages];  // leader only: the peer's TMA landed (remote arrive)
    __shared__ __align__(8) uint64_t empty_bar[kStages];      // the pair's MMAs finished reading the stage (leader's commit)
    __shared__ __align__(8) uint64_t tmem_full_bar;
    __shared__ uint32_t tmem_base_smem;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t crank = cluster_ctarank();
    const bool leader = crank == 0;
    const int n_tile = blockIdx.x >> 1, m_pair = blockIdx.y;
    const int row0 = m_pair * 2 * kBM + (int)crank * kBM;
    const int kbeg = blockIdx.z * p.k_chunk;
    const int kend = min(p.k_total, kbeg + p.k_chunk);
    const int num_kb = (kend > kbeg) ? (kend - kbeg + kBK - 1) / kBK : 0;
    const bool tail = p.bn_tail > 0 && p.bn_tail < kBN && n_tile == p.n_tiles - 1;
    const int bn = tail ? p.bn_tail : kBN;
    const int bh = bn / 2;                        // factor rows this CTA supplies
    const int bh_bytes = bh * kBK;

    // the zero plane of every stage (never touched by TMA)
    for (int st = 0; st < kStages; ++st)
        for (int i = threadIdx.x * 16; i < BH_BYTES; i += kThreads * 16)
            *reinterpret_cast<uint4*>(smem + st * STAGE_BYTES + S * A_BYTES + i) = make_uint4(0u, 0u, 0u, 0u);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapBh) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapBth) : "memory");
        for (int i = 0; i < kStages; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&peer_full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        mbar_init(&tmem_full_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc2(&tmem_base_smem, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;

    if (warp == 0) {
        // ===== TMA producer (both CTAs: own 128 rows of X~, own half of the factor planes) =====
        if (lane == 0) {
            for (int kb = 0; kb < num_kb; ++kb) {
                const int st = kb % kStages;
                const uint32_t ph = (kb / kStages) & 1;
                mbar_wait_cluster(&empty_bar[st], ph ^ 1);
                uint8_t* sa = smem + st * STAGE_BYTES;
                uint8_t* sb = sa + S * A_BYTES + bh_bytes;  // plane 0 follows the zero plane at the tile's own pitch
                mbar_expect_tx(&full_bar[st], S * (A_BYTES + bh_bytes));
                const int k0 = kbeg + kb * kBK;
#pragma unroll
                for (int s = 0; s < S; ++s) {
                    if (KMAJOR) tma_load_3d(sa + s * A_BYTES, &mapA, &full_bar[st], k0, row0, s);
                    else tma_load_3d(sa + s * A_BYTES, &mapA, &full_bar[st], row0, k0, s);
                    tma_load_3d(sb + s * bh_bytes, tail ? &mapBth : &mapBh, &full_bar[st], k0, n_tile * kBN + (int)crank * bh, s);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            if (leader) {
                // ===== MMA issuer for the pair =====
                for (int kb = 0; kb < num_kb; ++kb) {
                    const int st = kb % kStages;
                    const uint32_t ph = (kb / kStages) & 1;
                    mbar_wait(&full_bar[st], ph);
                    mbar_wait_cluster(&peer_full_bar[st], ph);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + st * STAGE_BYTES);
                    const uint32_t sbz = sa + S * A_BYTES;
                    if (bn == 64) issue_kblock_pair<KMAJOR, 64>(sa, sbz, tmem_base, kb == 0);
                    else if (bn == 48) issue_kblock_pair<KMAJOR, 48>(sa, sbz, tmem_base, kb == 0);
                    else issue_kblock_pair<KMAJOR, 32>(sa, sbz, tmem_base, kb == 0);
                    umma2_commit_mc(&empty_bar[st], (uint16_t)3);  // frees the stage in both CTAs
                }
                umma2_commit_mc(&tmem_full_bar, (uint16_t)3);
            } else {
                // ===== relay: tell the leader when this CTA's half of a stage has landed =====
                for (int kb = 0; kb < num_kb; ++kb) {
                    const int st = kb % kStages;
                    const uint32_t ph = (kb / kStages) & 1;
                    mbar_wait(&full_bar[st], ph);
                    mbar_arrive_remote(&peer_full_bar[st], 0u);
                }
            }
        }
    } else {
        // ===== epilogue: this CTA's 128 rows, TMEM -> registers -> fp64 recombination -> global =====
        const int quarter = warp & 3;
        const int row = row0 + quarter * 32 + lane;
        mbar_wait_cluster(&tmem_full_bar, 0);
        tc_fence_after();
        double* C = p.C + (long long)blockIdx.z * p.c_split_stride;
        const double rs = (p.row_scale != nullptr && row < p.rows) ? p.row_scale[row] : 1.0;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
#pragma unroll 1
        for (int c0 = 0; c0 < bn; c0 += 8) {   // 8-column chunks never straddle the half boundary (bh = 16, 24 or 32)
            const int half = c0 / bh, jp = c0 - half * bh;
            double acc[8];
            if (num_kb > 0) {
                uint32_t r[8];
#pragma unroll
                for (int g = S - 1; g >= 0; --g) {
                    tmem_ld8(lane_addr + (uint32_t)((g >> 1) * 2 * bn + half * bn + (g & 1) * bh + jp), r);
                    tmem_ld_wait();
                    if (g == S - 1) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) acc[j] = (double)(int)r[j];
                    } else {
#pragma unroll
                        for (int j = 0; j < 8; ++j) acc[j] = acc[j] * p.inv_radix + (double)(int)r[j];
                    }
                }
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[j] = 0.0;
            }
            if (row < p.rows) {
                const int col0 = n_tile * kBN + c0;
#pragma unroll
                for (int j = 0; j < 8; j += 2) {
                    const int col = col0 + j;
                    double v0 = acc[j] * (p.inv_radix * p.inv_radix) * rs, v1 = acc[j + 1] * (p.inv_radix * p.inv_radix) * rs;
                    if (p.col_scale != nullptr) {
                        if (col < p.cols) v0 *= p.col_scale[col];
                        if (col + 1 < p.cols) v1 *= p.col_scale[col + 1];
                    }
                    if (p.trans_out) {
                        if (col < p.cols) C[(long long)col * p.ldc + row] = v0;
                        if (col + 1 < p.cols) C[(long long)(col + 1) * p.ldc + row] = v1;
                    } else if (col + 1 < p.cols) {
                        *reinterpret_cast<double2*>(C + (long long)row * p.ldc + col) = make_double2(v0, v1);
                    } else if (col < p.cols) {
                        C[(long long)row * p.ldc + col] = v0;
                    }
                }
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    cluster_sync_all();  // nobody leaves (or frees tensor memory) while the pair's MMAs or remote arrives can still touch it
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc2(tmem_base, TMEM_COLS);
    }
}

// CTA-pair launch: grid.x = 2 CTAs per factor tile, grid.y = 256-row tiles, grid.z = K splits
template <bool KMAJOR>
inline int launch_oz_gemm_pair(const CUtensorMap& mapA, const CUtensorMap& mapBh, const CUtensorMap& mapBth, GemmParams p,
                               int n_tiles, int m_pairs, int splits, cudaStream_t st) {
    constexpr int SMEM = 3 * (6 * kBM * kBK + 7 * 32 * kBK + 1024) + 1024;
    static bool configured = false;
    auto kern = oz_gemm_pair_kernel<KMAJOR>;
    if (!configured) {
        LCX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
        configured = true;
    }
    p.n_tiles = n_tiles;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)(2 * n_tiles), (unsigned)m_pairs, (unsigned)splits);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = SMEM;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    LCX_CUDA(cudaLaunchKernelEx(&cfg, kern, mapA, mapBh, mapBth, p));
    return 0;
}

#endif
