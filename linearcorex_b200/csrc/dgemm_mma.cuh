// FP64 tensor-core GEMM for the Linear CorEx contractions (FP64-faithful mode).
//
// Every dense contraction of the fit loop is an instance of this kernel:
//   Y   = X~ A^T            linearcorex.py:247 / :210       (A K-contiguous, B K-contiguous)
//   P   = X~^T Y            linearcorex.py:259 / :211       (A M-contiguous, B N-contiguous, split-K)
//   ry  = W rho^T, H, cy    :261 / :294 / :355              (both K-contiguous, split-K over n)
//   Qij = ry rhoinvrho, H W :266 / :300                     (A K-contiguous, B N-contiguous)
//   cov = z^T z             :448                            (A M-contiguous, B N-contiguous)
//
// Shape of one CTA: 128 x (8*NT) output tile, 8 warps stacked along M (16 rows each, the full tile
// width), K consumed in 16-wide slabs through a STAGES-deep cp.async ring.  The inner product is
// DMMA.8x8x4 (mma.sync.m8n8k4.f64) -- on sm_100a every PTX f64 mma shape lowers to that SASS
// instruction, so it is issued directly.  tcgen05 has no f64 kind; the tcgen05/TMEM/TMA path is the
// split-integer engine of ozaki_i8.cuh, which replaces this kernel for the two X contractions in the
// default precision mode.
//
// Shared-memory tiles are padded so that the 8-byte fragment loads of a half-warp hit 16 distinct
// bank pairs:  K-contiguous tiles use a row stride of 20 doubles, M/N-contiguous tiles a stride
// congruent to 4 or 12 (mod 16).
//
// Edges are handled by zero-filling cp.async (src-size operand), so callers need no padding
// guarantees beyond: base pointers 16-byte aligned and leading dimensions even.
#pragma once
#include "common.cuh"

namespace lcx {

struct GemmArgs {
    const double* A;
    const double* B;
    double* C;
    const double* Cadd;      // optional, same layout as C: C = A*B + Cadd
    double* colsq_part;      // optional: [gridDim.x][ld_colsq] per-CTA column sums of squares of A*B
    int M, N, K;
    long long lda, ldb, ldc;
    int k_chunk;             // K range handled by one blockIdx.z (multiple of 16)
    long long c_split_stride;
    int trans_out;           // store C^T (element (r,c) at C[c*ldc + r])
    int ld_colsq;
};

template <int NT, bool A_KC, bool B_KC, int STAGES>
struct GemmCfg {
    static constexpr int BM = 128, BN = NT * 8, BK = 16;
    static constexpr int A_STRIDE = A_KC ? (BK + 4) : (BM + 4);
    static constexpr int A_ROWS = A_KC ? BM : BK;
    static constexpr int B_PAD = (BN % 16 == 0) ? 4 : 12;
    static constexpr int B_STRIDE = B_KC ? (BK + 4) : (BN + B_PAD);
    static constexpr int B_ROWS = B_KC ? BN : BK;
    static constexpr int A_TILE = A_ROWS * A_STRIDE;
    static constexpr int B_TILE = B_ROWS * B_STRIDE;
    static constexpr int SMEM_BYTES = STAGES * (A_TILE + B_TILE) * 8;
};

template <int NT, bool A_KC, bool B_KC, int STAGES>
__global__ void __launch_bounds__(256, 1) dgemm_mma_kernel(const GemmArgs p) {
    using Cfg = GemmCfg<NT, A_KC, B_KC, STAGES>;
    constexpr int BM = Cfg::BM, BN = Cfg::BN, BK = Cfg::BK;
    constexpr int A_STRIDE = Cfg::A_STRIDE, B_STRIDE = Cfg::B_STRIDE;
    constexpr int A_TILE = Cfg::A_TILE, B_TILE = Cfg::B_TILE;

    extern __shared__ __align__(16) double smem[];
    double* As = smem;
    double* Bs = smem + STAGES * A_TILE;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    const int kbeg = blockIdx.z * p.k_chunk;
    const int kend = min(p.K, kbeg + p.k_chunk);
    const int KT = (kend > kbeg) ? (kend - kbeg + BK - 1) / BK : 0;

    auto load_tile = [&](int stage, int kt) {
        const int k0 = kbeg + kt * BK;
        double* as = As + stage * A_TILE;
        double* bs = Bs + stage * B_TILE;
        if (A_KC) {  // 128 rows x 8 chunks of 2 doubles
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int idx = tid + i * 256;
                const int r = idx >> 3, c = idx & 7;
                const int row = m0 + r, k = k0 + 2 * c;
                const int valid = (row < p.M) ? max(0, min(2, kend - k)) : 0;
                const double* src = valid ? p.A + (long long)row * p.lda + k : p.A;
                cp_async_16(smem_u32(as + r * A_STRIDE + 2 * c), src, valid * 8);
            }
        } else {  // 16 k-rows x 64 chunks
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int idx = tid + i * 256;
                const int r = idx >> 6, c = idx & 63;
                const int k = k0 + r, col = m0 + 2 * c;
                const int valid = (k < kend) ? max(0, min(2, p.M - col)) : 0;
                const double* src = valid ? p.A + (long long)k * p.lda + col : p.A;
                cp_async_16(smem_u32(as + r * A_STRIDE + 2 * c), src, valid * 8);
            }
        }
        if (B_KC) {  // BN rows x 8 chunks
            constexpr int TOTAL = BN * 8;
#pragma unroll
            for (int i = 0; i < (TOTAL + 255) / 256; ++i) {
                const int idx = tid + i * 256;
                if (idx < TOTAL) {
                    const int r = idx >> 3, c = idx & 7;
                    const int row = n0 + r, k = k0 + 2 * c;
                    const int valid = (row < p.N) ? max(0, min(2, kend - k)) : 0;
                    const double* src = valid ? p.B + (long long)row * p.ldb + k : p.B;
                    cp_async_16(smem_u32(bs + r * B_STRIDE + 2 * c), src, valid * 8);
                }
            }
        } else {  // 16 k-rows x (BN/2) chunks
            constexpr int CPR = BN / 2, TOTAL = 16 * CPR;
#pragma unroll
            for (int i = 0; i < (TOTAL + 255) / 256; ++i) {
                const int idx = tid + i * 256;
                if (idx < TOTAL) {
                    const int r = idx / CPR, c = idx % CPR;
                    const int k = k0 + r, col = n0 + 2 * c;
                    const int valid = (k < kend) ? max(0, min(2, p.N - col)) : 0;
                    const double* src = valid ? p.B + (long long)k * p.ldb + col : p.B;
                    cp_async_16(smem_u32(bs + r * B_STRIDE + 2 * c), src, valid * 8);
                }
            }
        }
    };

    double acc[2][NT][2];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;

    // Row groups of 8 that lie wholly beyond M issue no DMMA (warp-uniform): with m = 100 factors on the M side only 13 of the
    // 16 groups of a tile do work, and the FP64 tensor pipe -- one per SM, shared by the 8 warps -- is what bounds these loops.
    const bool act0 = m0 + warp * 16 < p.M, act1 = m0 + warp * 16 + 8 < p.M;
    auto compute = [&](int stage) {
        const double* as = As + stage * A_TILE;
        const double* bs = Bs + stage * B_TILE;
        if (!act0) return;
#pragma unroll
        for (int ks = 0; ks < BK / 4; ++ks) {
            const int kk = ks * 4 + t;
            double a0, a1;
            if (A_KC) {
                a0 = as[(warp * 16 + g) * A_STRIDE + kk];
                a1 = as[(warp * 16 + 8 + g) * A_STRIDE + kk];
            } else {
                a0 = as[kk * A_STRIDE + warp * 16 + g];
                a1 = as[kk * A_STRIDE + warp * 16 + 8 + g];
            }
            if (act1) {
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) {
                    const double b = B_KC ? bs[(nt * 8 + g) * B_STRIDE + kk] : bs[kk * B_STRIDE + nt * 8 + g];
                    dmma884(acc[0][nt][0], acc[0][nt][1], a0, b);
                    dmma884(acc[1][nt][0], acc[1][nt][1], a1, b);
                }
            } else {
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) {
                    const double b = B_KC ? bs[(nt * 8 + g) * B_STRIDE + kk] : bs[kk * B_STRIDE + nt * 8 + g];
                    dmma884(acc[0][nt][0], acc[0][nt][1], a0, b);
                }
            }
        }
    };

#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < KT) load_tile(s, s);
        cp_async_commit();
    }
    for (int kt = 0; kt < KT; ++kt) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        const int nk = kt + STAGES - 1;
        if (nk < KT) load_tile(nk % STAGES, nk);
        cp_async_commit();
        compute(kt % STAGES);
    }
    cp_async_wait<0>();
    __syncthreads();

    // ---- epilogue -----------------------------------------------------------------------------
    double* Cz = p.C + (long long)blockIdx.z * p.c_split_stride;
    if (p.C != nullptr) {
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
            const int row = m0 + warp * 16 + mt * 8 + g;
            if (row < p.M) {
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) {
                    const int col = n0 + nt * 8 + 2 * t;
                    double v0 = acc[mt][nt][0], v1 = acc[mt][nt][1];
                    if (!p.trans_out) {
                        const long long off = (long long)row * p.ldc + col;
                        if (col + 1 < p.N) {
                            if (p.Cadd) {
                                const double2 c = *reinterpret_cast<const double2*>(p.Cadd + off);
                                v0 += c.x;
                                v1 += c.y;
                            }
                            *reinterpret_cast<double2*>(Cz + off) = make_double2(v0, v1);
                        } else if (col < p.N) {
                            if (p.Cadd) v0 += p.Cadd[off];
                            Cz[off] = v0;
                        }
                    } else {
                        if (col < p.N) {
                            const long long off = (long long)col * p.ldc + row;
                            Cz[off] = p.Cadd ? v0 + p.Cadd[off] : v0;
                        }
                        if (col + 1 < p.N) {
                            const long long off = (long long)(col + 1) * p.ldc + row;
                            Cz[off] = p.Cadd ? v1 + p.Cadd[off] : v1;
                        }
                    }
                }
            }
        }
    }
    if (p.colsq_part != nullptr) {  // sum_l Y_lj^2 for this CTA's 128 rows (linearcorex.py:248)
        double* red = smem;         // [8 warps][BN]; the pipeline buffers are idle now
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
            double s0 = acc[0][nt][0] * acc[0][nt][0] + acc[1][nt][0] * acc[1][nt][0];
            double s1 = acc[0][nt][1] * acc[0][nt][1] + acc[1][nt][1] * acc[1][nt][1];
#pragma unroll
            for (int o = 4; o < 32; o <<= 1) {
                s0 += __shfl_xor_sync(0xffffffffu, s0, o);
                s1 += __shfl_xor_sync(0xffffffffu, s1, o);
            }
            if (g == 0) {
                red[warp * BN + nt * 8 + 2 * t] = s0;
                red[warp * BN + nt * 8 + 2 * t + 1] = s1;
            }
        }
        __syncthreads();
        for (int c = tid; c < BN; c += 256) {
            double s = 0.0;
#pragma unroll
            for (int w = 0; w < 8; ++w) s += red[w * BN + c];
            if (n0 + c < p.N) p.colsq_part[(long long)blockIdx.x * p.ld_colsq + n0 + c] = s;
        }
    }
}

// Fixed-order reduction of split-K partials: out = sum_z part[z]  (deterministic: the association order depends only on
// LANES, never on timing).  Only the valid rows x cols region (leading dimension ld, even) is read or written.
// LANES threads share one output pair: lane q adds splits q, q+LANES, ... in order, then the lanes are combined by a
// fixed shuffle tree.  LANES = 8 is used for small outputs with many splits (the m x m products), 1 otherwise.
template <int LANES>
__global__ void __launch_bounds__(256) reduce_splits_kernel(const double* __restrict__ part, int splits, long long stride,
                                                            double* __restrict__ out, int rows, int cpairs, int cols,
                                                            long long ld, double diag_value, double* __restrict__ diag_out) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long p = t / LANES;
    const int q = (int)(t % LANES);
    const bool live = p < (long long)rows * cpairs;
    const int r = live ? (int)(p / cpairs) : 0;
    const int c2 = live ? (int)(p % cpairs) * 2 : 0;
    const long long o = (long long)r * ld + c2;
    const bool two = c2 + 1 < cols;
    double ax = 0.0, ay = 0.0;
    if (live) {
        int z = q;
        for (; z + 3 * LANES < splits; z += 4 * LANES) {  // four independent loads in flight
            const double* b0 = part + (long long)z * stride + o;
            const double* b1 = b0 + (long long)LANES * stride;
            const double* b2 = b1 + (long long)LANES * stride;
            const double* b3 = b2 + (long long)LANES * stride;
            if (two) {
                const double2 v0 = *reinterpret_cast<const double2*>(b0), v1 = *reinterpret_cast<const double2*>(b1);
                const double2 v2 = *reinterpret_cast<const double2*>(b2), v3 = *reinterpret_cast<const double2*>(b3);
                ax += v0.x; ay += v0.y; ax += v1.x; ay += v1.y; ax += v2.x; ay += v2.y; ax += v3.x; ay += v3.y;
            } else {
                const double v0 = *b0, v1 = *b1, v2 = *b2, v3 = *b3;
                ax += v0; ax += v1; ax += v2; ax += v3;
            }
        }
        for (; z < splits; z += LANES) {
            const double* b0 = part + (long long)z * stride + o;
            if (two) {
                const double2 v = *reinterpret_cast<const double2*>(b0);
                ax += v.x; ay += v.y;
            } else {
                ax += *b0;
            }
        }
    }
    if (LANES > 1) {
#pragma unroll
        for (int w = LANES / 2; w > 0; w >>= 1) {
            ax += __shfl_down_sync(0xffffffffu, ax, w, LANES);
            ay += __shfl_down_sync(0xffffffffu, ay, w, LANES);
        }
    }
    if (live && q == 0) {
        // optional np.fill_diagonal fused into the combine (square outputs): keep the raw diagonal, store diag_value
        if (diag_out != nullptr) {
            if (c2 == r) { diag_out[r] = ax; ax = diag_value; }
            if (two && c2 + 1 == r) { diag_out[r] = ay; ay = diag_value; }
        }
        if (two) *reinterpret_cast<double2*>(out + o) = make_double2(ax, ay);
        else out[o] = ax;
    }
}

// ---- host-side launcher -----------------------------------------------------------------------
enum GemmLayout { kLayoutKK = 0, kLayoutMN = 1, kLayoutKN = 2 };  // (A,B): (K,K) (M,N) (K,N)-contiguous

struct GemmPlan {
    int nt;       // n8 tiles per CTA
    int splits;   // gridDim.z
    int k_chunk;  // multiple of 16
    dim3 grid;
};

inline int pick_nt(int N) {
    const int need = (N + 7) / 8;
    const int opts[5] = {2, 4, 8, 13, 16};
    for (int i = 0; i < 5; ++i)
        if (opts[i] >= need) return opts[i];
    // wider than 128: prefer the option with the least padding
    int best = 16;
    long long best_cols = (long long)cdiv(N, 128) * 128;
    const long long c13 = (long long)cdiv(N, 104) * 104;
    if (c13 < best_cols) { best = 13; best_cols = c13; }
    return best;
}

// Choose the split count that best fills whole waves of `sms` CTAs (one CTA per SM).
inline GemmPlan plan_gemm(int M, int N, int K, int sms, int max_splits, bool allow_split) {
    GemmPlan pl;
    pl.nt = pick_nt(N);
    const int gm = cdiv(M, 128), gn = cdiv(N, pl.nt * 8);
    const long long tiles = (long long)gm * gn;
    const int ktiles = cdiv(K, 16);
    int best_s = 1;
    if (allow_split && max_splits > 1) {
        double best_score = -1.0;
        const int smax = (int)min((long long)max_splits, (long long)max(1, ktiles / 4));
        for (int s = 1; s <= smax; ++s) {
            const long long ctas = tiles * s;
            const long long waves = (ctas + sms - 1) / sms;
            double eff = (double)ctas / (double)(waves * sms);
            // each split pays a pipeline fill (~4 k-tiles) and a partial write/read
            const double kt_per = (double)ktiles / s;
            eff *= kt_per / (kt_per + 6.0);
            if (eff > best_score + 1e-9) { best_score = eff; best_s = s; }
        }
    }
    const int kt_per = cdiv(ktiles, best_s);
    pl.k_chunk = kt_per * 16;
    pl.splits = cdiv(K, pl.k_chunk);
    if (pl.splits < 1) pl.splits = 1;
    pl.grid = dim3(gm, gn, pl.splits);
    return pl;
}

// (m x n) = (m x m)(m x n) products (Qij = ry rhoinvrho, grad = G0 + H W): K = m is a handful of k tiles and the variables run
// along N, so the tile width decides how many SMs work at all -- 104-wide tiles of n = 10 000 are 97 CTAs on 148 SMs.  Pick
// the width whose waves x (width + a per-CTA fixed cost of ~3 n8 tiles) is smallest: 72 wide -> 139 CTAs at that shape.
inline GemmPlan plan_gemm_kn(int M, int N, int K, int sms) {
    GemmPlan pl = plan_gemm(M, N, K, sms, 1, false);
    const int opts[9] = {2, 4, 6, 8, 9, 11, 13, 16, 17};
    double best = 1e300;
    const int gm = cdiv(M, 128);
    for (int i = 0; i < 9; ++i) {
        const int nt = opts[i];
        if (nt * 8 > round_up(N, 8) && nt != pick_nt(N)) continue;
        const long long tiles = (long long)gm * cdiv(N, nt * 8);
        const long long waves = (tiles + sms - 1) / sms;
        const double cost = (double)waves * (nt + 3.0);
        if (cost < best - 1e-9) { best = cost; pl.nt = nt; }
    }
    pl.grid = dim3(gm, cdiv(N, pl.nt * 8), 1);
    return pl;
}

template <int NT, bool A_KC, bool B_KC>
inline int launch_gemm_inst(const GemmArgs& a, dim3 grid, cudaStream_t st) {
    constexpr int STAGES = 4;
    using Cfg = GemmCfg<NT, A_KC, B_KC, STAGES>;
    static PerDeviceOnce configured = {};
    auto kern = dgemm_mma_kernel<NT, A_KC, B_KC, STAGES>;
    if (configured.first_time())
        LCX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    kern<<<grid, 256, Cfg::SMEM_BYTES, st>>>(a);
    LCX_CUDA(cudaGetLastError());
    return 0;
}

template <bool A_KC, bool B_KC>
inline int launch_gemm_nt(int nt, const GemmArgs& a, dim3 grid, cudaStream_t st) {
    switch (nt) {
        case 2: return launch_gemm_inst<2, A_KC, B_KC>(a, grid, st);
        case 4: return launch_gemm_inst<4, A_KC, B_KC>(a, grid, st);
        case 8: return launch_gemm_inst<8, A_KC, B_KC>(a, grid, st);
        case 13: return launch_gemm_inst<13, A_KC, B_KC>(a, grid, st);
        case 16: return launch_gemm_inst<16, A_KC, B_KC>(a, grid, st);
    }
    if (A_KC && !B_KC) {  // the extra widths of plan_gemm_kn exist for the (K, N)-contiguous layout only
        switch (nt) {
            case 6: return launch_gemm_inst<6, true, false>(a, grid, st);
            case 9: return launch_gemm_inst<9, true, false>(a, grid, st);
            case 11: return launch_gemm_inst<11, true, false>(a, grid, st);
            case 17: return launch_gemm_inst<17, true, false>(a, grid, st);
        }
    }
    return fail(-1, "launch_gemm", "unsupported tile width");
}

inline int launch_gemm(GemmLayout lay, const GemmPlan& pl, GemmArgs a, cudaStream_t st) {
    LCX_REQUIRE(((uintptr_t)a.A % 16 == 0) && ((uintptr_t)a.B % 16 == 0) && (a.C == nullptr || (uintptr_t)a.C % 16 == 0),
                "GEMM operands must be 16-byte aligned");
    LCX_REQUIRE(a.lda % 2 == 0 && a.ldb % 2 == 0 && a.ldc % 2 == 0, "GEMM leading dimensions must be even");
    a.k_chunk = pl.k_chunk;
    if (a.M <= 0 || a.N <= 0) return 0;
    switch (lay) {
        case kLayoutKK: return launch_gemm_nt<true, true>(pl.nt, a, pl.grid, st);
        case kLayoutMN: return launch_gemm_nt<false, false>(pl.nt, a, pl.grid, st);
        case kLayoutKN: return launch_gemm_nt<true, false>(pl.nt, a, pl.grid, st);
    }
    return fail(-1, "launch_gemm", "unknown layout");
}

inline int launch_reduce_splits(const double* part, int splits, long long stride, double* out, int rows, int cols,
                                long long ld, cudaStream_t st, double* diag_out = nullptr, double diag_value = 0.0) {
    const int cpairs = (cols + 1) / 2;
    const long long outputs = (long long)rows * cpairs;
    if (outputs <= 0) return 0;
    if (outputs <= 65536 && splits >= 16) {
        reduce_splits_kernel<8><<<cdiv(outputs * 8, 256), 256, 0, st>>>(part, splits, stride, out, rows, cpairs, cols, ld, diag_value,
                                                                      diag_out);
    } else {
        reduce_splits_kernel<1><<<cdiv(outputs, 256), 256, 0, st>>>(part, splits, stride, out, rows, cpairs, cols, ld, diag_value,
                                                                  diag_out);
    }
    LCX_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace lcx
