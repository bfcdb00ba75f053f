// Fused m x n phase of a fit iteration for m <= 128 factors (FP64).
//
// Between two passes over the data an iteration of _update_ns (linearcorex.py:290-334) is a chain of m x n elementwise
// passes, per-variable reductions and four skinny products (ry = W rho^T, Qij = ry rinv, H = T rinv^T, grad = G0 + H W:
// 2 m^2 n flops each).  As separate launches (corex_kernels.cuh + dgemm_mma.cuh) that chain is ~15 kernels of 5-30 us whose
// cost is launch/ramp latency, not bytes or flops.  Here the products are folded into the elementwise kernels that produce
// or consume their operands -- one CTA keeps a strip of variables (all m factors) in shared memory and runs the small
// product with DMMA.8x8x4 next to the elementwise work:
//
//   strip_outer_kernel   rho / invrho / rinv / Si (and W + eta U) of a strip, then the strip's share of  A_strip B_strip^T
//                        (ry = W rho^T: :261; or, MODE 0, H = T rinv^T: :294) into a per-CTA partial; a fixed-order combine
//                        (reduce_splits_kernel, with np.fill_diagonal) follows.
//   strip_apply_kernel   Q_strip = M V_strip for the m x m matrix M held in shared memory, then
//                        EPI 0: Qij = ry rinv (:266), Qi - Si^2 (:269), the objective's log sums (:272-274; the last CTA
//                               finishes TC, uj and max uj), and -- for the set just evaluated -- T and G0 of the NEXT search
//                               direction (:294-299), so an accepted trial needs no further elementwise pass;
//                        EPI 1: grad = G0 + H W (:300) with per-CTA row maxima (the exponents of grad's digit planes) and
//                               partial Bj = sum_i rho grad (:302).
// Every reduction has a fixed order (per-CTA partials combined in index order): results are run-to-run deterministic.
#pragma once
#include "common.cuh"

namespace lcx {
namespace fs {

constexpr int kMaxM = 128;
constexpr int kOuterCols = 32;   // variables per sub-strip of strip_outer_kernel
constexpr int kApplyCols = 32;   // variables per sub-strip of strip_apply_kernel

struct OuterArgs {
    // MODE 0: the two operands
    const double* A;
    const double* B;
    // MODE 1 (rho = c1 D + e2 W) / MODE 2 (W2 = W + eta U, rho = rho0 + eta Rdir): stage 1 of the moments
    const double* D;
    const double* W;
    const double* U;
    const double* rho0;
    const double* Rdir;
    double eta, c1, e2;
    double* W2;
    double* rho;
    double* invrho;
    double* rinv;
    double* Si;
    // output: part[cta][a][b] (m x ldm each)
    double* part;
    long long part_stride;
    int m, n, cols_per_cta;
    long long ld, ldm;
};

// acc[ii][jj] += sum over the strip's k-steps of A-tile(wr + 2 ii) x B-tile(wc + 4 jj)^T, TCL live tile columns; the fragments of
// k-step ks + 1 are requested before the DMMAs of k-step ks are issued.
template <int TR, int TC, int TCL>
__device__ __forceinline__ void outer_product(const double* As, const double* Bs, int P, int wr, int wc, int fr, int fk, int nks,
                                              double (&acc)[TR][TC][2]) {
    constexpr int NB = TCL > 0 ? TCL : 1;
    double a0[TR], b0[NB], a1[TR], b1[NB];
    const double* ap = As + (wr * 8 + fr) * P + fk;
    const double* bp = Bs + (wc * 8 + fr) * P + fk;
#pragma unroll
    for (int ii = 0; ii < TR; ++ii) a0[ii] = ap[ii * 16 * P];
#pragma unroll
    for (int jj = 0; jj < TCL; ++jj) b0[jj] = bp[jj * 32 * P];
    for (int ks = 0; ks < nks; ks += 2) {
        const bool more1 = ks + 1 < nks, more2 = ks + 2 < nks;
        if (more1) {
#pragma unroll
            for (int ii = 0; ii < TR; ++ii) a1[ii] = ap[ii * 16 * P + (ks + 1) * 4];
#pragma unroll
            for (int jj = 0; jj < TCL; ++jj) b1[jj] = bp[jj * 32 * P + (ks + 1) * 4];
        }
#pragma unroll
        for (int ii = 0; ii < TR; ++ii)
#pragma unroll
            for (int jj = 0; jj < TCL; ++jj) dmma884(acc[ii][jj][0], acc[ii][jj][1], a0[ii], b0[jj]);
        if (more1) {
            if (more2) {
#pragma unroll
                for (int ii = 0; ii < TR; ++ii) a0[ii] = ap[ii * 16 * P + (ks + 2) * 4];
#pragma unroll
                for (int jj = 0; jj < TCL; ++jj) b0[jj] = bp[jj * 32 * P + (ks + 2) * 4];
            }
#pragma unroll
            for (int ii = 0; ii < TR; ++ii)
#pragma unroll
                for (int jj = 0; jj < TCL; ++jj) dmma884(acc[ii][jj][0], acc[ii][jj][1], a1[ii], b1[jj]);
        }
    }
}

// Warp w = threadIdx.y owns the 8 x 8 output tiles (tile row wr + 2 i, tile column wc + 4 j), wr = w / 4, wc = w % 4, and feeds
// them with DMMA.8x8x4 (mma.sync.m8n8k4.f64): both fragments are X[row0 + lane / 4][k0 + lane % 4] of the strip as stored
// ([factor][variable], pitch 36 = 4 mod 16 doubles: conflict-free), so the elementwise stage writes the strip untransposed.
// (The FP64 FMA pipe of this part runs at ~1/7 of the DMMA rate -- measured, profiles/r02_fused_fma_vs_dmma.txt.)
template <int NT8, int MODE>
__global__ void __launch_bounds__(256) strip_outer_kernel(const OuterArgs a) {
    constexpr int MP = 8 * NT8;       // padded factor count
    constexpr int P = kOuterCols + 4;
    constexpr int TR = NT8 / 2, TC = (NT8 + 3) / 4;
    static_assert(NT8 % 2 == 0, "tile rows split evenly over the two warp rows");
    extern __shared__ __align__(16) double smem[];
    double* As = smem;                 // [MP][P]
    double* Bs = smem + MP * P;        // [MP][P]
    double* red = Bs + MP * P;         // [8][32]
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int wr = ty >> 2, wc = ty & 3;
    const int fr = tx >> 2, fk = tx & 3;   // fragment row / k index of this lane
    double acc[TR][TC][2];
#pragma unroll
    for (int i = 0; i < TR; ++i)
#pragma unroll
        for (int j = 0; j < TC; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    const int c_begin = blockIdx.x * a.cols_per_cta;
    const int c_end = min(a.n, c_begin + a.cols_per_cta);
    const double* __restrict__ Ag = a.A;
    const double* __restrict__ Bg = a.B;
    const double* __restrict__ Dg = a.D;
    const double* __restrict__ Wg = a.W;
    const double* __restrict__ Ug = a.U;
    const double* __restrict__ R0g = a.rho0;
    const double* __restrict__ Rdg = a.Rdir;
    for (int c0 = c_begin; c0 < c_end; c0 += kOuterCols) {
        const int i = c0 + tx;
        const bool live = i < c_end;
        double si = 0.0;
        // One CTA per SM and 8 warps: the global loads of a thread's NT8 rows are issued together (register batches of up to 8
        // rows), so a sub-strip pays the L2 latency once, not once per row.
#pragma unroll
        for (int h = 0; h < NT8; h += 8) {
            constexpr int HB = 8;
            double x0[HB], x1[HB], x2[HB], x3[HB];
#pragma unroll
            for (int q = 0; q < HB; ++q) {
                const int j = ty + 8 * (h + q);
                x0[q] = x1[q] = x2[q] = x3[q] = 0.0;
                if (h + q < NT8 && live && j < a.m) {
                    const long long o = (long long)j * a.ld + i;
                    if (MODE == 0) {
                        x0[q] = Ag[o];
                        x1[q] = Bg[o];
                    } else if (MODE == 1) {
                        x0[q] = Wg[o];
                        x1[q] = Dg[o];
                    } else {
                        x0[q] = Wg[o];
                        x1[q] = Ug[o];
                        x2[q] = R0g[o];
                        x3[q] = Rdg[o];
                    }
                }
            }
#pragma unroll
            for (int q = 0; q < HB; ++q) {
                const int j = ty + 8 * (h + q);
                if (h + q >= NT8) continue;
                double av = 0.0, bv = 0.0;
                if (live && j < a.m) {
                    const long long o = (long long)j * a.ld + i;
                    if (MODE == 0) {
                        av = x0[q];
                        bv = x1[q];
                    } else {
                        double r;
                        if (MODE == 1) {
                            av = x0[q];
                            r = a.c1 * x1[q] + a.e2 * av;
                        } else {
                            av = x0[q] + a.eta * x1[q];
                            a.W2[o] = av;
                            r = x2[q] + a.eta * x3[q];
                        }
                        const double iv = 1.0 / (1.0 - r * r);
                        const double ri = r * iv;
                        a.rho[o] = r;
                        a.invrho[o] = iv;
                        a.rinv[o] = ri;
                        si += r * ri;
                        bv = r;
                    }
                }
                As[j * P + tx] = av;
                Bs[j * P + tx] = bv;
            }
        }
        if (MODE != 0) red[ty * 32 + tx] = si;
        __syncthreads();
        if (MODE != 0 && ty == 0 && live) {
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < 8; ++k) s += red[k * 32 + tx];
            a.Si[i] = s;
        }
        const int nks = (min(kOuterCols, c_end - c0) + 3) >> 2;   // columns past the range were stored as zeros
        // tile rows wr, wr + 2, ... are always inside (NT8 is even); the last tile column of warps wc >= NT8 % 4 may not exist
        if (wc + 4 * (TC - 1) < NT8) outer_product<TR, TC, TC>(As, Bs, P, wr, wc, fr, fk, nks, acc);
        else outer_product<TR, TC, TC - 1>(As, Bs, P, wr, wc, fr, fk, nks, acc);
        __syncthreads();
    }
    double* Pout = a.part + (long long)blockIdx.x * a.part_stride;
#pragma unroll
    for (int ii = 0; ii < TR; ++ii) {
        const int ra = (wr + 2 * ii) * 8 + fr;
#pragma unroll
        for (int jj = 0; jj < TC; ++jj) {
            const int cb = (wc + 4 * jj) * 8 + fk * 2;
            if (wr + 2 * ii < NT8 && wc + 4 * jj < NT8 && ra < a.m) {
                if (cb < a.m) Pout[(long long)ra * a.ldm + cb] = acc[ii][jj][0];
                if (cb + 1 < a.m) Pout[(long long)ra * a.ldm + cb + 1] = acc[ii][jj][1];
            }
        }
    }
}

struct ApplyArgs {
    const double* Q;   // m x ldm: ry (EPI 0) or H (EPI 1)
    const double* V;   // m x ld : rinv (EPI 0) or W (EPI 1)
    int m, n, cols_per_cta;
    long long ld, ldm;
    // EPI 0
    const double* rho;
    const double* invrho;
    const double* W;
    const double* Si;
    double* Qij;
    double* QiSi2;
    double* T;
    double* G0;
    int uj_mode;
    const double* s;
    const double* w2;
    const double* ujdiag;
    double c1, e2;
    double* uj;
    double* part;        // 2 per CTA
    unsigned* ticket;
    double* out;         // [0] TC, [1] max uj
    // EPI 1 (rho is shared with EPI 0)
    double* G;           // in: G0, out: grad
    double* pmax;        // [cta][ldm]
    double* pdot;        // [cta][ldm]
};

// acc[i][j] += M-tile(ty + 8 i) x V-tile(j) over the m factors, TRL live tile rows and TCL live column tiles of this warp.
template <int TR, int TRL, int TCL>
__device__ __forceinline__ void apply_product(const double* Ms, const double* Vs, int PM, int PV, int ty, int fr, int fk, int nks,
                                              double (&acc)[TR][4][2]) {
    double a0[TRL], b0[TCL], a1[TRL], b1[TCL];
    const double* ap = Ms + (ty * 8 + fr) * PM + fk;
    const double* bp = Vs + fk * PV + fr;
#pragma unroll
    for (int i = 0; i < TRL; ++i) a0[i] = ap[i * 64 * PM];
#pragma unroll
    for (int j = 0; j < TCL; ++j) b0[j] = bp[j * 8];
    for (int ks = 0; ks < nks; ks += 2) {
        const bool more1 = ks + 1 < nks, more2 = ks + 2 < nks;
        if (more1) {
#pragma unroll
            for (int i = 0; i < TRL; ++i) a1[i] = ap[i * 64 * PM + (ks + 1) * 4];
#pragma unroll
            for (int j = 0; j < TCL; ++j) b1[j] = bp[(ks + 1) * 4 * PV + j * 8];
        }
#pragma unroll
        for (int i = 0; i < TRL; ++i)
#pragma unroll
            for (int j = 0; j < TCL; ++j) dmma884(acc[i][j][0], acc[i][j][1], a0[i], b0[j]);
        if (more1) {
            if (more2) {
#pragma unroll
                for (int i = 0; i < TRL; ++i) a0[i] = ap[i * 64 * PM + (ks + 2) * 4];
#pragma unroll
                for (int j = 0; j < TCL; ++j) b0[j] = bp[(ks + 2) * 4 * PV + j * 8];
            }
#pragma unroll
            for (int i = 0; i < TRL; ++i)
#pragma unroll
                for (int j = 0; j < TCL; ++j) dmma884(acc[i][j][0], acc[i][j][1], a1[i], b1[j]);
        }
    }
}

// The m x m matrix stays in shared memory for the life of the CTA ([j][k], pitch MP + 4); per sub-strip of 32 variables warp
// w computes the 8 x 8 tiles (tile row w and w + 8) x (4 column tiles) with DMMA.8x8x4, parks them in shared memory, and the
// epilogue runs in the elementwise mapping (lane = variable, warp = factor row group ty, ty + 8, ...).
template <int NT8, int EPI>
__global__ void __launch_bounds__(256) strip_apply_kernel(const ApplyArgs a) {
    constexpr int MP = 8 * NT8;        // padded factor count (NT8 even: MP is a multiple of 16)
    constexpr int PM = MP + 4;         // 4 mod 16: conflict-free A fragments
    constexpr int PV = kApplyCols + 4; // 36
    constexpr int PQ = kApplyCols + 1; // 33
    constexpr int TR = (NT8 + 7) / 8;  // tile rows per warp (1 or 2)
    extern __shared__ __align__(16) double smem[];
    double* Ms = smem;                 // [MP][PM]
    double* Vs = Ms + MP * PM;         // [MP][PV]
    double* Qs = Vs + MP * PV;         // [MP][PQ]
    double* red = Qs + MP * PQ;        // [8][32]
    double* ujs = red + 8 * 32;        // [MP]
    __shared__ int is_last;
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int t = ty * 32 + tx;
    const int fr = tx >> 2, fk = tx & 3;
    const int m = a.m;
    {   // M -> shared memory, 16 loads in flight per thread
        const double* __restrict__ Qg = a.Q;
        constexpr int TOTAL = MP * MP;
        for (int base = 0; base < TOTAL; base += 256 * 16) {
            double v[16];
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                const int idx = base + q * 256 + t;
                const int j = idx / MP, k = idx - j * MP;
                v[q] = (idx < TOTAL && j < m && k < m) ? Qg[(long long)j * a.ldm + k] : 0.0;
            }
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                const int idx = base + q * 256 + t;
                const int j = idx / MP, k = idx - j * MP;
                if (idx < TOTAL) Ms[j * PM + k] = v[q];
            }
        }
    }
    if (EPI == 0) {
        for (int j = t; j < MP; j += 256)
            ujs[j] = (j < m) ? (a.uj_mode == 0 ? a.c1 * a.s[j] + a.e2 * a.w2[j] : a.ujdiag[j]) : 0.0;
    }
    double l1 = 0.0, l2 = 0.0;     // EPI 0: log sums (threads of warp 0)
    double mx[NT8], dt[NT8];       // EPI 1: row maxima / partial Bj of rows ty + 8 i
#pragma unroll
    for (int i = 0; i < NT8; ++i) mx[i] = dt[i] = 0.0;
    const int c_begin = blockIdx.x * a.cols_per_cta;
    const int c_end = min(a.n, c_begin + a.cols_per_cta);
    const int nks = (m + 3) >> 2;
    const double* __restrict__ Vg = a.V;
    const double* __restrict__ Rg = a.rho;
    const double* __restrict__ IVg = a.invrho;
    const double* __restrict__ Wg = a.W;
    const double* __restrict__ G0g = a.G;
    for (int c0 = c_begin; c0 < c_end; c0 += kApplyCols) {
        const int col = c0 + tx;
        const bool live = col < c_end;
        __syncthreads();   // the previous sub-strip's Vs / Qs / red are no longer read (also orders the Ms / ujs fill)
        {
            double v[NT8];
#pragma unroll
            for (int i = 0; i < NT8; ++i) {
                const int k = ty + 8 * i;
                v[i] = (live && k < m) ? Vg[(long long)k * a.ld + col] : 0.0;
            }
#pragma unroll
            for (int i = 0; i < NT8; ++i) Vs[(ty + 8 * i) * PV + tx] = v[i];
        }
        // (the epilogue's global operands are requested before the product so their latency hides behind the DMMAs)
        double e0[NT8], e1[NT8];
#pragma unroll
        for (int i = 0; i < NT8; ++i) {
            const int j = ty + 8 * i;
            e0[i] = e1[i] = 0.0;
            if (live && j < m) {
                const long long o = (long long)j * a.ld + col;
                e0[i] = Rg[o];                        // rho
                e1[i] = (EPI == 0) ? IVg[o] : G0g[o];  // invrho / G0
            }
        }
        __syncthreads();
        const int ntc = (min(kApplyCols, c_end - c0) + 7) >> 3;   // live column tiles
        double acc[TR][4][2];
#pragma unroll
        for (int i = 0; i < TR; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
        // warp-uniform choices made once per sub-strip, so the k loop carries no predicates
        const bool two = (TR == 2) && (ty + 8 < NT8);
        if (ntc == 4) {
            if (two) apply_product<TR, TR, 4>(Ms, Vs, PM, PV, ty, fr, fk, nks, acc);
            else apply_product<TR, 1, 4>(Ms, Vs, PM, PV, ty, fr, fk, nks, acc);
        } else if (ntc == 3) {
            if (two) apply_product<TR, TR, 3>(Ms, Vs, PM, PV, ty, fr, fk, nks, acc);
            else apply_product<TR, 1, 3>(Ms, Vs, PM, PV, ty, fr, fk, nks, acc);
        } else if (ntc == 2) {
            if (two) apply_product<TR, TR, 2>(Ms, Vs, PM, PV, ty, fr, fk, nks, acc);
            else apply_product<TR, 1, 2>(Ms, Vs, PM, PV, ty, fr, fk, nks, acc);
        } else {
            if (two) apply_product<TR, TR, 1>(Ms, Vs, PM, PV, ty, fr, fk, nks, acc);
            else apply_product<TR, 1, 1>(Ms, Vs, PM, PV, ty, fr, fk, nks, acc);
        }
#pragma unroll
        for (int i = 0; i < TR; ++i) {
            const int tr = ty + 8 * i;
            if (tr < NT8) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    Qs[(tr * 8 + fr) * PQ + j * 8 + fk * 2] = acc[i][j][0];
                    Qs[(tr * 8 + fr) * PQ + j * 8 + fk * 2 + 1] = acc[i][j][1];
                }
            }
        }
        __syncthreads();
        if (EPI == 0) {
            const double si = live ? a.Si[col] : 0.0;
            double partial = 0.0;
            double wv[NT8];
#pragma unroll
            for (int i = 0; i < NT8; ++i) {
                const int j = ty + 8 * i;
                wv[i] = 0.0;
                if (live && j < m) {
                    const long long o = (long long)j * a.ld + col;
                    wv[i] = Wg[o];
                    const double qv = Qs[j * PQ + tx];
                    a.Qij[o] = qv;
                    partial += Vs[j * PV + tx] * (qv - si * e0[i]);
                }
            }
            red[ty * 32 + tx] = partial;
            __syncthreads();
            if (live) {
                double qs = 0.0;
#pragma unroll
                for (int g = 0; g < 8; ++g) qs += red[g * 32 + tx];
                if (ty == 0) {
                    a.QiSi2[col] = qs;
                    l1 += log(1.0 + si);
                    l2 += log(1.0 + qs);
                }
                const double q1 = 1.0 + qs;
#pragma unroll
                for (int i = 0; i < NT8; ++i) {
                    const int j = ty + 8 * i;
                    if (j < m) {
                        const long long o = (long long)j * a.ld + col;
                        const double r = e0[i], iv = e1[i], ri = Vs[j * PV + tx];
                        const double rj = 1.0 - ujs[j];
                        a.T[o] = ri / q1;
                        double gval = wv[i] / rj;
                        gval -= 2.0 * iv * ri / (1.0 + si);
                        gval += iv * iv * ((1.0 + r * r) * Qs[j * PQ + tx] - 2.0 * r * si) / q1;
                        a.G0[o] = gval;
                    }
                }
            }
        } else if (live) {
#pragma unroll
            for (int i = 0; i < NT8; ++i) {
                const int j = ty + 8 * i;
                if (j < m) {
                    const long long o = (long long)j * a.ld + col;
                    const double g = e1[i] + Qs[j * PQ + tx];
                    a.G[o] = g;
                    mx[i] = lcx::amax_acc(mx[i], g);
                    dt[i] += e0[i] * g;
                }
            }
        }
    }
    if (EPI == 1) {
        // a warp is one ty: its 32 lanes hold the same rows ty + 8 i over different variables
#pragma unroll
        for (int i = 0; i < NT8; ++i) {
            const double vmax = warp_max(mx[i]);
            const double vdot = warp_sum(dt[i]);
            const int j = ty + 8 * i;
            if (tx == 0 && j < m) {
                a.pmax[(long long)blockIdx.x * a.ldm + j] = vmax;
                a.pdot[(long long)blockIdx.x * a.ldm + j] = vdot;
            }
        }
        return;
    }
    // EPI 0: per-CTA log sums -> last CTA finishes TC, uj, max uj (moments_finish_kernel's arithmetic)
    if (ty == 0) {
        l1 = warp_sum(l1);
        l2 = warp_sum(l2);
        if (tx == 0) {
            a.part[2 * blockIdx.x] = l1;
            a.part[2 * blockIdx.x + 1] = l2;
            __threadfence();
            const unsigned prev = atomicAdd(a.ticket, 1u);
            is_last = (prev == gridDim.x - 1) ? 1 : 0;
        }
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    double* scratch = red;   // 8 doubles
    double s1 = 0.0, s2 = 0.0;
    for (int b = t; b < (int)gridDim.x; b += 256) {
        s1 += __ldcg(&a.part[2 * b]);
        s2 += __ldcg(&a.part[2 * b + 1]);
    }
    auto bsum = [&](double v) {   // block_sum_256's fixed tree over the (32, 8) thread block
        v = warp_sum(v);
        __syncthreads();
        if (tx == 0) scratch[ty] = v;
        __syncthreads();
        double r = 0.0;
#pragma unroll
        for (int i = 0; i < 8; ++i) r += scratch[i];
        return r;
    };
    s1 = bsum(s1);
    s2 = bsum(s2);
    double l3 = 0.0, um = -1e300;
    for (int j = t; j < m; j += 256) {
        const double u = ujs[j];
        a.uj[j] = u;
        l3 += log(1.0 - u);
        um = fmax(um, u);
        if (!(u == u)) um = 1e300;
    }
    l3 = bsum(l3);
    um = warp_max(um);
    __syncthreads();
    if (tx == 0) scratch[ty] = um;
    __syncthreads();
    if (t == 0) {
        double r = scratch[0];
#pragma unroll
        for (int i = 1; i < 8; ++i) r = fmax(r, scratch[i]);
        a.out[0] = s1 - 0.5 * s2 + 0.5 * l3;
        a.out[1] = r;
        *a.ticket = 0u;
    }
}

// Bj[j] = sum over CTAs of the partial dots (index order) -- the DMMA mode has no slicing pass to carry this
__global__ void bj_finish_kernel(const double* __restrict__ pdot, int nparts, long long ldp, double* __restrict__ bj, int m) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    double s = 0.0;
    for (int p = 0; p < nparts; ++p) s += pdot[(long long)p * ldp + j];
    bj[j] = s;
}

// ---- host side ------------------------------------------------------------------------------------
inline size_t outer_smem(int nt8) { return (size_t)(2 * 8 * nt8 * (kOuterCols + 4) + 8 * 32) * sizeof(double); }
inline size_t apply_smem(int nt8) {
    const int mp = 8 * nt8;
    return (size_t)(mp * (mp + 4) + mp * (kApplyCols + 4) + mp * (kApplyCols + 1) + 8 * 32 + mp) * sizeof(double);
}

template <int NT8, int MODE>
inline int launch_outer_inst(const OuterArgs& a, int grid, cudaStream_t st) {
    static PerDeviceOnce configured = {};
    auto kern = strip_outer_kernel<NT8, MODE>;
    if (configured.first_time())
        LCX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)outer_smem(NT8)));
    kern<<<grid, dim3(32, 8), outer_smem(NT8), st>>>(a);
    LCX_CUDA(cudaGetLastError());
    return 0;
}

template <int MODE>
inline int launch_outer(const OuterArgs& a, int grid, cudaStream_t st) {
    switch ((a.m + 15) / 16) {   // NT8 = 2 * ceil(m / 16) tiles of 8 factors
        case 1: return launch_outer_inst<2, MODE>(a, grid, st);
        case 2: return launch_outer_inst<4, MODE>(a, grid, st);
        case 3: return launch_outer_inst<6, MODE>(a, grid, st);
        case 4: return launch_outer_inst<8, MODE>(a, grid, st);
        case 5: return launch_outer_inst<10, MODE>(a, grid, st);
        case 6: return launch_outer_inst<12, MODE>(a, grid, st);
        case 7: return launch_outer_inst<14, MODE>(a, grid, st);
        case 8: return launch_outer_inst<16, MODE>(a, grid, st);
    }
    return fail(-1, "launch_outer", "m exceeds the fused path");
}

template <int NT8, int EPI>
inline int launch_apply_inst(const ApplyArgs& a, int grid, cudaStream_t st) {
    static PerDeviceOnce configured = {};
    auto kern = strip_apply_kernel<NT8, EPI>;
    if (configured.first_time())
        LCX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)apply_smem(NT8)));
    kern<<<grid, dim3(32, 8), apply_smem(NT8), st>>>(a);
    LCX_CUDA(cudaGetLastError());
    return 0;
}

template <int EPI>
inline int launch_apply(const ApplyArgs& a, int grid, cudaStream_t st) {
    switch ((a.m + 15) / 16) {
        case 1: return launch_apply_inst<2, EPI>(a, grid, st);
        case 2: return launch_apply_inst<4, EPI>(a, grid, st);
        case 3: return launch_apply_inst<6, EPI>(a, grid, st);
        case 4: return launch_apply_inst<8, EPI>(a, grid, st);
        case 5: return launch_apply_inst<10, EPI>(a, grid, st);
        case 6: return launch_apply_inst<12, EPI>(a, grid, st);
        case 7: return launch_apply_inst<14, EPI>(a, grid, st);
        case 8: return launch_apply_inst<16, EPI>(a, grid, st);
    }
    return fail(-1, "launch_apply", "m exceeds the fused path");
}

}  // namespace fs
}  // namespace lcx
