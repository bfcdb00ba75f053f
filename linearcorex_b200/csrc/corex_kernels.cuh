// Fused elementwise / reduction kernels of the Linear CorEx fit loop (FP64).
//
// All m x n arrays ("factor-major") are stored row-major with leading dimension ld >= n; only
// columns < n are ever read or written.  Column reductions (over factors j, one result per
// variable i) use a 32-column strip per CTA with 8 row groups (blockDim = 32 x 8) and a fixed-order
// shared-memory combine, so every result is run-to-run deterministic.  Scalar reductions go
// through per-CTA partials that a single-CTA "finish" kernel adds in index order.
//
// Reference formulas: linearcorex/linearcorex.py, lines cited per kernel.
#pragma once
#include "common.cuh"

namespace lcx {

constexpr int kStripCols = 32;
constexpr int kStripRows = 8;

// ---- small row/diag helpers ---------------------------------------------------------------------

// out[j] = sum_i a[j][i] * (b ? b[j][i] : 1)      one CTA (256 threads) per row.  (:249 sum(ws**2), :302 Bj)
__global__ void row_dot_kernel(const double* __restrict__ a, const double* __restrict__ b, double* __restrict__ out,
                               int n, long long ld) {
    __shared__ double scratch[8];
    const int j = blockIdx.x;
    const double* ra = a + (long long)j * ld;
    const double* rb = b ? b + (long long)j * ld : nullptr;
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) s += rb ? ra[i] * rb[i] : ra[i];
    s = block_sum_256(s, scratch);
    if (threadIdx.x == 0) out[j] = s;
}

// a[j][:] *= f[j]
__global__ void scale_rows_kernel(double* __restrict__ a, const double* __restrict__ f, int m, int n, long long ld) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y;
    if (i < n && j < m) a[(long long)j * ld + i] *= f[j];
}

// dst[j][:] = src[j][:] * f[j]
__global__ void scale_rows_out_kernel(const double* __restrict__ src, const double* __restrict__ f,
                                      double* __restrict__ dst, int m, int n, long long ld) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y;
    if (i < n && j < m) dst[(long long)j * ld + i] = src[(long long)j * ld + i] * f[j];
}

// diag_out[j] = a[j][j] (optional); a[j][j] = value.   (np.fill_diagonal, :263 / :295 / :379)
__global__ void diag_fix_kernel(double* __restrict__ a, long long ld, int m, double value, double* __restrict__ diag_out) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < m) {
        if (diag_out) diag_out[j] = a[(long long)j * ld + j];
        a[(long long)j * ld + j] = value;
    }
}

// s[j] = sum over CTAs of the K1 column-sum-of-squares partials (fixed order).
__global__ void reduce_colsq_kernel(const double* __restrict__ part, int nparts, int ldp, double* __restrict__ s, int m) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < m) {
        double acc = 0.0;
        for (int b = 0; b < nparts; ++b) acc += part[(long long)b * ldp + j];
        s[j] = acc;
    }
}

// ---- moments, stage 1 ---------------------------------------------------------------------------
// rho -> invrho, rhoinvrho, Si.                                                     (:260, :264-265, :268)
//   FROM_D : rho = c1 * D + e2 * W           (c1 = (1-eps^2)/N, e2 = eps^2; D = X~^T (X~ W^T) summed over ranks)
//   !FROM_D: W2 = W + eta U, rho = rho0 + eta R  (line-search trial evaluated through the linearity of _sig;
//            SURVEY.md section 7.8), W2 is written as well.
template <bool FROM_D>
__global__ void __launch_bounds__(256) moments_stage1_kernel(
    const double* __restrict__ Dsrc, const double* __restrict__ W, const double* __restrict__ U,
    const double* __restrict__ rho0, const double* __restrict__ Rdir, double eta, double c1, double e2,
    double* __restrict__ W2, double* __restrict__ rho, double* __restrict__ invrho, double* __restrict__ rinv,
    double* __restrict__ Si, int m, int n, long long ld) {
    __shared__ double red[kStripRows][kStripCols];
    const int i = blockIdx.x * kStripCols + threadIdx.x;
    double si = 0.0;
    if (i < n) {
        for (int j = threadIdx.y; j < m; j += kStripRows) {
            const long long o = (long long)j * ld + i;
            double r;
            if (FROM_D) {
                r = c1 * Dsrc[o] + e2 * W[o];
            } else {
                W2[o] = W[o] + eta * U[o];
                r = rho0[o] + eta * Rdir[o];
            }
            const double iv = 1.0 / (1.0 - r * r);
            const double ri = r * iv;
            rho[o] = r;
            invrho[o] = iv;
            rinv[o] = ri;
            si += r * ri;
        }
    }
    red[threadIdx.y][threadIdx.x] = si;
    __syncthreads();
    if (threadIdx.y == 0 && i < n) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < kStripRows; ++k) s += red[k][threadIdx.x];
        Si[i] = s;
    }
}

// ---- moments, stage 2 ---------------------------------------------------------------------------
// Qi-Si^2 = sum_k rinv_ki (Qij_ki - Si_i rho_ki)   (:269) and the two log-sums of the objective (:272-273).
// part[blockIdx.x] = { sum log(1+Si), sum log(1+Qi-Si^2) } over this strip.
__global__ void __launch_bounds__(256) moments_stage2_kernel(
    const double* __restrict__ rho, const double* __restrict__ rinv, const double* __restrict__ Qij,
    const double* __restrict__ Si, double* __restrict__ QiSi2, double* __restrict__ part, int m, int n, long long ld) {
    __shared__ double red[kStripRows][kStripCols];
    const int i = blockIdx.x * kStripCols + threadIdx.x;
    double acc = 0.0;
    double si = 0.0;
    if (i < n) {
        si = Si[i];
        for (int j = threadIdx.y; j < m; j += kStripRows) {
            const long long o = (long long)j * ld + i;
            acc += rinv[o] * (Qij[o] - si * rho[o]);
        }
    }
    red[threadIdx.y][threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.y == 0) {  // one warp finishes the strip
        double l1 = 0.0, l2 = 0.0;
        if (i < n) {
            double q = 0.0;
#pragma unroll
            for (int k = 0; k < kStripRows; ++k) q += red[k][threadIdx.x];
            QiSi2[i] = q;
            l1 = log(1.0 + si);
            l2 = log(1.0 + q);
        }
        l1 = warp_sum(l1);
        l2 = warp_sum(l2);
        if (threadIdx.x == 0) {
            part[2 * blockIdx.x] = l1;
            part[2 * blockIdx.x + 1] = l2;
        }
    }
}

// Scalars of one moment evaluation.  Single CTA, 256 threads.
//   uj_mode 0: uj = c1 * s + e2 * w2        (:249; s = sum_l Y_lj^2 over all ranks, w2 = sum_i W_ji^2)
//   uj_mode 1: uj = ujdiag                  (diag(W rho^T), the same quantity through linearity)
// out[0] = TC (:272-274), out[1] = max_j uj (:250).
__global__ void moments_finish_kernel(const double* __restrict__ part, int nparts, int uj_mode,
                                      const double* __restrict__ s, const double* __restrict__ w2,
                                      const double* __restrict__ ujdiag, double c1, double e2,
                                      double* __restrict__ uj, int m, double* __restrict__ out) {
    __shared__ double scratch[8];
    double l1 = 0.0, l2 = 0.0;
    for (int b = threadIdx.x; b < nparts; b += 256) {
        l1 += part[2 * b];
        l2 += part[2 * b + 1];
    }
    l1 = block_sum_256(l1, scratch);
    l2 = block_sum_256(l2, scratch);
    double l3 = 0.0, mx = -1e300;
    for (int j = threadIdx.x; j < m; j += 256) {
        const double u = (uj_mode == 0) ? c1 * s[j] + e2 * w2[j] : ujdiag[j];
        uj[j] = u;
        l3 += log(1.0 - u);
        mx = fmax(mx, u);
        if (!(u == u)) mx = 1e300;  // NaN uj counts as invalid, like `np.max(uj) >= 1` on a NaN-free path never would
    }
    l3 = block_sum_256(l3, scratch);
    mx = block_max_256(mx, scratch);
    if (threadIdx.x == 0) {
        out[0] = l1 - 0.5 * l2 + 0.5 * l3;
        out[1] = mx;
    }
}

// uj only (used by _norm at init, :228): uj = c1*s + e2*w2, f = 1 / (10 sqrt(uj))          (:117)
__global__ void init_scale_kernel(const double* __restrict__ s, const double* __restrict__ w2, double c1, double e2,
                                  double* __restrict__ f, int m) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < m) f[j] = 1.0 / (10.0 * sqrt(c1 * s[j] + e2 * w2[j]));
}

// Anneal-stage rescale factor (:130-133): f = 0.001 floor(1000 a), a = sqrt((1-e0^2)/((1-e^2)(1+delta)))
__global__ void stage_scale_kernel(const double* __restrict__ wmag, const double* __restrict__ uj, double eps,
                                   double eps0, double* __restrict__ f, int m) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < m) {
        const double delta = (eps * eps - eps0 * eps0) / (1.0 - eps * eps) * wmag[j] / uj[j];
        const double a = sqrt((1.0 - eps0 * eps0) / ((1.0 - eps * eps) * (1.0 + delta)));
        f[j] = 0.001 * floor(1000.0 * a);
    }
}

// ---- search direction (:292-305) ----------------------------------------------------------------
// T = rinv / (1 + Qi-Si^2)  (left factor of H, :294) and the H-free part of the gradient (:296-299):
// G0 = W/rj - 2 invrho rinv/(1+Si) + invrho^2 ((1+rho^2) Qij - 2 rho Si)/(1+Qi-Si^2)
__global__ void direction_stage1_kernel(const double* __restrict__ W, const double* __restrict__ rho,
                                        const double* __restrict__ invrho, const double* __restrict__ rinv,
                                        const double* __restrict__ Qij, const double* __restrict__ Si,
                                        const double* __restrict__ QiSi2, const double* __restrict__ uj,
                                        double* __restrict__ T, double* __restrict__ G0, int m, int n, long long ld) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y;
    if (i >= n || j >= m) return;
    const long long o = (long long)j * ld + i;
    const double rj = 1.0 - uj[j];
    const double si = Si[i], q1 = 1.0 + QiSi2[i];
    const double r = rho[o], iv = invrho[o], ri = rinv[o];
    T[o] = ri / q1;
    double gval = W[o] / rj;
    gval -= 2.0 * iv * ri / (1.0 + si);
    gval += iv * iv * ((1.0 + r * r) * Qij[o] - 2.0 * r * si) / q1;
    G0[o] = gval;
}

// update = -rj (G - 2 W/(2-rj) Bj) (:303);  sigG = c1 D_G + e2 G (:212);  tangent = sum sigG*update (:305).
// Rdir = -rj (sigG - 2 rho/(2-rj) Bj) = _sig(update): rho(W + eta U) = rho + eta Rdir by linearity.
// part[blockIdx.y * gridDim.x + blockIdx.x] = partial tangent.  DG may be given as dg_splits split-K partials (dg_stride apart),
// summed here in index order -- the Gram route's product hands its partials over without a combine pass.
__global__ void __launch_bounds__(256) direction_stage2_kernel(
    const double* __restrict__ W, const double* __restrict__ rho, const double* __restrict__ G,
    const double* __restrict__ DG, const double* __restrict__ uj, const double* __restrict__ Bj, double c1, double e2,
    double* __restrict__ U, double* __restrict__ Rdir, double* __restrict__ part, int m, int n, long long ld,
    int dg_splits = 1, long long dg_stride = 0, unsigned* __restrict__ ticket = nullptr, double* __restrict__ out = nullptr) {
    __shared__ double scratch[8];
    __shared__ int is_last;
    const int i = blockIdx.x * 256 + threadIdx.x;
    const int j = blockIdx.y;
    double tang = 0.0;
    if (i < n) {
        const long long o = (long long)j * ld + i;
        const double rj = 1.0 - uj[j];
        const double bj = Bj[j];
        const double gval = G[o];
        double dg = DG[o];
        for (int z = 1; z < dg_splits; ++z) dg += DG[(long long)z * dg_stride + o];  // split-K partials, fixed order
        const double sg = c1 * dg + e2 * gval;
        const double u = -rj * (gval - 2.0 * W[o] / (2.0 - rj) * bj);
        U[o] = u;
        Rdir[o] = -rj * (sg - 2.0 * rho[o] / (2.0 - rj) * bj);
        tang = sg * u;
    }
    tang = block_sum_256(tang, scratch);
    if (threadIdx.x == 0) part[(long long)blockIdx.y * gridDim.x + blockIdx.x] = tang;
    if (ticket == nullptr) return;
    // the last CTA to arrive adds all partials in index order (sum_partials_kernel's arithmetic, one launch less)
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned total = gridDim.x * gridDim.y;
        is_last = (atomicAdd(ticket, 1u) == total - 1) ? 1 : 0;
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    const int nparts = (int)(gridDim.x * gridDim.y);
    double s = 0.0;
    for (int b = threadIdx.x; b < nparts; b += 256) s += __ldcg(&part[b]);
    s = block_sum_256(s, scratch);
    if (threadIdx.x == 0) {
        out[0] = s;
        *ticket = 0u;
    }
}

// The 16 scalars of an iteration (TC, max uj, tangent, ..., the solve status) stored straight into the session's page-locked
// mailbox -- host memory, device-addressable under UVA -- followed by a sequence number the host polls: no copy engine, no
// stream synchronisation between the last kernel of a trial and the host's accept / backtrack decision.  One warp.
__global__ void post_mailbox_kernel(const double* __restrict__ scalars, volatile double* box, unsigned long long seq) {
    if (threadIdx.x < 16) box[threadIdx.x] = scalars[threadIdx.x];
    __threadfence_system();
    __syncwarp();
    if (threadIdx.x == 0) *reinterpret_cast<volatile unsigned long long*>(box + 16) = seq;
}

// out[0] = sum of partials (fixed order).  Single CTA.
__global__ void sum_partials_kernel(const double* __restrict__ part, int nparts, double* __restrict__ out) {
    __shared__ double scratch[8];
    double s = 0.0;
    for (int b = threadIdx.x; b < nparts; b += 256) s += part[b];
    s = block_sum_256(s, scratch);
    if (threadIdx.x == 0) out[0] = s;
}

// ---- details (quick=False extras, :277-287) -------------------------------------------------------
// Per column: MI = -0.5 log1p(-rho^2) (written), sum_j MI, max_j MI, X_i^2|Y = clip(1 - sum_j XZ_ji XYorRho_ji, 1e-6),
// I(X_i;Y) = -0.5 log(X_i^2|Y).  part[b] = { sum_i max_j MI, sum_i I(X_i;Y), sum_i (sum_j MI - I(X_i;Y)) }.
__global__ void __launch_bounds__(256) details_cols_kernel(
    const double* __restrict__ rho, const double* __restrict__ XZ, const double* __restrict__ other,
    double* __restrict__ MI, double* __restrict__ X2Y, double* __restrict__ IXY, double* __restrict__ part, int m, int n,
    long long ld) {
    __shared__ double red[3][kStripRows][kStripCols];
    const int i = blockIdx.x * kStripCols + threadIdx.x;
    double smi = 0.0, mmi = -1e300, sxz = 0.0;
    if (i < n) {
        for (int j = threadIdx.y; j < m; j += kStripRows) {
            const long long o = (long long)j * ld + i;
            const double r = rho[o];
            const double mi = -0.5 * log1p(-r * r);
            MI[o] = mi;
            smi += mi;
            mmi = fmax(mmi, mi);
            sxz += XZ[o] * other[o];
        }
    }
    red[0][threadIdx.y][threadIdx.x] = smi;
    red[1][threadIdx.y][threadIdx.x] = mmi;
    red[2][threadIdx.y][threadIdx.x] = sxz;
    __syncthreads();
    if (threadIdx.y == 0) {
        double p0 = 0.0, p1 = 0.0, p2 = 0.0;
        if (i < n) {
            double a = 0.0, b = -1e300, c = 0.0;
#pragma unroll
            for (int k = 0; k < kStripRows; ++k) {
                a += red[0][k][threadIdx.x];
                b = fmax(b, red[1][k][threadIdx.x]);
                c += red[2][k][threadIdx.x];
            }
            const double x2y = fmax(1.0 - c, 1e-6);
            const double ixy = -0.5 * log(x2y);
            X2Y[i] = x2y;
            IXY[i] = ixy;
            p0 = b;
            p1 = ixy;
            p2 = a - ixy;
        }
        p0 = warp_sum(p0);
        p1 = warp_sum(p1);
        p2 = warp_sum(p2);
        if (threadIdx.x == 0) {
            part[3 * blockIdx.x] = p0;
            part[3 * blockIdx.x + 1] = p1;
            part[3 * blockIdx.x + 2] = p2;
        }
    }
}

// Per-factor scalars of the details path.  Single CTA.
//   Yj2 = yscale^2/(1-uj) (:262) [ns]  or given [syn];  IYX = 0.5 log Yj2 (:282, yscale = 1);  TCs = rowMI - IYX (:284)
//   out[0] = TC_no_overlap (:285), out[1] = sum_i I(X_i;Y), out[2] = additivity (:287), out[3] = sum_j IYX
//   TC_direct_j = out[1] - IYX_j (:286);  sqrtY = sqrt(Yj2) (row scale of X_i Y_j, :279)
__global__ void details_finish_kernel(const double* __restrict__ part, int nparts, const double* __restrict__ uj,
                                      const double* __restrict__ yj2_in, const double* __restrict__ rowMI,
                                      double* __restrict__ Yj2, double* __restrict__ IYX, double* __restrict__ TCs,
                                      double* __restrict__ TCdirect, double* __restrict__ sqrtY, int m,
                                      double* __restrict__ out) {
    __shared__ double scratch[8];
    double p0 = 0.0, p1 = 0.0, p2 = 0.0;
    for (int b = threadIdx.x; b < nparts; b += 256) {
        p0 += part[3 * b];
        p1 += part[3 * b + 1];
        p2 += part[3 * b + 2];
    }
    p0 = block_sum_256(p0, scratch);
    p1 = block_sum_256(p1, scratch);
    p2 = block_sum_256(p2, scratch);
    double siyx = 0.0;
    for (int j = threadIdx.x; j < m; j += 256) {
        const double y2 = yj2_in ? yj2_in[j] : 1.0 / (1.0 - uj[j]);
        const double iyx = 0.5 * log(y2);
        Yj2[j] = y2;
        IYX[j] = iyx;
        sqrtY[j] = sqrt(y2);
        TCs[j] = rowMI[j] - iyx;
        TCdirect[j] = p1 - iyx;
        siyx += iyx;
    }
    siyx = block_sum_256(siyx, scratch);
    if (threadIdx.x == 0) {
        out[0] = p0 - siyx;
        out[1] = p1;
        out[2] = p2;
        out[3] = siyx;
    }
}

// ---- synergistic variant (:336-384) ---------------------------------------------------------------
// From cy (m x m): Yj2 = diag(cy) (:356), ry = cy / sqrt(Yj2 Yj2^T) (:357), isq = 1/sqrt(Yj2).
__global__ void syn_ry_kernel(const double* __restrict__ cy, long long ldm, int m, double* __restrict__ ry,
                              double* __restrict__ Yj2, double* __restrict__ isq) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    if (c < m && r < m) {
        const double dr = cy[(long long)r * ldm + r], dc = cy[(long long)c * ldm + c];
        ry[(long long)r * ldm + c] = cy[(long long)r * ldm + c] / (sqrt(dc) * sqrt(dr));
        if (c == r) {
            Yj2[r] = dr;
            isq[r] = sqrt(dr);
        }
    }
}

// XY = D / N (:354); adds yscale^2 = 1 to the diagonal of cy happens in diag code of the caller.
// rho = XY / sqrt(Yj2) (:358), invrho, rinv, Si; also Qi when Qij is available is done by syn_stage2.
__global__ void __launch_bounds__(256) syn_stage1_kernel(const double* __restrict__ XY, const double* __restrict__ sq,
                                                         double* __restrict__ rho, double* __restrict__ invrho,
                                                         double* __restrict__ rinv, double* __restrict__ Si, int m, int n,
                                                         long long ld) {
    __shared__ double red[kStripRows][kStripCols];
    const int i = blockIdx.x * kStripCols + threadIdx.x;
    double si = 0.0;
    if (i < n) {
        for (int j = threadIdx.y; j < m; j += kStripRows) {
            const long long o = (long long)j * ld + i;
            const double r = XY[o] / sq[j];
            const double iv = 1.0 / (1.0 - r * r);
            rho[o] = r;
            invrho[o] = iv;
            rinv[o] = r * iv;
            si += r * r * iv;
        }
    }
    red[threadIdx.y][threadIdx.x] = si;
    __syncthreads();
    if (threadIdx.y == 0 && i < n) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < kStripRows; ++k) s += red[k][threadIdx.x];
        Si[i] = s;
    }
}

// Qi = sum_k rinv_ki Qij_ki (:362)
__global__ void __launch_bounds__(256) syn_qi_kernel(const double* __restrict__ rinv, const double* __restrict__ Qij,
                                                     double* __restrict__ Qi, int m, int n, long long ld) {
    __shared__ double red[kStripRows][kStripCols];
    const int i = blockIdx.x * kStripCols + threadIdx.x;
    double acc = 0.0;
    if (i < n)
        for (int j = threadIdx.y; j < m; j += kStripRows) acc += rinv[(long long)j * ld + i] * Qij[(long long)j * ld + i];
    red[threadIdx.y][threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.y == 0 && i < n) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < kStripRows; ++k) s += red[k][threadIdx.x];
        Qi[i] = s;
    }
}

// a[j][j] += v
__global__ void diag_add_kernel(double* __restrict__ a, long long ld, int m, double v) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < m) a[(long long)j * ld + j] += v;
}

// _update_syn pieces (:378-382): Rm = XZ / X2Y (column scale; this is both R and the left factor of H)
__global__ void syn_colscale_kernel(const double* __restrict__ XZ, const double* __restrict__ X2Y, double* __restrict__ Rm,
                                    int m, int n, long long ld) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y;
    if (i < n && j < m) Rm[(long long)j * ld + i] = XZ[(long long)j * ld + i] / X2Y[i];
}

// W2 = (1-eta) W + eta (R - S)   (:382)
__global__ void syn_mix_kernel(const double* W, const double* __restrict__ Rm, const double* __restrict__ S, double eta,
                               double* W2, int m, int n, long long ld) {  // W2 may alias W
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y;
    if (i < n && j < m) {
        const long long o = (long long)j * ld + i;
        W2[o] = (1.0 - eta) * W[o] + eta * (Rm[o] - S[o]);
    }
}

// syn objective scalars: out[0] = TC = sum_i I(X_i;Y) - sum_j I(Y_j;X) (:372) given details_finish's out[1], out[3].
__global__ void syn_tc_kernel(const double* __restrict__ dout, double* __restrict__ out) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        out[0] = dout[1] - dout[3];
        out[1] = 0.0;
    }
}

// ---- covariance reconstruction (:443-455) ---------------------------------------------------------
// z = rinv / (1 + Si)   (:447)
__global__ void cov_z_kernel(const double* __restrict__ rinv, const double* __restrict__ Si, double* __restrict__ z, int m,
                             int n, long long ld) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y;
    if (i < n && j < m) z[(long long)j * ld + i] = rinv[(long long)j * ld + i] / (1.0 + Si[i]);
}

// cov = cov * scale; diag = 1; cov_ik *= sd_i sd_k   (:449-451); rows [row0, row0+rows) of the n x n result
__global__ void cov_finish_kernel(double* __restrict__ cov, long long ldc, int row0, int rows, int n, double inv_scale,
                                  const double* __restrict__ sd) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    if (k < n && r < rows) {
        const int i = row0 + r;
        double v = cov[(long long)r * ldc + k];
        v /= inv_scale;
        if (i == k) v = 1.0;
        cov[(long long)r * ldc + k] = sd[i] * sd[k] * v;
    }
}

}  // namespace lcx
