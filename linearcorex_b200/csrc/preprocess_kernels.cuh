// Column statistics, mean imputation and standardisation (linearcorex.py:397-429, :483-510).
//
// Raw X arrives as float32 or float64, row-major N x n (ld = ldx).  Three passes, each a column
// reduction over a row slab per CTA (blockDim = 32 x 8, grid = strips x slabs) followed by a
// fixed-order combine of the slab partials.  When rows are sharded over ranks the caller
// all-reduces the combined vectors between passes.
#pragma once
#include "common.cuh"

namespace lcx {

template <typename T>
__device__ __forceinline__ bool is_missing(T v, int has_marker, double marker, int marker_is_nan) {
    if (!has_marker) return false;
    const double d = (double)v;
    return (d != d) || (!marker_is_nan && d == marker);  // NaN is always missing once a marker is set (:505)
}

// pass 1: part_sum[slab][i] = sum of observed x, part_cnt[slab][i] = number observed (finite, not missing)
template <typename T>
__global__ void __launch_bounds__(256) colstats_sum_kernel(const T* __restrict__ x, long long N, int n, long long ldx,
                                                           int rows_per_slab, int has_marker, double marker,
                                                           int marker_is_nan, double* __restrict__ part_sum,
                                                           double* __restrict__ part_cnt, long long ldp) {
    __shared__ double rs[8][32], rc[8][32];
    const int i = blockIdx.x * 32 + threadIdx.x;
    const long long r0 = (long long)blockIdx.y * rows_per_slab;
    const long long r1 = min(N, r0 + rows_per_slab);
    double s = 0.0, c = 0.0;
    if (i < n) {
        for (long long r = r0 + threadIdx.y; r < r1; r += 8) {
            const T v = x[r * ldx + i];
            const bool obs = has_marker ? (!is_missing(v, has_marker, marker, marker_is_nan) && isfinite((double)v)) : true;
            if (obs) { s += (double)v; c += 1.0; }
        }
    }
    rs[threadIdx.y][threadIdx.x] = s;
    rc[threadIdx.y][threadIdx.x] = c;
    __syncthreads();
    if (threadIdx.y == 0 && i < n) {
        double a = 0.0, b = 0.0;
#pragma unroll
        for (int k = 0; k < 8; ++k) { a += rs[k][threadIdx.x]; b += rc[k][threadIdx.x]; }
        part_sum[(long long)blockIdx.y * ldp + i] = a;
        part_cnt[(long long)blockIdx.y * ldp + i] = b;
    }
}

// pass 2: part[slab][i] = sum over observed rows of (x - mean_i)^2   (imputed entries contribute 0)
template <typename T>
__global__ void __launch_bounds__(256) colstats_sqdev_kernel(const T* __restrict__ x, long long N, int n, long long ldx,
                                                             int rows_per_slab, int has_marker, double marker,
                                                             int marker_is_nan, const double* __restrict__ mean,
                                                             double* __restrict__ part, double* __restrict__ part_max,
                                                             long long ldp) {
    // part_max (optional): per-slab max |x - mean| -- bounds |X~| before X~ exists (streamed digit slicing)
    __shared__ double rs[8][32], rm[8][32];
    const int i = blockIdx.x * 32 + threadIdx.x;
    const long long r0 = (long long)blockIdx.y * rows_per_slab;
    const long long r1 = min(N, r0 + rows_per_slab);
    double s = 0.0, mx = 0.0;
    if (i < n) {
        const double mu = mean[i];
        for (long long r = r0 + threadIdx.y; r < r1; r += 8) {
            const T v = x[r * ldx + i];
            if (!is_missing(v, has_marker, marker, marker_is_nan)) {
                const double d = (double)v - mu;
                s += d * d;
                mx = amax_acc(mx, d);
            }
        }
    }
    rs[threadIdx.y][threadIdx.x] = s;
    rm[threadIdx.y][threadIdx.x] = mx;
    __syncthreads();
    if (threadIdx.y == 0 && i < n) {
        double a = 0.0, b = 0.0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            a += rs[k][threadIdx.x];
            b = fmax(b, rm[k][threadIdx.x]);
        }
        part[(long long)blockIdx.y * ldp + i] = a;
        if (part_max) part_max[(long long)blockIdx.y * ldp + i] = b;
    }
}

// out[i] = max_slab part[slab][i]
__global__ void combine_slabs_max_kernel(const double* __restrict__ part, int slabs, long long ldp, double* __restrict__ out,
                                         int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        double a = 0.0;
        for (int s = 0; s < slabs; ++s) a = fmax(a, part[(long long)s * ldp + i]);
        out[i] = a;
    }
}

// out[i] = sum_slab part[slab][i]
__global__ void combine_slabs_kernel(const double* __restrict__ part, int slabs, long long ldp, double* __restrict__ out,
                                     int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        double a = 0.0;
        for (int s = 0; s < slabs; ++s) a += part[(long long)s * ldp + i];
        out[i] = a;
    }
}

// mean = sum / cnt
__global__ void finish_mean_kernel(const double* __restrict__ sum, const double* __restrict__ cnt, double* __restrict__ mean,
                                   int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) mean[i] = sum[i] / cnt[i];
}

// std = max(sqrt(sq / denom), 1e-10); denom = n_obs ('standard', :413) or N_total ('outliers', np.std :421)
__global__ void finish_std_kernel(const double* __restrict__ sq, const double* __restrict__ cnt, double n_total,
                                  int use_nobs, double* __restrict__ sd, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) sd[i] = fmax(sqrt(sq[i] / (use_nobs ? cnt[i] : n_total)), 1e-10);
}

__device__ __forceinline__ double squash_tails(double z) {  // g(), :483-487
    const double core = fmin(fmax(z, -4.0), 4.0);
    return core + tanh(z - core);
}

// X~ = (x - mean)/std  [+ g()];  missing entries take the column mean first (:404, :415, :423).
// mode: 0 = 'standard', 1 = 'outliers', 2 = 'none' (cast only).  Output fp64, ld = ldo, columns >= n zeroed.
template <typename T>
__global__ void standardize_kernel(const T* __restrict__ x, long long N, int n, long long ldx, int has_marker,
                                   double marker, int marker_is_nan, int mode, const double* __restrict__ impute,
                                   const double* __restrict__ mean, const double* __restrict__ sd,
                                   double* __restrict__ out, long long ldo) {
    const int i = blockIdx.y * blockDim.x + threadIdx.x;  // rows on grid.x (up to 2^31-1), column blocks on grid.y
    const long long r = blockIdx.x;
    if (i >= ldo || r >= N) return;
    double o = 0.0;
    if (i < n) {
        const T v = x[r * ldx + i];
        double d = (double)v;
        if (is_missing(v, has_marker, marker, marker_is_nan)) d = impute[i];
        if (mode == 2) {
            o = d;
        } else {
            o = (d - mean[i]) / sd[i];
            if (mode == 1) o = squash_tails(o);
        }
    }
    out[r * ldo + i] = o;
}

}  // namespace lcx
