// Split-K reduction of X~^T Y fused with the cross-GPU sum over sample shards, in ONE kernel over NVLink peer memory.
//
// Reference context: in the single-process reference the sum over samples is inside one GEMM (linearcorex.py:211,
// :259).  Sharded over ranks it becomes  D = sum_r sum_z partial[r][z]  -- a split-K combine followed by an all-reduce.
// Instead of `reduce_splits_kernel` + an NCCL call, this kernel does both: every rank
//   A. adds its own split-K partials (fixed order) into its slot of a symmetric buffer,
//   -- cross-GPU barrier (release/acquire flags written into the peers' memory) --
//   B. owns 1/P of the elements: reads that slice from every rank's slot over NVLink (P2P loads), adds in rank order,
//      and stores the result into EVERY rank's output buffer (P2P stores)   [reduce-scatter + all-gather, two-shot],
//   -- cross-GPU barrier --
// so each rank ends with bit-identical D (same additions in the same order everywhere), which keeps the replicated
// line-search / convergence decisions in lock-step without further communication.
//
// The symmetric buffer (allocated by the host through torch.distributed._symmetric_memory) is laid out as
//   [ slot0 : count ][ slot1 : count ][ out : count ][ flags : 2 * kMaxRanks uint64 ][ grid barrier : 2 uint64 ]
// Slots alternate per call so a fast rank's next phase A never overwrites what a slow rank is still reading.
#pragma once
#include <cooperative_groups.h>

#include "common.cuh"

namespace lcx {
namespace far {

namespace cg = cooperative_groups;
constexpr int kMaxRanks = 8;
constexpr unsigned long long kBarrierTimeoutNs = 60ULL * 1000 * 1000 * 1000;  // 60 s

struct Peers {
    double* base[kMaxRanks];  // symmetric buffer base of every rank (peer-mapped)
    int world, rank;
    long long count;          // doubles per slot
};

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ unsigned long long* flags_of(double* base, long long count) {
    return reinterpret_cast<unsigned long long*>(base + 3 * count);
}

// All ranks arrive with `epoch`; returns when every rank has.  Called by every CTA after a grid-wide sync.
__device__ __forceinline__ void cross_gpu_barrier(const Peers& p, unsigned long long epoch, cg::grid_group& grid) {
    __threadfence_system();
    grid.sync();
    if (blockIdx.x == 0 && threadIdx.x < p.world)
        st_release_sys(flags_of(p.base[threadIdx.x], p.count) + p.rank, epoch);
    if (threadIdx.x < p.world) {
        const unsigned long long* mine = flags_of(p.base[p.rank], p.count) + threadIdx.x;
        // A peer that died never arrives: after kBarrierTimeoutNs of polling the kernel traps, which surfaces as a CUDA error at
        // the caller's next synchronisation instead of a hang (a live rank is at most one iteration -- milliseconds -- behind).
        unsigned long long t0 = 0;
        unsigned polls = 0;
        while (ld_acquire_sys(mine) < epoch) {
            if ((++polls & 0x3ffu) == 0) {
                unsigned long long now;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
                if (t0 == 0) t0 = now;
                else if (now - t0 > kBarrierTimeoutNs) asm volatile("trap;");
            }
        }
    }
    __syncthreads();
}

// part: this rank's split-K partials [splits][rows][ld] (valid cols < cols); tail: `ntail` extra doubles (the column sums
// of squares) appended after the rows*ld block.  epoch0 = 2 * call_index (flags are monotonically increasing).
__global__ void __launch_bounds__(512) reduce_allreduce_kernel(Peers p, const double* __restrict__ part, int splits,
                                                               long long stride, int rows, int cols, long long ld,
                                                               const double* __restrict__ tail, int ntail, int slot,
                                                               unsigned long long epoch0) {
    cg::grid_group grid = cg::this_grid();
    const long long body = (long long)rows * ld;
    const long long count = body + ntail;
    double* my_slot = p.base[p.rank] + (long long)slot * p.count;
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long nthr = (long long)gridDim.x * blockDim.x;
    // ---- A: local split-K combine (fixed order z = 0, 1, ...), 16-byte accesses, four loads in flight; padding
    //         columns are written as zeros.  body and ld are even, so pairs never straddle a row. ----
    const unsigned ld2 = (unsigned)(ld >> 1);
    const long long body2 = body >> 1;
    for (long long i2 = tid; i2 < body2; i2 += nthr) {
        const unsigned c = ((unsigned)i2 % ld2) * 2u;  // body < 2^32 doubles is enforced by the caller
        double2 acc = make_double2(0.0, 0.0);
        if ((int)c < cols) {
            const double2* src = reinterpret_cast<const double2*>(part) + i2;
            const long long s2 = stride >> 1;
            acc = src[0];
            int z = 1;
            for (; z + 3 < splits; z += 4) {
                const double2 v0 = src[(long long)z * s2], v1 = src[(long long)(z + 1) * s2];
                const double2 v2 = src[(long long)(z + 2) * s2], v3 = src[(long long)(z + 3) * s2];
                acc.x += v0.x; acc.y += v0.y;
                acc.x += v1.x; acc.y += v1.y;
                acc.x += v2.x; acc.y += v2.y;
                acc.x += v3.x; acc.y += v3.y;
            }
            for (; z < splits; ++z) {
                const double2 v = src[(long long)z * s2];
                acc.x += v.x; acc.y += v.y;
            }
            if ((int)c + 1 >= cols) acc.y = 0.0;
        }
        reinterpret_cast<double2*>(my_slot)[i2] = acc;
    }
    for (long long i = body + tid; i < count; i += nthr) my_slot[i] = tail[i - body];
    cross_gpu_barrier(p, epoch0 + 1, grid);
    // ---- B: this rank owns 1/P of the elements: P2P loads of that slice from every rank's slot (all issued before the
    //         first add), sum in rank order, P2P stores of the result into every rank's output ----
    const long long count2 = (count + 1) >> 1;             // slots are padded to an even length by the host
    const long long per = (count2 + p.world - 1) / p.world;
    const long long lo = per * p.rank, hi = min(count2, lo + per);
    const long long slot_off2 = ((long long)slot * p.count) >> 1, out_off2 = p.count;  // (2 * count) / 2
    for (long long i2 = lo + tid; i2 < hi; i2 += nthr) {
        double2 v[kMaxRanks];
#pragma unroll
        for (int r = 0; r < kMaxRanks; ++r)
            if (r < p.world) v[r] = (reinterpret_cast<const double2*>(p.base[r]) + slot_off2)[i2];
        double2 acc = v[0];
#pragma unroll
        for (int r = 1; r < kMaxRanks; ++r)
            if (r < p.world) { acc.x += v[r].x; acc.y += v[r].y; }
#pragma unroll
        for (int r = 0; r < kMaxRanks; ++r)
            if (r < p.world) (reinterpret_cast<double2*>(p.base[r]) + out_off2)[i2] = acc;
    }
    cross_gpu_barrier(p, epoch0 + 2, grid);
}

// Gram route across ranks: every rank has computed the columns [col0, col0 + ncols) of D = (G A^T)^T (its share of the rows of
// G) into its own output region; this kernel hands that slab to every peer (P2P stores over NVLink) -- an all-gather in
// place.  First barrier: every rank has finished reading the previous D and holds its new slab; second: all slabs landed.
// Each element is produced by exactly one rank, so all ranks end with bit-identical D.
__global__ void __launch_bounds__(512) gather_cols_kernel(Peers p, int rows, long long ld, int col0, int ncols,
                                                          unsigned long long epoch0) {
    cg::grid_group grid = cg::this_grid();
    cross_gpu_barrier(p, epoch0 + 1, grid);
    const long long out_off = 2 * p.count;   // the output region of the symmetric buffer (where D lives when sharded)
    const double* mine = p.base[p.rank] + out_off;
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long nthr = (long long)gridDim.x * blockDim.x;
    const int pairs = ncols >> 1;            // col0 and ncols are even (tiles of 128 variables; the last slab ends at ld)
    for (long long i = tid; i < (long long)rows * pairs; i += nthr) {
        const int j = (int)(i / pairs), c = (int)(i - (long long)j * pairs) * 2;
        const long long o = (long long)j * ld + col0 + c;
        const double2 v = *reinterpret_cast<const double2*>(mine + o);
#pragma unroll
        for (int r = 0; r < kMaxRanks; ++r)
            if (r < p.world && r != p.rank) *reinterpret_cast<double2*>(p.base[r] + out_off + o) = v;
    }
    cross_gpu_barrier(p, epoch0 + 2, grid);
}

}  // namespace far
}  // namespace lcx
