// C ABI of liblcx_b200.so -- see include/lcx_b200.h for the contract and the reference lines each
// entry point replaces.  Host orchestration only; the arithmetic lives in dgemm_mma.cuh,
// corex_kernels.cuh and preprocess_kernels.cuh.
#include "../../include/lcx_b200.h"

#include <float.h>
#include <math.h>

#include "common.cuh"
#include "corex_kernels.cuh"
#include "dgemm_mma.cuh"
#include "fused_allreduce.cuh"
#include "ozaki_i8.cuh"
#include "preprocess_kernels.cuh"

namespace lcx {
thread_local char g_err[512] = "";
constexpr int kSMs = 148;  // B200; plans (and therefore workspace sizes) are fixed for this part
constexpr int kMaxSplitsX = 32;
constexpr int kMaxSplitsSmall = 148;

__global__ void axpy_kernel(const double* __restrict__ W, const double* __restrict__ U, double eta, double* __restrict__ W2,
                            int m, int n, long long ld) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y;
    if (i < n && j < m) W2[(long long)j * ld + i] = W[(long long)j * ld + i] + eta * U[(long long)j * ld + i];
}

// out = c1 * D + e2 * u     (_sig, linearcorex.py:212)
__global__ void sig_finish_kernel(const double* __restrict__ D, const double* __restrict__ u, double c1, double e2,
                                  double* __restrict__ out, int m, int n, long long ld) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y;
    if (i < n && j < m) out[(long long)j * ld + i] = c1 * D[(long long)j * ld + i] + e2 * u[(long long)j * ld + i];
}
}  // namespace lcx

using namespace lcx;

// internal (non-exported) workspace slots appended after the public enum
enum {
    I_T = LCX_A_COUNT,  // m x ld   rinv/(1+Qi-Si^2), also z of get_covariance and R of _update_syn
    I_PART,             // split-K partials
    I_COLSQ,            // K1 per-CTA column-sum-of-squares partials
    I_SPART,            // scalar partials
    I_W2,               // m   sum_i W^2
    I_BJ,               // m
    I_F,                // m   row scale factors
    I_UJDIAG,           // m   diag(W rho^T)
    I_ROWMI,            // m
    I_SQRTY,            // m
    I_RYINV,            // m x ldm
    I_AUG,              // m x 2m
    I_STATUS,           // 2 doubles (int status of the inverse)
    I_XS,               // split modes: int8 digit slices of X~   [S][N_local][ld8]
    I_AS,               //              int8 digit slices of A    [S][m][ld8]
    I_YS,               //              int8 digit slices of Y, transposed  [S][m][ldk8]
    I_OZV,              //              scales: x(16) | a(ldm) | c(ldm) | y(ldm) | d(ldm)
    I_YSTAT,            //              per-slab column max / sum of squares of Y
    I_AMAX,             //              per-CTA partial max |X~|
    I_MMA,              // m x m x n products on the int8 engine: row-scaled digit slices of the left operand  [S][m][ld8]
    I_MMB,              //              row-scaled digit slices of the right operand (K-major, contraction over variables)
    I_MMC,              //              column-scaled digit slices of the operand contracted over its rows (factors)
    I_MMQ,              //              row-scaled digit slices of the m x m factor (ry or H)  [S][m][ldm8]
    I_MMV,              //              scales: col partial max (32 x ld) | col scale (ld) | row scales a, b, q (3 x ldm)
    I_COUNT
};

struct Slot {
    long long off, rows, cols, ld;
};

struct Layout {
    Slot slot[I_COUNT][2];
    long long total;
    long long ld, ldm, ldy;
    GemmPlan plan_k1, plan_k2, plan_mm, plan_mn;  // K1, K2, (m x m over n), (m x n over m)
    int nstrips;
    // split-integer modes (ozaki_i8.cuh)
    int S;                       // digits per operand, 0 = DMMA mode
    long long ld8, ldk8;         // byte leading dimensions of the X~/A slices and of the transposed Y slices (samples)
    int oz_splits, oz_chunk;     // split-K of the second contraction (over samples)
    int oz1_splits, oz1_chunk;   // split-K of the first contraction (over variables), only when row tiles are scarce
    int ystat_slabs;
    int radix;                   // 128: 7-bit signed digits (|d| <= 64); 254: full int8 range (|d| <= 127)
    int oz_kmax;                 // longest contraction one int32 accumulator group may see: 2^31 / ((R/2)^2 S)
    // the four m x m x n products of an iteration (ry, Qij, H, H W) on the same int8 engine (large m only)
    int mm_i8;                   // 0 = DMMA (dgemm_mma.cuh)
    long long ldm8;              // byte leading dimension of the digit slices of an m x m matrix
    int mm_splits, mm_chunk;     // split-K over the variables of the m x m outputs
    int mm_slabs, mm_slab_rows;  // row slabs of the per-column maximum
};

static int radix_for() {
    const char* env = getenv("LCX_SPLIT_RADIX");
    return (env && atoi(env) == 128) ? 128 : 254;  // 254: digits use the full int8 range (measured 100x tighter parity)
}

static int digits_for(int precision) {
    if (precision == LCX_PRECISION_FP64) return 0;
    const char* env = getenv("LCX_SPLIT_DIGITS");
    if (env && atoi(env) >= 3 && atoi(env) <= 7) return atoi(env);
    if (precision == LCX_PRECISION_FAST) return 3;          // 24 bits: fp32-equivalent products
    if (precision == LCX_PRECISION_FP64_SPLIT5) return 5;   // 40 bits
    if (precision == LCX_PRECISION_FP64_SPLIT7) return 7;   // 56 bits: finer than binary64's own 53-bit significand
    return 6;                                               // 48 bits: truncation at the level of binary64 rounding
}
// LCX_MM_I8=1 / 0 forces the int8 engine for the m x m x n products on / off; otherwise it is used from m = 384 factors
// (measured on a B200: m = 500, n = 50 000: direction 3.70 -> 3.19 ms, trial 2.25 -> 1.50 ms; m = 256, n = 8 000: 5 % slower --
// the products grow with m^2 n, the extra digit slicing with m n).
static int mm_i8_for(int S, int n, int m) {
    if (S <= 0 || round_up(m, 64) > (1 << 14)) return 0;
    const char* env = getenv("LCX_MM_I8");
    if (env) return atoi(env) != 0;
    return m >= 384 && n >= 2048;
}
constexpr int kYStatRows = 512;
constexpr int kAmaxCtas = 592;

static long long align16(long long v) { return round_up(v, 16); }

static Layout make_layout(long long Nl, int n, int m, int precision) {
    Layout L;
    memset(&L, 0, sizeof(L));
    L.S = digits_for(precision);
    L.radix = radix_for();
    if (L.S > 0) {
        const long long half = L.radix / 2;
        L.oz_kmax = (int)(((1LL << 31) / (half * half * L.S)) / 64 * 64);
        if (L.oz_kmax > 65536) L.oz_kmax = 65536;
    }
    L.ld = round_up(n, 16);
    L.ldm = round_up(m, 16);
    L.ldy = round_up(m, 8);
    // few samples (N << 128 * 148 rows): split the first contraction over the variables as well
    const bool k1_split = (long long)cdiv(Nl, 128) * cdiv(m, 128) < kSMs / 2;
    L.plan_k1 = plan_gemm((int)Nl, m, n, kSMs, k1_split ? 16 : 1, k1_split);
    L.plan_k2 = plan_gemm(n, m, (int)Nl, kSMs, kMaxSplitsX, true);
    L.plan_mm = plan_gemm(m, m, n, kSMs, kMaxSplitsSmall, true);
    L.plan_mn = plan_gemm(m, n, m, kSMs, 1, false);
    L.nstrips = cdiv(n, kStripCols);
    long long cur = 0;
    auto put = [&](int id, int set, long long rows, long long cols, long long ld) {
        L.slot[id][set] = Slot{cur, rows, cols, ld};
        cur = align16(cur + rows * ld);
    };
    const long long mn = m;
    for (int set = 0; set < 2; ++set) {
        put(LCX_A_W, set, mn, n, L.ld);
        put(LCX_A_RHO, set, mn, n, L.ld);
        put(LCX_A_INVRHO, set, mn, n, L.ld);
        put(LCX_A_RHOINVRHO, set, mn, n, L.ld);
        put(LCX_A_QIJ, set, mn, n, L.ld);
        put(LCX_A_SI, set, 1, n, L.ld);
        put(LCX_A_QISI2, set, 1, n, L.ld);
        put(LCX_A_RY, set, mn, m, L.ldm);
        put(LCX_A_UJ, set, 1, m, L.ldm);
    }
    auto put1 = [&](int id, long long rows, long long cols, long long ld) {
        put(id, 0, rows, cols, ld);
        L.slot[id][1] = L.slot[id][0];
    };
    put1(LCX_A_GRAD, mn, n, L.ld);
    put1(LCX_A_UPDATE, mn, n, L.ld);
    put1(LCX_A_RDIR, mn, n, L.ld);
    put1(LCX_A_D, mn + cdiv(m, L.ld), n, L.ld);  // D (m x ld) immediately followed by s (m values)
    put1(LCX_A_MI, mn, n, L.ld);
    put1(LCX_A_XZ, mn, n, L.ld);
    put1(LCX_A_XY, mn, n, L.ld);
    put1(LCX_A_X2Y, 1, n, L.ld);
    put1(LCX_A_IXY, 1, n, L.ld);
    put1(LCX_A_YJ2, 1, m, L.ldm);
    put1(LCX_A_IYX, 1, m, L.ldm);
    put1(LCX_A_TCS, 1, m, L.ldm);
    put1(LCX_A_TCDIRECT, 1, m, L.ldm);
    put1(LCX_A_CY, mn, m, L.ldm);
    put1(LCX_A_Y, Nl, m, L.ldy);
    put1(LCX_A_SCALARS, 1, 16, 16);
    put1(I_T, mn, n, L.ld);
    long long part = 0;
    if (L.plan_k2.splits > 1) part = max(part, (long long)L.plan_k2.splits * mn * L.ld);
    if (L.plan_mm.splits > 1) part = max(part, (long long)L.plan_mm.splits * mn * L.ldm);
    if (L.plan_k1.splits > 1) part = max(part, (long long)L.plan_k1.splits * Nl * L.ldy);
    put1(I_PART, 1, max(part, 16LL), max(part, 16LL));
    put1(I_COLSQ, L.plan_k1.grid.x, m, L.ldy);
    const long long spart = max(3LL * L.nstrips, (long long)m * cdiv(n, 256));
    put1(I_SPART, 1, spart, spart);
    put1(I_W2, 1, m, L.ldm);
    put1(I_BJ, 1, m, L.ldm);
    put1(I_F, 1, m, L.ldm);
    put1(I_UJDIAG, 1, m, L.ldm);
    put1(I_ROWMI, 1, m, L.ldm);
    put1(I_SQRTY, 1, m, L.ldm);
    put1(I_RYINV, mn, m, L.ldm);
    put1(I_AUG, mn, 2 * mn, 2 * mn);
    put1(I_STATUS, 1, 2, 2);
    L.ystat_slabs = cdiv(Nl, kYStatRows);
    put1(I_OZV, 1, 16 + 4 * L.ldm, 16 + 4 * L.ldm);
    put1(I_YSTAT, (long long)L.ystat_slabs * 2, m, L.ldm);
    if (L.S > 0) {
        L.ld8 = round_up(n, 128);
        L.ldk8 = round_up(Nl, 128);
        // second contraction: 128 x 64 tiles over (variables x factors), split over samples to fill whole waves;
        // at most oz_kmax samples per split keeps every int32 accumulator exact
        const long long tiles = (long long)cdiv(n, oz::kBM) * cdiv(m, oz::bn_max(L.S));
        const int kblocks = cdiv(Nl, oz::kBK);
        // time model in units of one 64-deep K block: waves x (K blocks per CTA + fixed prologue/TMEM-drain/store cost
        // of ~16 blocks) + the partial-buffer round trip; measured at 12.5k and 100k rows per GPU
        int best = 1;
        double best_cost = 1e300;
        const int smin = cdiv(Nl, L.oz_kmax), smax = (int)min(64LL, (long long)max(1, kblocks / 8));
        for (int sp = smin; sp <= max(smin, smax); ++sp) {
            const long long ctas = tiles * sp;
            const long long waves = (ctas + kSMs - 1) / kSMs;
            const double kb = ceil((double)kblocks / sp);
            const double cost = (double)waves * (kb + 16.0) + (sp > 1 ? 1.5 * sp : 0.0);
            if (cost < best_cost - 1e-9) { best_cost = cost; best = sp; }
        }
        if (const char* env = getenv("LCX_OZ_SPLITS")) {  // experiment override; never below the int32-exact minimum
            const int v = atoi(env);
            if (v >= smin && v <= kblocks) best = v;
        }
        L.oz_chunk = (int)round_up(cdiv(Nl, best), oz::kBK);
        L.oz_splits = cdiv(Nl, L.oz_chunk);
        {   // first contraction: same cost model over its (row tile x factor tile) grid
            const long long tiles1 = (long long)cdiv(Nl, oz::kBM) * cdiv(m, oz::bn_max(L.S));
            const int kblocks1 = cdiv(n, oz::kBK);
            int b1 = cdiv(n, L.oz_kmax);
            double c1best = 1e300;
            const int s1min = cdiv(n, L.oz_kmax);  // int32 exactness of every accumulator group
            for (int sp = s1min; sp <= max(s1min, min(8, kblocks1 / 16)); ++sp) {
                const long long waves = (tiles1 * sp + kSMs - 1) / kSMs;
                const double cost = (double)waves * (ceil((double)kblocks1 / sp) + 16.0) + (sp > 1 ? 4.0 * sp : 0.0);
                if (cost < c1best - 1e-9) { c1best = cost; b1 = sp; }
            }
            L.oz1_chunk = (int)round_up(cdiv(n, b1), oz::kBK);
            L.oz1_splits = cdiv(n, L.oz1_chunk);
        }
        const long long part = max((long long)L.oz_splits * mn * L.ld, L.oz1_splits > 1 ? (long long)L.oz1_splits * Nl * L.ldy : 0LL);
        if (part > L.slot[I_PART][0].cols) {  // grow the split-K partial buffer (it is the last big slot before these)
            put1(I_PART, 1, part, part);
        }
        const long long xs8 = ((long long)L.S * Nl * L.ld8 + 7) / 8;   // int8 planes counted in doubles (64-bit sizes:
        const long long as8 = ((long long)L.S * mn * L.ld8 + 7) / 8;   // the target shape has 1.5e10 doubles of planes)
        const long long ys8 = ((long long)L.S * mn * L.ldk8 + 7) / 8;
        put1(I_XS, 1, xs8, xs8);
        put1(I_AS, 1, as8, as8);
        put1(I_YS, 1, ys8, ys8);
        put1(I_AMAX, 1, kAmaxCtas, kAmaxCtas);
        L.mm_i8 = mm_i8_for(L.S, n, m);
        if (L.mm_i8) {
            L.ldm8 = round_up(m, 128);
            {   // m x m outputs, contraction over the variables: split to fill the SMs, never beyond the int32-exact length
                const long long tiles_mm = (long long)cdiv(m, oz::kBM) * cdiv(m, oz::bn_max(L.S));
                const int kblocks_mm = cdiv(n, oz::kBK);
                const int smin_mm = cdiv(n, L.oz_kmax);
                int bmm = smin_mm;
                double cbest = 1e300;
                for (int sp = smin_mm; sp <= max(smin_mm, min(64, kblocks_mm / 8)); ++sp) {
                    const long long waves = (tiles_mm * sp + kSMs - 1) / kSMs;
                    const double cost = (double)waves * (ceil((double)kblocks_mm / sp) + 16.0) + 1.5 * sp;
                    if (cost < cbest - 1e-9) { cbest = cost; bmm = sp; }
                }
                L.mm_chunk = (int)round_up(cdiv(n, bmm), oz::kBK);
                L.mm_splits = cdiv(n, L.mm_chunk);
            }
            L.mm_slabs = (int)min(32LL, (long long)cdiv(m, 8));
            L.mm_slab_rows = cdiv(m, L.mm_slabs);
            const long long need = (long long)L.mm_splits * mn * L.ldm;
            if (need > L.slot[I_PART][0].cols) put1(I_PART, 1, need, need);
            const long long pl8 = ((long long)L.S * mn * L.ld8 + 7) / 8;
            const long long q8 = ((long long)L.S * mn * L.ldm8 + 7) / 8;
            put1(I_MMA, 1, pl8, pl8);
            put1(I_MMB, 1, pl8, pl8);
            put1(I_MMC, 1, pl8, pl8);
            put1(I_MMQ, 1, q8, q8);
            put1(I_MMV, 1, 33 * L.ld + 3 * L.ldm, 33 * L.ld + 3 * L.ldm);
        }
    }
    L.total = cur;
    return L;
}

struct lcx_session {
    int device, precision;
    cudaStream_t stream;
    lcx_allreduce_fn hook;
    void* hook_user;
    long long launches;
    double* mailbox;  // pinned host, 16 doubles
    bool bound;
    const double* xt;
    long long Nl, Nt, ldx;
    int n, m;
    double* ws;
    Layout L;
    int cur;  // which physical set is "set 0" (current)
    // optional device-side timing of the two X contractions (bench.py roofline); events are
    // recorded on the session stream around each launch and resolved at lcx_profile_read
    bool prof_on;
    cudaEvent_t* prof_ev;      // 4 per pair: [0] before K1, [1] after K1, [3] before K2, [2] after K2 (+ split reduction)
    int prof_pending, prof_cap;
    double prof_k1_ms, prof_k2_ms;
    long long prof_pairs;
    // sample sharding over NVLink peers (fused_allreduce.cuh); peers.world <= 1 means off
    far::Peers peers;
    unsigned long long ar_calls;
    // split-integer modes: TMA descriptors over the digit slices
    CUtensorMap map_x_k1, map_a_k1, map_a_k1_tail, map_x_k2, map_y_k2, map_y_k2_tail;
    int oz_bn_tail;   // width of the last factor tile of the first contraction (multiple of 16)
    // m x m x n products on the int8 engine (L.mm_i8)
    CUtensorMap map_mm_a, map_mm_b, map_mm_b_tail, map_mn_c, map_mn_q, map_mn_q_tail;
    int8_t* mma() const { return (int8_t*)(ws + L.slot[I_MMA][0].off); }
    int8_t* mmb() const { return (int8_t*)(ws + L.slot[I_MMB][0].off); }
    int8_t* mmc() const { return (int8_t*)(ws + L.slot[I_MMC][0].off); }
    int8_t* mmq() const { return (int8_t*)(ws + L.slot[I_MMQ][0].off); }
    double* mm_colpart() const { return ws + L.slot[I_MMV][0].off; }
    double* mm_colscale() const { return ws + L.slot[I_MMV][0].off + 32 * L.ld; }
    double* mm_scale_a() const { return ws + L.slot[I_MMV][0].off + 33 * L.ld; }
    double* mm_scale_b() const { return ws + L.slot[I_MMV][0].off + 33 * L.ld + L.ldm; }
    double* mm_scale_q() const { return ws + L.slot[I_MMV][0].off + 33 * L.ld + 2 * L.ldm; }
    int8_t* xs() const { return (int8_t*)(ws + L.slot[I_XS][0].off); }
    int8_t* as() const { return (int8_t*)(ws + L.slot[I_AS][0].off); }
    int8_t* ys() const { return (int8_t*)(ws + L.slot[I_YS][0].off); }
    double* oz_xscale() const { return ws + L.slot[I_OZV][0].off; }
    double* oz_ascale() const { return ws + L.slot[I_OZV][0].off + 16; }
    double* oz_cscale() const { return ws + L.slot[I_OZV][0].off + 16 + L.ldm; }
    double* oz_yscale() const { return ws + L.slot[I_OZV][0].off + 16 + 2 * L.ldm; }
    double* oz_dscale() const { return ws + L.slot[I_OZV][0].off + 16 + 3 * L.ldm; }

    double* ptr(int id, int set = 0) const {
        const int phys = (id <= LCX_A_UJ) ? (set ^ cur) : 0;
        return ws + L.slot[id][phys].off;
    }
    long long off(int id, int set = 0) const {
        const int phys = (id <= LCX_A_UJ) ? (set ^ cur) : 0;
        return L.slot[id][phys].off;
    }
};

#define S_REQUIRE_BOUND(s)                                                        \
    do {                                                                          \
        if (!(s)) return fail(LCX_ERR_ARG, "session", "null session");            \
        if (!(s)->bound) return fail(LCX_ERR_STATE, "session", "no bound problem"); \
        LCX_CUDA(cudaSetDevice((s)->device));                                     \
    } while (0)

#define LAUNCHED(s) ((s)->launches++)

static int combine_and_allreduce(lcx_session* s, const double* part, int splits, long long stride, int rows, int cols,
                                 long long ld, double* dst_body, double* tail, int ntail);

static dim3 grid_mn(int m, int n) { return dim3(cdiv(n, 256), m); }

// ---- lifecycle ---------------------------------------------------------------------------------
extern "C" int lcx_version(void) { return 100; }
extern "C" const char* lcx_last_error(void) { return g_err; }

extern "C" int lcx_session_create(lcx_session** out, int device, int precision) {
    LCX_REQUIRE(out != nullptr, "out is null");
    LCX_REQUIRE(precision >= LCX_PRECISION_FP64 && precision <= LCX_PRECISION_FP64_SPLIT7, "unknown precision mode");
    LCX_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    LCX_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) return fail(LCX_ERR_STATE, "lcx_session_create", "this library is built for sm_100a (B200) only");
    lcx_session* s = new lcx_session();
    memset(s, 0, sizeof(*s));
    s->device = device;
    s->precision = precision;
    s->stream = 0;
    LCX_CUDA(cudaHostAlloc((void**)&s->mailbox, 16 * sizeof(double), cudaHostAllocDefault));
    *out = s;
    return 0;
}

extern "C" int lcx_profile_enable(lcx_session* s, int on) {
    LCX_REQUIRE(s != nullptr, "null session");
    LCX_CUDA(cudaSetDevice(s->device));
    if (on && s->prof_ev == nullptr) {
        s->prof_cap = 4096;
        s->prof_ev = new cudaEvent_t[4 * s->prof_cap];
        for (int i = 0; i < 4 * s->prof_cap; ++i) LCX_CUDA(cudaEventCreate(&s->prof_ev[i]));
    }
    s->prof_on = on != 0;
    return 0;
}

extern "C" int lcx_profile_read(lcx_session* s, double* k1_ms, double* k2_ms, long long* pairs, int reset) {
    LCX_REQUIRE(s != nullptr, "null session");
    LCX_CUDA(cudaSetDevice(s->device));
    LCX_CUDA(cudaStreamSynchronize(s->stream));
    for (int i = 0; i < s->prof_pending; ++i) {
        cudaEvent_t* ev = s->prof_ev + 4 * i;
        float a = 0.f, b = 0.f;
        LCX_CUDA(cudaEventElapsedTime(&a, ev[0], ev[1]));
        LCX_CUDA(cudaEventElapsedTime(&b, ev[3], ev[2]));
        s->prof_k1_ms += a;
        s->prof_k2_ms += b;
        s->prof_pairs++;
    }
    s->prof_pending = 0;
    if (k1_ms) *k1_ms = s->prof_k1_ms;
    if (k2_ms) *k2_ms = s->prof_k2_ms;
    if (pairs) *pairs = s->prof_pairs;
    if (reset) {
        s->prof_k1_ms = s->prof_k2_ms = 0.0;
        s->prof_pairs = 0;
    }
    return 0;
}

extern "C" int lcx_session_destroy(lcx_session* s) {
    if (!s) return 0;
    if (s->prof_ev) {
        for (int i = 0; i < 4 * s->prof_cap; ++i) cudaEventDestroy(s->prof_ev[i]);
        delete[] s->prof_ev;
    }
    if (s->mailbox) cudaFreeHost(s->mailbox);
    delete s;
    return 0;
}

extern "C" int lcx_set_stream(lcx_session* s, void* cuda_stream) {
    LCX_REQUIRE(s != nullptr, "null session");
    s->stream = (cudaStream_t)cuda_stream;
    return 0;
}

extern "C" int lcx_set_allreduce(lcx_session* s, lcx_allreduce_fn fn, void* user) {
    LCX_REQUIRE(s != nullptr, "null session");
    s->hook = fn;
    s->hook_user = user;
    return 0;
}

extern "C" long long lcx_peer_buffer_doubles(int n_vars, int n_factors) {
    if (n_vars <= 0 || n_factors <= 0) return -1;
    const long long count = (long long)n_factors * round_up(n_vars, 16) + round_up(n_factors, 16);
    return 3 * count + 4 * far::kMaxRanks;
}

extern "C" int lcx_set_peer_allreduce(lcx_session* s, int world, int rank, void* const* bases, long long buffer_doubles) {
    S_REQUIRE_BOUND(s);
    if (world <= 1 || bases == nullptr) {
        s->peers.world = 0;
        return 0;
    }
    LCX_REQUIRE(world <= far::kMaxRanks && rank >= 0 && rank < world, "bad world / rank (at most 8 ranks)");
    LCX_REQUIRE(buffer_doubles >= lcx_peer_buffer_doubles(s->n, s->m), "peer buffer too small (see lcx_peer_buffer_doubles)");
    for (int r = 0; r < world; ++r) {
        LCX_REQUIRE(bases[r] != nullptr && ((uintptr_t)bases[r] % 16 == 0), "bad peer pointer");
        s->peers.base[r] = (double*)bases[r];
    }
    s->peers.world = world;
    s->peers.rank = rank;
    s->peers.count = (long long)s->m * s->L.ld + s->L.ldm;
    s->ar_calls = 0;
    return 0;
}

extern "C" int lcx_launch_count(lcx_session* s, long long* launches) {
    LCX_REQUIRE(s != nullptr && launches != nullptr, "null argument");
    *launches = s->launches;
    return 0;
}

// ---- layout ------------------------------------------------------------------------------------
extern "C" long long lcx_ld(int n_vars) { return round_up(n_vars, 16); }
extern "C" long long lcx_ldy(int n_factors) { return round_up(n_factors, 8); }

extern "C" long long lcx_workspace_doubles(long long n_rows_local, int n_vars, int n_factors, int precision) {
    if (n_rows_local < 0 || n_vars <= 0 || n_factors <= 0 || precision < 0 || precision > LCX_PRECISION_FP64_SPLIT7) return -1;
    return make_layout(n_rows_local, n_vars, n_factors, precision).total;
}

// ---- split-integer plumbing (ozaki_i8.cuh) -------------------------------------------------------
template <int S>
static int oz_slice_x_t(lcx_session* s) {
    const Layout& L = s->L;
    oz::absmax_partial_kernel<<<kAmaxCtas, 256, 0, s->stream>>>(s->xt, s->ldx, s->Nl, s->n, s->ws + L.slot[I_AMAX][0].off);
    LAUNCHED(s);
    oz::absmax_finish_kernel<<<1, 256, 0, s->stream>>>(s->ws + L.slot[I_AMAX][0].off, kAmaxCtas, s->oz_xscale());
    LAUNCHED(s);
    dim3 grid((unsigned)s->Nl, cdiv(L.ld8, 4 * 128));
    oz::slice_rows_kernel<S><<<grid, 128, 0, s->stream>>>(s->xt, s->ldx, (int)s->Nl, s->n, nullptr, s->oz_xscale(), s->xs(), L.ld8,
                                                        s->Nl * L.ld8, (double)L.radix);
    LAUNCHED(s);
    LCX_CUDA(cudaGetLastError());
    return 0;
}

static int oz_prepare(lcx_session* s, bool streamed) {
    const Layout& L = s->L;
    LCX_REQUIRE(L.oz1_chunk <= L.oz_kmax && L.oz_chunk <= L.oz_kmax, "contraction chunk exceeds the int32-exact length");
    if (!streamed) switch (L.S) {
        case 3: LCX_TRY(oz_slice_x_t<3>(s)); break;
        case 4: LCX_TRY(oz_slice_x_t<4>(s)); break;
        case 5: LCX_TRY(oz_slice_x_t<5>(s)); break;
        case 6: LCX_TRY(oz_slice_x_t<6>(s)); break;
        case 7: LCX_TRY(oz_slice_x_t<7>(s)); break;
        default: return fail(LCX_ERR_STATE, "oz_prepare", "bad digit count");
    }
    // X~ slices as the M operand of Y = X~ A^T (K-major: inner = variables, 64 B boxes) and as the M operand of
    // D = X~^T Y (MN-major: inner = variables, 128 B boxes over 64 sample rows); the factor-side operands are K-major:
    // A slices (inner = variables) and the transposed Y slices (inner = samples).
    LCX_TRY(oz::make_slice_map(&s->map_x_k1, s->xs(), s->n, s->Nl, L.S, L.ld8, s->Nl * L.ld8, oz::kBK, oz::kBM, false));
    const int bnm = oz::bn_max(L.S);
    LCX_TRY(oz::make_slice_map(&s->map_a_k1, s->as(), s->n, s->m, L.S, L.ld8, (long long)s->m * L.ld8, oz::kBK, bnm, false));
    s->oz_bn_tail = (int)round_up(s->m - (cdiv(s->m, bnm) - 1) * bnm, 16);
    LCX_TRY(oz::make_slice_map(&s->map_a_k1_tail, s->as(), s->n, s->m, L.S, L.ld8, (long long)s->m * L.ld8, oz::kBK, s->oz_bn_tail,
                               false));
    LCX_TRY(oz::make_slice_map(&s->map_x_k2, s->xs(), s->n, s->Nl, L.S, L.ld8, s->Nl * L.ld8, oz::kBM, oz::kBK, true));
    LCX_TRY(oz::make_slice_map(&s->map_y_k2, s->ys(), s->Nl, s->m, L.S, L.ldk8, (long long)s->m * L.ldk8, oz::kBK, bnm, false));
    LCX_TRY(oz::make_slice_map(&s->map_y_k2_tail, s->ys(), s->Nl, s->m, L.S, L.ldk8, (long long)s->m * L.ldk8, oz::kBK,
                               s->oz_bn_tail, false));
    if (L.mm_i8) {
        LCX_REQUIRE(L.mm_chunk <= L.oz_kmax && round_up(s->m, oz::kBK) <= L.oz_kmax, "contraction chunk exceeds the int32-exact length");
        const long long st_n = (long long)s->m * L.ld8, st_q = (long long)s->m * L.ldm8;
        // ry = W rho^T, H = T rinv^T: both operands K-major over the variables (M side 128-row boxes, N side bn-row boxes)
        LCX_TRY(oz::make_slice_map(&s->map_mm_a, s->mma(), s->n, s->m, L.S, L.ld8, st_n, oz::kBK, oz::kBM, false));
        LCX_TRY(oz::make_slice_map(&s->map_mm_b, s->mmb(), s->n, s->m, L.S, L.ld8, st_n, oz::kBK, bnm, false));
        LCX_TRY(oz::make_slice_map(&s->map_mm_b_tail, s->mmb(), s->n, s->m, L.S, L.ld8, st_n, oz::kBK, s->oz_bn_tail, false));
        // Qij = ry rinv, grad += H W: M side = variables of the m x n operand (MN-major, contraction over its m rows),
        // N side = the m x m factor, K-major
        LCX_TRY(oz::make_slice_map(&s->map_mn_c, s->mmc(), s->n, s->m, L.S, L.ld8, st_n, oz::kBM, oz::kBK, true));
        LCX_TRY(oz::make_slice_map(&s->map_mn_q, s->mmq(), s->m, s->m, L.S, L.ldm8, st_q, oz::kBK, bnm, false));
        LCX_TRY(oz::make_slice_map(&s->map_mn_q_tail, s->mmq(), s->m, s->m, L.S, L.ldm8, st_q, oz::kBK, s->oz_bn_tail, false));
    }
    return 0;
}

static int oz_cluster() {  // LCX_OZ_CLUSTER=1|2|4 overrides the cluster size of the split-integer contractions
    const char* env = getenv("LCX_OZ_CLUSTER");
    return env ? atoi(env) : 2;  // pairs: multicast does not lower the bytes delivered per SM, and clusters of 4 fit only 132 SMs
}

template <int S>
static int oz_pair_t(lcx_session* s, const double* A, double* svec, cudaEvent_t* ev, bool first_only, bool want_tail) {
    const Layout& L = s->L;
    const int m = s->m, n = s->n;
    double* Y = s->ptr(LCX_A_Y);
    double* D = s->ptr(LCX_A_D);
    // ---- Y = X~ A^T ----
    oz::row_scale_kernel<<<m, 256, 0, s->stream>>>(A, L.ld, n, s->oz_ascale());
    LAUNCHED(s);
    oz::mul_scale_kernel<<<cdiv(m, 128), 128, 0, s->stream>>>(s->oz_xscale(), s->oz_ascale(), s->oz_cscale(), m);
    LAUNCHED(s);
    oz::slice_rows_kernel<S><<<dim3(m, cdiv(L.ld8, 4 * 128)), 128, 0, s->stream>>>(A, L.ld, m, n, s->oz_ascale(), nullptr, s->as(), L.ld8,
                                                                              (long long)m * L.ld8, (double)L.radix);
    LAUNCHED(s);
    {
        oz::GemmParams p;
        memset(&p, 0, sizeof(p));
        const bool split1 = L.oz1_splits > 1;
        p.C = split1 ? s->ptr(I_PART) : Y;
        p.ldc = L.ldy; p.c_split_stride = split1 ? s->Nl * L.ldy : 0;
        p.col_scale = s->oz_cscale();
        p.inv_radix = 1.0 / (double)L.radix;
        p.rows = (int)s->Nl; p.cols = m; p.k_total = n; p.k_chunk = L.oz1_chunk;
        p.bn_tail = s->oz_bn_tail;
        LCX_TRY((oz::launch_oz_gemm<S, true>(s->map_x_k1, s->map_a_k1, s->map_a_k1_tail, p,
                                             dim3(cdiv(m, oz::bn_max(S)), cdiv(s->Nl, oz::kBM), L.oz1_splits), s->stream, oz_cluster())));
        LAUNCHED(s);
        if (split1) {
            LCX_TRY(launch_reduce_splits(s->ptr(I_PART), L.oz1_splits, s->Nl * L.ldy, Y, (int)s->Nl, m, L.ldy, s->stream));
            LAUNCHED(s);
        }
    }
    if (ev) LCX_CUDA(cudaEventRecord(ev[1], s->stream));
    if (ev) LCX_CUDA(cudaEventRecord(ev[3], s->stream));
    // ---- column max / sum of squares of Y, digit slices of Y ----
    double* ystat = s->ws + L.slot[I_YSTAT][0].off;
    oz::y_stats_kernel<<<dim3(cdiv(m, 32), L.ystat_slabs), dim3(32, 8), 0, s->stream>>>(Y, L.ldy, s->Nl, m, kYStatRows, ystat, L.ldm);
    LAUNCHED(s);
    oz::y_stats_finish_kernel<<<m, 256, 0, s->stream>>>(ystat, L.ystat_slabs, L.ldm, m, s->oz_xscale(), svec, s->oz_yscale(),
                                                                  s->oz_dscale());
    LAUNCHED(s);
    if (first_only) {  // _norm (:215-228): only Y and its column sums of squares are needed
        LCX_CUDA(cudaGetLastError());
        return 0;
    }
    oz::slice_cols_t_kernel<S><<<dim3((unsigned)cdiv(s->Nl, 128), cdiv(m, 32)), dim3(32, 8), 0, s->stream>>>(
        Y, L.ldy, s->Nl, m, s->oz_yscale(), s->ys(), L.ldk8, (long long)m * L.ldk8, (double)L.radix);
    LAUNCHED(s);
    // ---- D = (X~^T Y)^T: tiles of 128 variables x 64 factors, stored factor-major, split over samples ----
    {
        oz::GemmParams p;
        memset(&p, 0, sizeof(p));
        const bool split = L.oz_splits > 1;
        p.C = split ? s->ptr(I_PART) : D;
        p.ldc = L.ld; p.c_split_stride = split ? (long long)m * L.ld : 0;
        p.col_scale = s->oz_dscale();
        p.inv_radix = 1.0 / (double)L.radix;
        p.rows = n; p.cols = m; p.k_total = (int)s->Nl; p.k_chunk = L.oz_chunk;
        p.bn_tail = s->oz_bn_tail;
        p.trans_out = 1;
        LCX_TRY((oz::launch_oz_gemm<S, false>(s->map_x_k2, s->map_y_k2, s->map_y_k2_tail, p,
                                              dim3(cdiv(m, oz::bn_max(S)), cdiv(n, oz::kBM), L.oz_splits), s->stream, oz_cluster())));
        LAUNCHED(s);
        LCX_TRY(combine_and_allreduce(s, split ? s->ptr(I_PART) : D, split ? L.oz_splits : 1, (long long)m * L.ld, m, n, L.ld, D, svec,
                                      want_tail ? m : 0));
    }
    LCX_CUDA(cudaGetLastError());
    return 0;
}

static int oz_pair(lcx_session* s, const double* A, double* svec, cudaEvent_t* ev, bool first_only, bool want_tail) {
    switch (s->L.S) {
        case 3: return oz_pair_t<3>(s, A, svec, ev, first_only, want_tail);
        case 4: return oz_pair_t<4>(s, A, svec, ev, first_only, want_tail);
        case 5: return oz_pair_t<5>(s, A, svec, ev, first_only, want_tail);
        case 6: return oz_pair_t<6>(s, A, svec, ev, first_only, want_tail);
        case 7: return oz_pair_t<7>(s, A, svec, ev, first_only, want_tail);
    }
    return fail(LCX_ERR_STATE, "oz_pair", "bad digit count");
}

// ---- the m x m x n products of an iteration on the int8 engine (L.mm_i8; same kernel, same digit scheme) ----------
// out (m x m) = left right^T over the variables, both m x n: each operand gets one exponent per factor row; split-K
// partials are combined in fixed order and np.fill_diagonal is applied there (raw diagonal -> diag_out).
template <int S>
static int oz_square_t(lcx_session* s, const double* left, const double* right, double* out, double diag_value,
                       double* diag_out) {
    const Layout& L = s->L;
    const int m = s->m, n = s->n;
    const dim3 gs(m, cdiv(L.ld8, 4 * 128));
    const long long st_n = (long long)m * L.ld8;
    oz::row_scale_kernel<<<m, 256, 0, s->stream>>>(left, L.ld, n, s->mm_scale_a());
    LAUNCHED(s);
    oz::slice_rows_kernel<S><<<gs, 128, 0, s->stream>>>(left, L.ld, m, n, s->mm_scale_a(), nullptr, s->mma(), L.ld8, st_n,
                                                       (double)L.radix);
    LAUNCHED(s);
    oz::row_scale_kernel<<<m, 256, 0, s->stream>>>(right, L.ld, n, s->mm_scale_b());
    LAUNCHED(s);
    oz::slice_rows_kernel<S><<<gs, 128, 0, s->stream>>>(right, L.ld, m, n, s->mm_scale_b(), nullptr, s->mmb(), L.ld8, st_n,
                                                       (double)L.radix);
    LAUNCHED(s);
    oz::GemmParams p;
    memset(&p, 0, sizeof(p));
    const long long out_count = (long long)m * L.ldm;
    p.C = s->ptr(I_PART);
    p.ldc = L.ldm; p.c_split_stride = out_count;
    p.row_scale = s->mm_scale_a();
    p.col_scale = s->mm_scale_b();
    p.inv_radix = 1.0 / (double)L.radix;
    p.rows = m; p.cols = m; p.k_total = n; p.k_chunk = L.mm_chunk;
    p.bn_tail = s->oz_bn_tail;
    LCX_TRY((oz::launch_oz_gemm<S, true>(s->map_mm_a, s->map_mm_b, s->map_mm_b_tail, p,
                                         dim3(cdiv(m, oz::bn_max(S)), cdiv(m, oz::kBM), L.mm_splits), s->stream, oz_cluster())));
    LAUNCHED(s);
    LCX_TRY(launch_reduce_splits(s->ptr(I_PART), L.mm_splits, out_count, out, m, m, L.ldm, s->stream,
                                 diag_out ? diag_out : s->ptr(I_F), diag_value));
    LAUNCHED(s);
    LCX_CUDA(cudaGetLastError());
    return 0;
}

// out (m x n, factor-major) = c_add + Q V with Q m x m and V m x n: the contraction runs over V's rows, so V gets one
// exponent per COLUMN (variable) and Q one per row; tiles of 128 variables x 64 factors like the second X contraction.
// unit_diag: Q has an exact unit diagonal (ry after np.fill_diagonal, :263).  Its digits would be spent on that 1 while the
// off-diagonal correlations are 1e-2 and below, so the product runs on Q - I and the caller passes c_add = V.
template <int S>
static int oz_mn_t(lcx_session* s, const double* Q, const double* V, double* out, const double* c_add, bool unit_diag) {
    const Layout& L = s->L;
    const int m = s->m, n = s->n;
    oz::col_absmax_partial_kernel<<<dim3(cdiv(n, 512), L.mm_slabs), 256, 0, s->stream>>>(V, L.ld, m, n, L.mm_slab_rows,
                                                                                       s->mm_colpart(), L.ld);
    LAUNCHED(s);
    oz::col_scale_finish_kernel<<<cdiv(n, 256), 256, 0, s->stream>>>(s->mm_colpart(), L.mm_slabs, L.ld, n, s->mm_colscale());
    LAUNCHED(s);
    oz::slice_colscaled_kernel<S><<<dim3(m, cdiv(L.ld8, 4 * 128)), 128, 0, s->stream>>>(V, L.ld, m, n, s->mm_colscale(), s->mmc(),
                                                                                      L.ld8, (long long)m * L.ld8, (double)L.radix);
    LAUNCHED(s);
    oz::row_scale_kernel<<<m, 256, 0, s->stream>>>(Q, L.ldm, m, s->mm_scale_q(), unit_diag ? 1 : 0);
    LAUNCHED(s);
    const dim3 gq(m, cdiv(L.ldm8, 4 * 128));
    if (unit_diag)
        oz::slice_rows_kernel<S, true><<<gq, 128, 0, s->stream>>>(Q, L.ldm, m, m, s->mm_scale_q(), nullptr, s->mmq(), L.ldm8,
                                                                 (long long)m * L.ldm8, (double)L.radix);
    else
        oz::slice_rows_kernel<S><<<gq, 128, 0, s->stream>>>(Q, L.ldm, m, m, s->mm_scale_q(), nullptr, s->mmq(), L.ldm8,
                                                           (long long)m * L.ldm8, (double)L.radix);
    LAUNCHED(s);
    oz::GemmParams p;
    memset(&p, 0, sizeof(p));
    p.C = out;
    p.ldc = L.ld; p.c_split_stride = 0;
    p.row_scale = s->mm_colscale();
    p.col_scale = s->mm_scale_q();
    p.inv_radix = 1.0 / (double)L.radix;
    p.rows = n; p.cols = m; p.k_total = m; p.k_chunk = (int)round_up(m, oz::kBK);
    p.bn_tail = s->oz_bn_tail;
    p.trans_out = 1;
    p.c_add = c_add;
    LCX_TRY((oz::launch_oz_gemm<S, false, true>(s->map_mn_c, s->map_mn_q, s->map_mn_q_tail, p,
                                          dim3(cdiv(m, oz::bn_max(S)), cdiv(n, oz::kBM), 1), s->stream, oz_cluster())));
    LAUNCHED(s);
    LCX_CUDA(cudaGetLastError());
    return 0;
}

static int oz_square(lcx_session* s, const double* left, const double* right, double* out, double diag_value, double* diag_out) {
    switch (s->L.S) {
        case 3: return oz_square_t<3>(s, left, right, out, diag_value, diag_out);
        case 4: return oz_square_t<4>(s, left, right, out, diag_value, diag_out);
        case 5: return oz_square_t<5>(s, left, right, out, diag_value, diag_out);
        case 6: return oz_square_t<6>(s, left, right, out, diag_value, diag_out);
        case 7: return oz_square_t<7>(s, left, right, out, diag_value, diag_out);
    }
    return fail(LCX_ERR_STATE, "oz_square", "bad digit count");
}

static int oz_mn(lcx_session* s, const double* Q, const double* V, double* out, const double* c_add, bool unit_diag) {
    switch (s->L.S) {
        case 3: return oz_mn_t<3>(s, Q, V, out, c_add, unit_diag);
        case 4: return oz_mn_t<4>(s, Q, V, out, c_add, unit_diag);
        case 5: return oz_mn_t<5>(s, Q, V, out, c_add, unit_diag);
        case 6: return oz_mn_t<6>(s, Q, V, out, c_add, unit_diag);
        case 7: return oz_mn_t<7>(s, Q, V, out, c_add, unit_diag);
    }
    return fail(LCX_ERR_STATE, "oz_mn", "bad digit count");
}

extern "C" int lcx_bind(lcx_session* s, const double* xt, long long n_rows_local, long long n_rows_total, int n_vars,
                        long long ldx, int n_factors, double* workspace, long long workspace_doubles) {
    LCX_REQUIRE(s != nullptr, "null session");
    LCX_REQUIRE(workspace != nullptr, "null device pointer");
    LCX_REQUIRE(xt != nullptr || s->precision != LCX_PRECISION_FP64,
                "xt may be NULL only in the split modes (digit planes are then filled by lcx_slice_block)");
    LCX_REQUIRE(n_rows_local > 0 && n_rows_local < (1LL << 31) && n_rows_total >= n_rows_local, "bad row counts");
    LCX_REQUIRE(n_vars > 0 && n_factors > 0, "bad shape");
    LCX_REQUIRE(ldx >= n_vars && ldx % 2 == 0, "ldx must be even and >= n_vars");
    LCX_REQUIRE(((uintptr_t)xt % 16 == 0) && ((uintptr_t)workspace % 128 == 0), "misaligned device pointer");
    const bool streamed = (xt == nullptr);
    Layout L = make_layout(n_rows_local, n_vars, n_factors, s->precision);
    LCX_REQUIRE(workspace_doubles >= L.total, "workspace too small (see lcx_workspace_doubles)");
    LCX_CUDA(cudaSetDevice(s->device));
    s->xt = xt;
    s->Nl = n_rows_local;
    s->Nt = n_rows_total;
    s->ldx = ldx;
    s->n = n_vars;
    s->m = n_factors;
    s->ws = workspace;
    s->L = L;
    s->cur = 0;
    s->bound = true;
    // everything except Y starts at zero so padding never carries NaNs into an all-reduce
    const long long y_off = L.slot[LCX_A_Y][0].off;
    LCX_CUDA(cudaMemsetAsync(workspace, 0, (size_t)y_off * sizeof(double), s->stream));
    const long long y_end = align16(y_off + n_rows_local * L.ldy);
    LCX_CUDA(cudaMemsetAsync(workspace + y_end, 0, (size_t)(L.total - y_end) * sizeof(double), s->stream));
    if (L.S > 0) {
        LCX_TRY(oz_prepare(s, streamed));  // digit slices of X~; after this X~ itself is never read again in the split modes,
        LCX_CUDA(cudaStreamSynchronize(s->stream));  // so the caller may release it as soon as lcx_bind returns
        s->xt = nullptr;
    }
    return 0;
}

// ---- streamed digit slicing (X~ never materialised as a whole: the 1M x 20k target on one GPU) -------------------
extern "C" int lcx_set_x_scale(lcx_session* s, double max_abs) {
    S_REQUIRE_BOUND(s);
    LCX_REQUIRE(s->L.S > 0, "only meaningful in the split modes");
    LCX_REQUIRE(!(max_abs < 0.0), "max_abs must be a non-negative bound on |X~|");
    int e = 0;
    double scale = 1.0;
    if (!(max_abs <= DBL_MAX)) {
        scale = NAN;  // non-finite data: every product comes out NaN, as in the reference's float64 path
    } else if (max_abs > 0.0) {
        frexp(max_abs, &e);
        scale = ldexp(1.0, e + 1);
    }
    LCX_CUDA(cudaMemcpyAsync(s->oz_xscale(), &scale, sizeof(double), cudaMemcpyHostToDevice, s->stream));
    LCX_CUDA(cudaStreamSynchronize(s->stream));
    return 0;
}

template <int S>
static int oz_slice_block_t(lcx_session* s, const double* xt, long long row0, long long rows, long long ldx) {
    const Layout& L = s->L;
    dim3 grid((unsigned)rows, cdiv(L.ld8, 4 * 128));
    oz::slice_rows_kernel<S><<<grid, 128, 0, s->stream>>>(xt, ldx, (int)rows, s->n, nullptr, s->oz_xscale(), s->xs() + row0 * L.ld8,
                                                        L.ld8, s->Nl * L.ld8, (double)L.radix);
    LAUNCHED(s);
    LCX_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int lcx_slice_block(lcx_session* s, const double* xt, long long row0, long long rows, long long ldx) {
    S_REQUIRE_BOUND(s);
    LCX_REQUIRE(s->L.S > 0, "only meaningful in the split modes");
    LCX_REQUIRE(xt != nullptr && row0 >= 0 && rows > 0 && row0 + rows <= s->Nl && ldx >= s->n, "bad row block");
    switch (s->L.S) {
        case 3: return oz_slice_block_t<3>(s, xt, row0, rows, ldx);
        case 4: return oz_slice_block_t<4>(s, xt, row0, rows, ldx);
        case 5: return oz_slice_block_t<5>(s, xt, row0, rows, ldx);
        case 6: return oz_slice_block_t<6>(s, xt, row0, rows, ldx);
        case 7: return oz_slice_block_t<7>(s, xt, row0, rows, ldx);
    }
    return fail(LCX_ERR_STATE, "lcx_slice_block", "bad digit count");
}

extern "C" int lcx_array_info(lcx_session* s, int array_id, int set, long long* offset, long long* rows, long long* cols,
                              long long* ld) {
    S_REQUIRE_BOUND(s);
    LCX_REQUIRE(array_id >= 0 && array_id < LCX_A_COUNT && (set == 0 || set == 1), "bad array id / set");
    const int phys = (array_id <= LCX_A_UJ) ? (set ^ s->cur) : 0;
    const Slot& sl = s->L.slot[array_id][phys];
    if (offset) *offset = sl.off;
    if (rows) *rows = (array_id == LCX_A_D) ? s->m : sl.rows;
    if (cols) *cols = sl.cols;
    if (ld) *ld = sl.ld;
    return 0;
}

// Where the int8 digit planes and their power-of-two scales live (split modes; for inspection and bit-exact tests).
extern "C" int lcx_digit_planes_info(lcx_session* s, int which, long long* offset, int* digits, long long* rows,
                                     long long* cols, long long* ld_bytes, long long* scale_offset, int* radix) {
    S_REQUIRE_BOUND(s);
    const Layout& L = s->L;
    LCX_REQUIRE(L.S > 0, "no digit planes in the DMMA mode");
    LCX_REQUIRE(which >= 0 && which <= 2, "which: 0 = X~, 1 = A (last small operand), 2 = Y");
    const int slot = which == 0 ? I_XS : (which == 1 ? I_AS : I_YS);
    if (offset) *offset = L.slot[slot][0].off;
    if (digits) *digits = L.S;
    if (rows) *rows = which == 0 ? s->Nl : s->m;          // the Y planes are stored transposed: [factor][sample]
    if (cols) *cols = which == 2 ? s->Nl : s->n;
    if (ld_bytes) *ld_bytes = which == 2 ? L.ldk8 : L.ld8;
    if (scale_offset) *scale_offset = L.slot[I_OZV][0].off + (which == 0 ? 0 : (which == 1 ? 16 : 16 + 2 * L.ldm));
    if (radix) *radix = L.radix;
    return 0;
}

// ---- preprocessing -----------------------------------------------------------------------------
static const int kSlabRows = 4096;

extern "C" long long lcx_colstats_scratch_doubles(long long n_rows, int n_vars) {
    return 2LL * cdiv(n_rows, kSlabRows) * round_up(n_vars, 16) + 32;
}

// 16-byte row loads are possible when the base pointer and the leading dimension are multiples of 4 elements
template <typename T>
static bool vec4_ok(const T* x, long long ldx) { return ((uintptr_t)x % 16 == 0) && (ldx % 4 == 0); }

template <typename T>
static int colstats_sum_t(lcx_session* s, const T* x, long long N, int n, long long ldx, int has_marker, double marker,
                          double* sum, double* cnt, double* scratch) {
    const int slabs = cdiv(N, kSlabRows);
    const long long ldp = round_up(n, 16);
    double* ps = scratch;
    double* pc = scratch + (long long)slabs * ldp;
    const dim3 block(32, 8);
    if (vec4_ok(x, ldx))
        colstats_sum_kernel<T, 4><<<dim3(cdiv(n, 128), slabs), block, 0, s->stream>>>(x, N, n, ldx, kSlabRows, has_marker, marker,
                                                                                    marker != marker, ps, pc, ldp);
    else
        colstats_sum_kernel<T, 1><<<dim3(cdiv(n, 32), slabs), block, 0, s->stream>>>(x, N, n, ldx, kSlabRows, has_marker, marker,
                                                                                   marker != marker, ps, pc, ldp);
    LAUNCHED(s);
    combine_slabs_kernel<<<cdiv(n, 256), 256, 0, s->stream>>>(ps, slabs, ldp, sum, n);
    LAUNCHED(s);
    combine_slabs_kernel<<<cdiv(n, 256), 256, 0, s->stream>>>(pc, slabs, ldp, cnt, n);
    LAUNCHED(s);
    LCX_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int lcx_colstats_sum(lcx_session* s, const void* x, int dtype, long long n_rows, int n_vars, long long ldx,
                                int has_marker, double marker, double* sum, double* cnt, double* scratch,
                                long long scratch_doubles) {
    LCX_REQUIRE(s && x && sum && cnt && scratch, "null argument");
    LCX_REQUIRE(n_rows > 0 && n_vars > 0 && ldx >= n_vars, "bad shape");
    LCX_REQUIRE(scratch_doubles >= lcx_colstats_scratch_doubles(n_rows, n_vars), "scratch too small");
    LCX_CUDA(cudaSetDevice(s->device));
    if (dtype == LCX_F32) return colstats_sum_t<float>(s, (const float*)x, n_rows, n_vars, ldx, has_marker, marker, sum, cnt, scratch);
    if (dtype == LCX_F64) return colstats_sum_t<double>(s, (const double*)x, n_rows, n_vars, ldx, has_marker, marker, sum, cnt, scratch);
    return fail(LCX_ERR_ARG, "lcx_colstats_sum", "unknown dtype");
}

extern "C" int lcx_colstats_mean(lcx_session* s, const double* sum, const double* cnt, double* mean, int n_vars) {
    LCX_REQUIRE(s && sum && cnt && mean && n_vars > 0, "bad argument");
    LCX_CUDA(cudaSetDevice(s->device));
    finish_mean_kernel<<<cdiv(n_vars, 256), 256, 0, s->stream>>>(sum, cnt, mean, n_vars);
    LAUNCHED(s);
    LCX_CUDA(cudaGetLastError());
    return 0;
}

template <typename T>
static int colstats_sqdev_t(lcx_session* s, const T* x, long long N, int n, long long ldx, int has_marker, double marker,
                            const double* mean, double* sq, double* maxdev, double* scratch) {
    const int slabs = cdiv(N, kSlabRows);
    const long long ldp = round_up(n, 16);
    double* pmax = maxdev ? scratch + (long long)slabs * ldp : nullptr;
    const dim3 block(32, 8);
    if (vec4_ok(x, ldx))
        colstats_sqdev_kernel<T, 4><<<dim3(cdiv(n, 128), slabs), block, 0, s->stream>>>(x, N, n, ldx, kSlabRows, has_marker, marker,
                                                                                      marker != marker, mean, scratch, pmax, ldp);
    else
        colstats_sqdev_kernel<T, 1><<<dim3(cdiv(n, 32), slabs), block, 0, s->stream>>>(x, N, n, ldx, kSlabRows, has_marker, marker,
                                                                                     marker != marker, mean, scratch, pmax, ldp);
    LAUNCHED(s);
    combine_slabs_kernel<<<cdiv(n, 256), 256, 0, s->stream>>>(scratch, slabs, ldp, sq, n);
    LAUNCHED(s);
    if (maxdev) {
        combine_slabs_max_kernel<<<cdiv(n, 256), 256, 0, s->stream>>>(pmax, slabs, ldp, maxdev, n);
        LAUNCHED(s);
    }
    LCX_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int lcx_colstats_sqdev(lcx_session* s, const void* x, int dtype, long long n_rows, int n_vars, long long ldx,
                                  int has_marker, double marker, const double* mean, double* sq, double* maxdev,
                                  double* scratch, long long scratch_doubles) {
    LCX_REQUIRE(s && x && mean && sq && scratch, "null argument");
    LCX_REQUIRE(n_rows > 0 && n_vars > 0 && ldx >= n_vars, "bad shape");
    LCX_REQUIRE(scratch_doubles >= lcx_colstats_scratch_doubles(n_rows, n_vars), "scratch too small");
    LCX_CUDA(cudaSetDevice(s->device));
    if (dtype == LCX_F32) return colstats_sqdev_t<float>(s, (const float*)x, n_rows, n_vars, ldx, has_marker, marker, mean, sq, maxdev, scratch);
    if (dtype == LCX_F64) return colstats_sqdev_t<double>(s, (const double*)x, n_rows, n_vars, ldx, has_marker, marker, mean, sq, maxdev, scratch);
    return fail(LCX_ERR_ARG, "lcx_colstats_sqdev", "unknown dtype");
}

extern "C" int lcx_colstats_std(lcx_session* s, const double* sq, const double* cnt, double n_rows_total, int use_nobs,
                                double* sd, int n_vars) {
    LCX_REQUIRE(s && sq && cnt && sd && n_vars > 0, "bad argument");
    LCX_CUDA(cudaSetDevice(s->device));
    finish_std_kernel<<<cdiv(n_vars, 256), 256, 0, s->stream>>>(sq, cnt, n_rows_total, use_nobs, sd, n_vars);
    LAUNCHED(s);
    LCX_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int lcx_standardize(lcx_session* s, const void* x, int dtype, long long n_rows, int n_vars, long long ldx,
                               int has_marker, double marker, int gauss_mode, const double* impute, const double* mean,
                               const double* sd, double* out, long long ldo) {
    LCX_REQUIRE(s && x && out, "null argument");
    LCX_REQUIRE(gauss_mode == LCX_GAUSS_NONE || (mean && sd), "mean/std required");
    LCX_REQUIRE(has_marker == 0 || impute != nullptr, "imputation means required when a missing marker is set");
    LCX_REQUIRE(n_rows > 0 && n_vars > 0 && ldx >= n_vars && ldo >= n_vars, "bad shape");
    LCX_CUDA(cudaSetDevice(s->device));
    const int nan_marker = marker != marker;
    const bool v4 = (ldo % 4 == 0) && ((uintptr_t)out % 16 == 0) &&
                    (dtype == LCX_F32 ? vec4_ok((const float*)x, ldx) : vec4_ok((const double*)x, ldx));
    const dim3 grid((unsigned)n_rows, cdiv(ldo, v4 ? 1024 : 256));
    if (dtype == LCX_F32 && v4)
        standardize_kernel<float, 4><<<grid, 256, 0, s->stream>>>((const float*)x, n_rows, n_vars, ldx, has_marker, marker,
                                                                nan_marker, gauss_mode, impute, mean, sd, out, ldo);
    else if (dtype == LCX_F32)
        standardize_kernel<float, 1><<<grid, 256, 0, s->stream>>>((const float*)x, n_rows, n_vars, ldx, has_marker, marker,
                                                                nan_marker, gauss_mode, impute, mean, sd, out, ldo);
    else if (dtype == LCX_F64 && v4)
        standardize_kernel<double, 4><<<grid, 256, 0, s->stream>>>((const double*)x, n_rows, n_vars, ldx, has_marker, marker,
                                                                 nan_marker, gauss_mode, impute, mean, sd, out, ldo);
    else if (dtype == LCX_F64)
        standardize_kernel<double, 1><<<grid, 256, 0, s->stream>>>((const double*)x, n_rows, n_vars, ldx, has_marker, marker,
                                                                 nan_marker, gauss_mode, impute, mean, sd, out, ldo);
    else
        return fail(LCX_ERR_ARG, "lcx_standardize", "unknown dtype");
    LAUNCHED(s);
    LCX_CUDA(cudaGetLastError());
    return 0;
}

// ---- GEMM plumbing -----------------------------------------------------------------------------
// Sum over ranks of [body = sum_z part[z] (rows x ld, valid cols) | tail (ntail doubles)] -> dst_body / dst_tail.
// Peer path: ONE kernel (split-K combine + two-shot all-reduce over NVLink peer memory).  Otherwise the fixed-order
// local combine followed by the installed hook (NCCL through torch.distributed), or nothing on a single rank.
static int combine_and_allreduce(lcx_session* s, const double* part, int splits, long long stride, int rows, int cols,
                                 long long ld, double* dst_body, double* tail, int ntail) {
    const long long body = (long long)rows * ld;
    if (s->peers.world > 1) {
        LCX_REQUIRE(body + ntail <= s->peers.count, "peer buffer too small");
        const int slot = (int)(s->ar_calls & 1);
        unsigned long long epoch0 = 2ULL * s->ar_calls;
        s->ar_calls++;
        far::Peers pp = s->peers;
        const double* tl = tail;
        void* args[] = {&pp, &part, &splits, &stride, &rows, &cols, &ld, &tl, &ntail, (void*)&slot, &epoch0};
        LCX_CUDA(cudaLaunchCooperativeKernel((void*)far::reduce_allreduce_kernel, dim3(kSMs), dim3(512), args, 0, s->stream));
        LAUNCHED(s);
        const double* out = s->peers.base[s->peers.rank] + 2 * s->peers.count;
        if (rows > 0 && ntail > 0 && tail == dst_body + body) {  // D and the sums of squares are adjacent: one copy
            LCX_CUDA(cudaMemcpyAsync(dst_body, out, (size_t)(body + ntail) * sizeof(double), cudaMemcpyDeviceToDevice, s->stream));
            return 0;
        }
        if (rows > 0)
            LCX_CUDA(cudaMemcpyAsync(dst_body, out, (size_t)body * sizeof(double), cudaMemcpyDeviceToDevice, s->stream));
        if (ntail > 0)
            LCX_CUDA(cudaMemcpyAsync(tail, out + body, (size_t)ntail * sizeof(double), cudaMemcpyDeviceToDevice, s->stream));
        return 0;
    }
    if (rows > 0 && (splits > 1 || part != dst_body)) {
        LCX_TRY(launch_reduce_splits(part, splits, stride, dst_body, rows, cols, ld, s->stream));
        LAUNCHED(s);
    }
    if (s->hook) {
        // body and tail are contiguous in the workspace (D is followed by the column sums of squares)
        const long long off = (rows > 0 ? dst_body : tail) - s->ws;
        if (s->hook(s->hook_user, off, (rows > 0 ? body : 0) + ntail) != 0)
            return fail(LCX_ERR_STATE, "allreduce hook", "hook reported failure");
    }
    return 0;
}

static int run_gemm(lcx_session* s, GemmLayout lay, const GemmPlan& pl, GemmArgs a, double* part, long long out_count,
                    bool leave_partials = false) {
    // out_count = number of doubles of one full output (rows * ldc) -- the split stride
    if (pl.splits > 1 && leave_partials) {
        LCX_REQUIRE(a.Cadd == nullptr, "split-K with Cadd is not supported");
        a.C = part;
        a.c_split_stride = out_count;
        LCX_TRY(launch_gemm(lay, pl, a, s->stream));
        LAUNCHED(s);
    } else if (pl.splits > 1) {
        double* final_c = a.C;
        const double* cadd = a.Cadd;
        LCX_REQUIRE(cadd == nullptr, "split-K with Cadd is not supported");
        a.C = part;
        a.c_split_stride = out_count;
        LCX_TRY(launch_gemm(lay, pl, a, s->stream));
        LAUNCHED(s);
        const int out_rows = a.trans_out ? a.N : a.M, out_cols = a.trans_out ? a.M : a.N;
        LCX_TRY(launch_reduce_splits(part, pl.splits, out_count, final_c, out_rows, out_cols, a.ldc, s->stream));
        LAUNCHED(s);
    } else {
        a.c_split_stride = 0;
        LCX_TRY(launch_gemm(lay, pl, a, s->stream));
        LAUNCHED(s);
    }
    return 0;
}

// m x m product over the variables (ry, H): split-K GEMM whose fixed-order combine also performs np.fill_diagonal
// (raw diagonal -> diag_out if given, diag_value stored) -- one launch less than combine + diag kernel.
static int run_square_gemm(lcx_session* s, GemmArgs a, double diag_value, double* diag_out) {
    const Layout& L = s->L;
    const int m = s->m;
    const long long out_count = (long long)m * L.ldm;
    if (L.plan_mm.splits > 1) {
        double* final_c = a.C;
        a.C = s->ptr(I_PART);
        a.c_split_stride = out_count;
        LCX_TRY(launch_gemm(kLayoutKK, L.plan_mm, a, s->stream));
        LAUNCHED(s);
        LCX_TRY(launch_reduce_splits(s->ptr(I_PART), L.plan_mm.splits, out_count, final_c, m, m, L.ldm, s->stream,
                                     diag_out ? diag_out : s->ptr(I_F), diag_value));
        LAUNCHED(s);
    } else {
        double* final_c = a.C;
        LCX_TRY(run_gemm(s, kLayoutKK, L.plan_mm, a, s->ptr(I_PART), out_count));
        diag_fix_kernel<<<cdiv(m, 128), 128, 0, s->stream>>>(final_c, L.ldm, m, diag_value, diag_out);
        LAUNCHED(s);
    }
    return 0;
}

// Y = X~ A^T (+ colsq into D's tail), D = X~^T Y summed over splits, then the rank all-reduce.
static int xpair(lcx_session* s, const double* A, bool want_colsq) {
    const Layout& L = s->L;
    const int m = s->m, n = s->n;
    double* Y = s->ptr(LCX_A_Y);
    double* D = s->ptr(LCX_A_D);
    double* svec = D + (long long)m * L.ld;
    cudaEvent_t* ev = nullptr;
    if (s->prof_on && s->prof_pending < s->prof_cap) ev = s->prof_ev + 4 * s->prof_pending;
    if (ev) LCX_CUDA(cudaEventRecord(ev[0], s->stream));
    if (L.S > 0) {
        LCX_TRY(oz_pair(s, A, svec, ev, false, want_colsq));
        if (ev) {
            LCX_CUDA(cudaEventRecord(ev[2], s->stream));
            s->prof_pending++;
        }
    } else {
    {   // K1
        GemmArgs a;
        memset(&a, 0, sizeof(a));
        a.A = s->xt; a.B = A; a.C = Y;
        a.M = (int)s->Nl; a.N = m; a.K = n;
        a.lda = s->ldx; a.ldb = L.ld; a.ldc = L.ldy;
        const bool k1_split = L.plan_k1.splits > 1;
        a.colsq_part = (want_colsq && !k1_split) ? s->ptr(I_COLSQ) : nullptr;
        a.ld_colsq = (int)L.ldy;
        LCX_TRY(run_gemm(s, kLayoutKK, L.plan_k1, a, k1_split ? s->ptr(I_PART) : nullptr, s->Nl * L.ldy));
        if (ev) LCX_CUDA(cudaEventRecord(ev[1], s->stream));
        if (want_colsq && k1_split) {  // split over variables: the sums of squares come from the reduced Y
            double* ystat = s->ws + L.slot[I_YSTAT][0].off;
            oz::y_stats_kernel<<<dim3(cdiv(m, 32), L.ystat_slabs), dim3(32, 8), 0, s->stream>>>(Y, L.ldy, s->Nl, m, kYStatRows, ystat,
                                                                                          L.ldm);
            LAUNCHED(s);
            oz::y_stats_finish_kernel<<<m, 256, 0, s->stream>>>(ystat, L.ystat_slabs, L.ldm, m, s->oz_xscale(), svec, s->oz_yscale(),
                                                               s->oz_dscale());
            LAUNCHED(s);
        } else if (want_colsq) {
            reduce_colsq_kernel<<<cdiv(m, 128), 128, 0, s->stream>>>(s->ptr(I_COLSQ), L.plan_k1.grid.x, (int)L.ldy, svec, m);
            LAUNCHED(s);
        }
    }
    {   // K2: (X~^T Y)^T written factor-major
        GemmArgs a;
        memset(&a, 0, sizeof(a));
        a.A = s->xt; a.B = Y; a.C = D;
        a.M = n; a.N = m; a.K = (int)s->Nl;
        a.lda = s->ldx; a.ldb = L.ldy; a.ldc = L.ld;
        a.trans_out = 1;
        if (ev) LCX_CUDA(cudaEventRecord(ev[3], s->stream));  // K2 starts after the (tiny) colsq reduction
        LCX_TRY(run_gemm(s, kLayoutMN, L.plan_k2, a, s->ptr(I_PART), (long long)m * L.ld, true));
        if (ev) {
            LCX_CUDA(cudaEventRecord(ev[2], s->stream));
            s->prof_pending++;
        }
        const bool split = L.plan_k2.splits > 1;
        LCX_TRY(combine_and_allreduce(s, split ? s->ptr(I_PART) : D, split ? L.plan_k2.splits : 1, (long long)m * L.ld, m, n, L.ld,
                                      D, svec, want_colsq ? m : 0));
    }
    }
    LCX_CUDA(cudaGetLastError());
    return 0;
}

static int read_mailbox(lcx_session* s) {
    LCX_CUDA(cudaMemcpyAsync(s->mailbox, s->ptr(LCX_A_SCALARS), 16 * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    LCX_CUDA(cudaStreamSynchronize(s->stream));
    return 0;
}

// ry, Qij, Qi-Si^2, TC, uj for `set`, given rho/invrho/rinv/Si (and W) of that set.
static int moments_tail(lcx_session* s, int set, double c1, double e2, int uj_mode) {
    const Layout& L = s->L;
    const int m = s->m, n = s->n;
    double* W = s->ptr(LCX_A_W, set);
    double* rho = s->ptr(LCX_A_RHO, set);
    double* rinv = s->ptr(LCX_A_RHOINVRHO, set);
    double* ry = s->ptr(LCX_A_RY, set);
    double* Qij = s->ptr(LCX_A_QIJ, set);
    if (L.mm_i8) {  // both products as exact int8 digit-plane products on tcgen05
        LCX_TRY(oz_square(s, W, rho, ry, 1.0, s->ptr(I_UJDIAG)));
        LCX_TRY(oz_mn(s, ry, rinv, Qij, rinv, true));  // Qij = rinv + (ry - I) rinv
    } else {
    {   // ry = W rho^T  (:261), diag -> 1 (:263); the diagonal before the fill is uj by linearity
        GemmArgs a;
        memset(&a, 0, sizeof(a));
        a.A = W; a.B = rho; a.C = ry;
        a.M = m; a.N = m; a.K = n;
        a.lda = L.ld; a.ldb = L.ld; a.ldc = L.ldm;
        LCX_TRY(run_square_gemm(s, a, 1.0, s->ptr(I_UJDIAG)));
    }
    {   // Qij = ry rinv  (:266)
        GemmArgs a;
        memset(&a, 0, sizeof(a));
        a.A = ry; a.B = rinv; a.C = Qij;
        a.M = m; a.N = n; a.K = m;
        a.lda = L.ldm; a.ldb = L.ld; a.ldc = L.ld;
        LCX_TRY(run_gemm(s, kLayoutKN, L.plan_mn, a, nullptr, 0));
    }
    }
    moments_stage2_kernel<<<L.nstrips, dim3(kStripCols, kStripRows), 0, s->stream>>>(
        rho, rinv, Qij, s->ptr(LCX_A_SI, set), s->ptr(LCX_A_QISI2, set), s->ptr(I_SPART), m, n, L.ld);
    LAUNCHED(s);
    moments_finish_kernel<<<1, 256, 0, s->stream>>>(s->ptr(I_SPART), L.nstrips, uj_mode,
                                                   s->ptr(LCX_A_D) + (long long)m * L.ld, s->ptr(I_W2), s->ptr(I_UJDIAG), c1,
                                                   e2, s->ptr(LCX_A_UJ, set), m, s->ptr(LCX_A_SCALARS));
    LAUNCHED(s);
    LCX_CUDA(cudaGetLastError());
    return 0;
}

// full from-X moment evaluation of W(set) into `set`
static int moments_from_x(lcx_session* s, int set, double eps) {
    const Layout& L = s->L;
    const int m = s->m, n = s->n;
    const double c1 = (1.0 - eps * eps) / (double)s->Nt, e2 = eps * eps;
    double* W = s->ptr(LCX_A_W, set);
    LCX_TRY(xpair(s, W, true));
    row_dot_kernel<<<m, 256, 0, s->stream>>>(W, W, s->ptr(I_W2), n, L.ld);
    LAUNCHED(s);
    moments_stage1_kernel<true><<<L.nstrips, dim3(kStripCols, kStripRows), 0, s->stream>>>(
        s->ptr(LCX_A_D), W, nullptr, nullptr, nullptr, 0.0, c1, e2, nullptr, s->ptr(LCX_A_RHO, set),
        s->ptr(LCX_A_INVRHO, set), s->ptr(LCX_A_RHOINVRHO, set), s->ptr(LCX_A_SI, set), m, n, L.ld);
    LAUNCHED(s);
    return moments_tail(s, set, c1, e2, 0);
}

// ---- exported steps ----------------------------------------------------------------------------
extern "C" long long lcx_project_scratch_doubles(long long n_rows, int n_factors) {
    return (long long)cdiv(n_rows, 128) * round_up(n_factors, 8) + 16;
}

extern "C" int lcx_project(lcx_session* s, const double* xt, long long n_rows, int n_vars, long long ldx, const double* a,
                           long long lda, int n_factors, double* y, long long ldy, double* colsq, double* scratch,
                           long long scratch_doubles) {
    LCX_REQUIRE(s && xt && a && y, "null argument");
    LCX_REQUIRE(n_rows > 0 && n_rows < (1LL << 31) && n_vars > 0 && n_factors > 0, "bad shape");
    LCX_REQUIRE(colsq == nullptr || (scratch != nullptr && scratch_doubles >= lcx_project_scratch_doubles(n_rows, n_factors)),
                "scratch too small for the column sums");
    LCX_CUDA(cudaSetDevice(s->device));
    GemmPlan pl = plan_gemm((int)n_rows, n_factors, n_vars, kSMs, 1, false);
    GemmArgs g;
    memset(&g, 0, sizeof(g));
    g.A = xt; g.B = a; g.C = y;
    g.M = (int)n_rows; g.N = n_factors; g.K = n_vars;
    g.lda = ldx; g.ldb = lda; g.ldc = ldy;
    const int ldp = (int)round_up(n_factors, 8);
    g.colsq_part = colsq ? scratch : nullptr;
    g.ld_colsq = ldp;
    LCX_TRY(launch_gemm(kLayoutKK, pl, g, s->stream));
    LAUNCHED(s);
    if (colsq) {
        reduce_colsq_kernel<<<cdiv(n_factors, 128), 128, 0, s->stream>>>(scratch, pl.grid.x, ldp, colsq, n_factors);
        LAUNCHED(s);
    }
    LCX_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int lcx_sig(lcx_session* s, const double* u, double eps, double* out) {
    S_REQUIRE_BOUND(s);
    LCX_REQUIRE(u && out, "null argument");
    LCX_TRY(xpair(s, u, false));
    const double c1 = (1.0 - eps * eps) / (double)s->Nt, e2 = eps * eps;
    sig_finish_kernel<<<grid_mn(s->m, s->n), 256, 0, s->stream>>>(s->ptr(LCX_A_D), u, c1, e2, out, s->m, s->n, s->L.ld);
    LAUNCHED(s);
    LCX_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int lcx_set_w(lcx_session* s, const double* host_w, long long host_ld) {
    S_REQUIRE_BOUND(s);
    LCX_REQUIRE(host_w && host_ld >= s->n, "bad host array");
    LCX_CUDA(cudaMemcpy2DAsync(s->ptr(LCX_A_W), s->L.ld * sizeof(double), host_w, host_ld * sizeof(double),
                               (size_t)s->n * sizeof(double), s->m, cudaMemcpyHostToDevice, s->stream));
    LCX_CUDA(cudaStreamSynchronize(s->stream));
    return 0;
}

extern "C" int lcx_get_w(lcx_session* s, double* host_w, long long host_ld) {
    S_REQUIRE_BOUND(s);
    LCX_REQUIRE(host_w && host_ld >= s->n, "bad host array");
    LCX_CUDA(cudaMemcpy2DAsync(host_w, host_ld * sizeof(double), s->ptr(LCX_A_W), s->L.ld * sizeof(double),
                               (size_t)s->n * sizeof(double), s->m, cudaMemcpyDeviceToHost, s->stream));
    LCX_CUDA(cudaStreamSynchronize(s->stream));
    return 0;
}

extern "C" int lcx_init_scale(lcx_session* s, double eps) {
    S_REQUIRE_BOUND(s);
    const Layout& L = s->L;
    const int m = s->m, n = s->n;
    double* W = s->ptr(LCX_A_W);
    double* svec = s->ptr(LCX_A_D) + (long long)m * L.ld;
    if (L.S > 0) {
        LCX_TRY(oz_pair(s, W, svec, nullptr, true, true));
    } else {
        LCX_TRY(lcx_project(s, s->xt, s->Nl, n, s->ldx, W, L.ld, m, s->ptr(LCX_A_Y), L.ldy, svec, s->ptr(I_COLSQ),
                            lcx_project_scratch_doubles(s->Nl, m)));
    }
    LCX_TRY(combine_and_allreduce(s, nullptr, 1, 0, 0, 0, L.ld, nullptr, svec, m));  // sum of Y^2 over ranks
    row_dot_kernel<<<m, 256, 0, s->stream>>>(W, W, s->ptr(I_W2), n, L.ld);
    LAUNCHED(s);
    const double c1 = (1.0 - eps * eps) / (double)s->Nt, e2 = eps * eps;
    init_scale_kernel<<<cdiv(m, 128), 128, 0, s->stream>>>(svec, s->ptr(I_W2), c1, e2, s->ptr(I_F), m);
    LAUNCHED(s);
    scale_rows_kernel<<<grid_mn(m, n), 256, 0, s->stream>>>(W, s->ptr(I_F), m, n, L.ld);
    LAUNCHED(s);
    LCX_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int lcx_stage_rescale(lcx_session* s, double eps, double eps_prev) {
    S_REQUIRE_BOUND(s);
    const Layout& L = s->L;
    const int m = s->m, n = s->n;
    double* W = s->ptr(LCX_A_W);
    row_dot_kernel<<<m, 256, 0, s->stream>>>(W, W, s->ptr(I_W2), n, L.ld);
    LAUNCHED(s);
    stage_scale_kernel<<<cdiv(m, 128), 128, 0, s->stream>>>(s->ptr(I_W2), s->ptr(LCX_A_UJ), eps, eps_prev, s->ptr(I_F), m);
    LAUNCHED(s);
    scale_rows_kernel<<<grid_mn(m, n), 256, 0, s->stream>>>(W, s->ptr(I_F), m, n, L.ld);
    LAUNCHED(s);
    LCX_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int lcx_permute_rows(lcx_session* s, const int* host_order) {
    S_REQUIRE_BOUND(s);
    LCX_REQUIRE(host_order != nullptr, "null order");
    const Layout& L = s->L;
    double* W = s->ptr(LCX_A_W);
    double* tmp = s->ptr(LCX_A_GRAD);
    for (int j = 0; j < s->m; ++j) {
        LCX_REQUIRE(host_order[j] >= 0 && host_order[j] < s->m, "order out of range");
        LCX_CUDA(cudaMemcpyAsync(tmp + (long long)j * L.ld, W + (long long)host_order[j] * L.ld, L.ld * sizeof(double),
                                 cudaMemcpyDeviceToDevice, s->stream));
    }
    LCX_CUDA(cudaMemcpyAsync(W, tmp, (size_t)s->m * L.ld * sizeof(double), cudaMemcpyDeviceToDevice, s->stream));
    return 0;
}

extern "C" int lcx_moments_ns(lcx_session* s, double eps, int check_uj, double* tc, double* max_uj) {
    S_REQUIRE_BOUND(s);
    LCX_TRY(moments_from_x(s, 0, eps));
    LCX_TRY(read_mailbox(s));
    if (tc) *tc = s->mailbox[0];
    if (max_uj) *max_uj = s->mailbox[1];
    return (check_uj && s->mailbox[1] >= 1.0) ? LCX_QUICK_FAIL : LCX_OK;
}

static int enqueue_direction(lcx_session* s, double eps) {
    const Layout& L = s->L;
    const int m = s->m, n = s->n;
    const double c1 = (1.0 - eps * eps) / (double)s->Nt, e2 = eps * eps;
    double* W = s->ptr(LCX_A_W);
    double* rho = s->ptr(LCX_A_RHO);
    double* rinv = s->ptr(LCX_A_RHOINVRHO);
    double* G = s->ptr(LCX_A_GRAD);
    double* T = s->ptr(I_T);
    double* H = s->ptr(I_RYINV);  // m x ldm scratch (the inverse buffer is idle outside the details path)
    direction_stage1_kernel<<<grid_mn(m, n), 256, 0, s->stream>>>(W, rho, s->ptr(LCX_A_INVRHO), rinv, s->ptr(LCX_A_QIJ),
                                                                s->ptr(LCX_A_SI), s->ptr(LCX_A_QISI2), s->ptr(LCX_A_UJ), T, G,
                                                                m, n, L.ld);
    LAUNCHED(s);
    if (L.mm_i8) {
        LCX_TRY(oz_square(s, T, rinv, H, 0.0, nullptr));  // H = T rinv^T, diag -> 0 (:294-295)
        LCX_TRY(oz_mn(s, H, W, G, G, false));             // grad = G0 + H W (:300)
    } else {
    {   // H = T rinv^T, diag -> 0 (:294-295)
        GemmArgs a;
        memset(&a, 0, sizeof(a));
        a.A = T; a.B = rinv; a.C = H;
        a.M = m; a.N = m; a.K = n;
        a.lda = L.ld; a.ldb = L.ld; a.ldc = L.ldm;
        LCX_TRY(run_square_gemm(s, a, 0.0, nullptr));
    }
    {   // grad = G0 + H W (:300), in place over G0
        GemmArgs a;
        memset(&a, 0, sizeof(a));
        a.A = H; a.B = W; a.C = G; a.Cadd = G;
        a.M = m; a.N = n; a.K = m;
        a.lda = L.ldm; a.ldb = L.ld; a.ldc = L.ld;
        LCX_TRY(run_gemm(s, kLayoutKN, L.plan_mn, a, nullptr, 0));
    }
    }
    LCX_TRY(xpair(s, G, false));  // X~^T (X~ grad^T): the one pass over X of this iteration (:301)
    row_dot_kernel<<<m, 256, 0, s->stream>>>(rho, G, s->ptr(I_BJ), n, L.ld);  // Bj (:302)
    LAUNCHED(s);
    const dim3 g2(cdiv(n, 256), m);
    direction_stage2_kernel<<<g2, 256, 0, s->stream>>>(W, rho, G, s->ptr(LCX_A_D), s->ptr(LCX_A_UJ), s->ptr(I_BJ), c1, e2,
                                                     s->ptr(LCX_A_UPDATE), s->ptr(LCX_A_RDIR), s->ptr(I_SPART), m, n, L.ld);
    LAUNCHED(s);
    sum_partials_kernel<<<1, 256, 0, s->stream>>>(s->ptr(I_SPART), (int)(g2.x * g2.y), s->ptr(LCX_A_SCALARS) + 2);
    LAUNCHED(s);
    LCX_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int lcx_direction_ns(lcx_session* s, double eps, double* tangent) {
    S_REQUIRE_BOUND(s);
    LCX_TRY(enqueue_direction(s, eps));
    LCX_TRY(read_mailbox(s));
    if (tangent) *tangent = s->mailbox[2];
    return 0;
}

static int enqueue_trial(lcx_session* s, double eps, double eta, int exact) {
    const Layout& L = s->L;
    const int m = s->m, n = s->n;
    const double c1 = (1.0 - eps * eps) / (double)s->Nt, e2 = eps * eps;
    if (exact) {
        axpy_kernel<<<grid_mn(m, n), 256, 0, s->stream>>>(s->ptr(LCX_A_W), s->ptr(LCX_A_UPDATE), eta, s->ptr(LCX_A_W, 1), m, n,
                                                        L.ld);
        LAUNCHED(s);
        LCX_TRY(moments_from_x(s, 1, eps));
    } else {
        moments_stage1_kernel<false><<<L.nstrips, dim3(kStripCols, kStripRows), 0, s->stream>>>(
            nullptr, s->ptr(LCX_A_W), s->ptr(LCX_A_UPDATE), s->ptr(LCX_A_RHO), s->ptr(LCX_A_RDIR), eta, c1, e2,
            s->ptr(LCX_A_W, 1), s->ptr(LCX_A_RHO, 1), s->ptr(LCX_A_INVRHO, 1), s->ptr(LCX_A_RHOINVRHO, 1),
            s->ptr(LCX_A_SI, 1), m, n, L.ld);
        LAUNCHED(s);
        LCX_TRY(moments_tail(s, 1, c1, e2, 1));
    }
    return 0;
}

extern "C" int lcx_trial_ns(lcx_session* s, double eps, double eta, int exact, double* tc, double* max_uj) {
    S_REQUIRE_BOUND(s);
    LCX_TRY(enqueue_trial(s, eps, eta, exact));
    LCX_TRY(read_mailbox(s));
    if (tc) *tc = s->mailbox[0];
    if (max_uj) *max_uj = s->mailbox[1];
    return (s->mailbox[1] >= 1.0) ? LCX_QUICK_FAIL : LCX_OK;
}

// Direction (:292-305) and the first backtracking trial at `eta` (:320-321) enqueued back to back, ONE host
// synchronisation for update_tangent, TC and max uj.  The trial is speculative: if update_tangent >= 0 the caller
// discards it (the reference returns before trying, :306-311).  Linear trials only (an exact trial costs a pass pair).
extern "C" int lcx_direction_trial_ns(lcx_session* s, double eps, double eta, double* tangent, double* tc, double* max_uj) {
    S_REQUIRE_BOUND(s);
    LCX_TRY(enqueue_direction(s, eps));
    LCX_TRY(enqueue_trial(s, eps, eta, 0));
    LCX_TRY(read_mailbox(s));
    if (tangent) *tangent = s->mailbox[2];
    if (tc) *tc = s->mailbox[0];
    if (max_uj) *max_uj = s->mailbox[1];
    return (s->mailbox[1] >= 1.0) ? LCX_QUICK_FAIL : LCX_OK;
}

extern "C" int lcx_accept_trial(lcx_session* s) {
    S_REQUIRE_BOUND(s);
    s->cur ^= 1;
    return 0;
}

static int run_inverse(lcx_session* s, const double* a, long long lda, int m, double* out, long long ldo, double* aug,
                       int* status) {
    gauss_jordan_inverse_kernel<<<1, 1024, 0, s->stream>>>(a, lda, m, aug, out, ldo, status);
    LAUNCHED(s);
    LCX_CUDA(cudaGetLastError());
    return 0;
}

// shared tail of the details computations: MI, X_i^2|Y, I(X_i;Y), TCs, TC_no_overlap, TC_direct, additivity
static int details_tail(lcx_session* s, const double* other, const double* yj2_in) {
    const Layout& L = s->L;
    const int m = s->m, n = s->n;
    details_cols_kernel<<<L.nstrips, dim3(kStripCols, kStripRows), 0, s->stream>>>(
        s->ptr(LCX_A_RHO), s->ptr(LCX_A_XZ), other, s->ptr(LCX_A_MI), s->ptr(LCX_A_X2Y), s->ptr(LCX_A_IXY), s->ptr(I_SPART),
        m, n, L.ld);
    LAUNCHED(s);
    row_dot_kernel<<<m, 256, 0, s->stream>>>(s->ptr(LCX_A_MI), nullptr, s->ptr(I_ROWMI), n, L.ld);
    LAUNCHED(s);
    details_finish_kernel<<<1, 256, 0, s->stream>>>(s->ptr(I_SPART), L.nstrips, s->ptr(LCX_A_UJ), yj2_in, s->ptr(I_ROWMI),
                                                   s->ptr(LCX_A_YJ2), s->ptr(LCX_A_IYX), s->ptr(LCX_A_TCS),
                                                   s->ptr(LCX_A_TCDIRECT), s->ptr(I_SQRTY), m, s->ptr(LCX_A_SCALARS) + 4);
    LAUNCHED(s);
    LCX_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int lcx_details_ns(lcx_session* s, double* tc_no_overlap, double* additivity) {
    S_REQUIRE_BOUND(s);
    const Layout& L = s->L;
    const int m = s->m, n = s->n;
    double* rho = s->ptr(LCX_A_RHO);
    // X_i Z_j = solve(ry, rho)^T (:280) as ry^-1 rho
    LCX_TRY(run_inverse(s, s->ptr(LCX_A_RY), L.ldm, m, s->ptr(I_RYINV), L.ldm, s->ptr(I_AUG), (int*)s->ptr(I_STATUS)));
    {
        GemmArgs a;
        memset(&a, 0, sizeof(a));
        a.A = s->ptr(I_RYINV); a.B = rho; a.C = s->ptr(LCX_A_XZ);
        a.M = m; a.N = n; a.K = m;
        a.lda = L.ldm; a.ldb = L.ld; a.ldc = L.ld;
        LCX_TRY(run_gemm(s, kLayoutKN, L.plan_mn, a, nullptr, 0));
    }
    LCX_TRY(details_tail(s, rho, nullptr));
    // X_i Y_j = rho^T sqrt(Y_j^2) (:279)
    scale_rows_out_kernel<<<grid_mn(m, n), 256, 0, s->stream>>>(rho, s->ptr(I_SQRTY), s->ptr(LCX_A_XY), m, n, L.ld);
    LAUNCHED(s);
    LCX_CUDA(cudaGetLastError());
    LCX_TRY(read_mailbox(s));
    if (tc_no_overlap) *tc_no_overlap = s->mailbox[4];
    if (additivity) *additivity = s->mailbox[6];
    return 0;
}

extern "C" int lcx_moments_syn(lcx_session* s, double* tc, double* additivity) {
    S_REQUIRE_BOUND(s);
    const Layout& L = s->L;
    const int m = s->m, n = s->n;
    double* W = s->ptr(LCX_A_W);
    double* XY = s->ptr(LCX_A_XY);
    double* cy = s->ptr(LCX_A_CY);
    double* rinv = s->ptr(LCX_A_RHOINVRHO);
    LCX_TRY(xpair(s, W, false));
    // X_i Y_j = X~^T Y / N (:354)
    sig_finish_kernel<<<grid_mn(m, n), 256, 0, s->stream>>>(s->ptr(LCX_A_D), s->ptr(LCX_A_D), 1.0 / (double)s->Nt, 0.0, XY, m, n,
                                                          L.ld);
    LAUNCHED(s);
    {   // cy = W XY + yscale^2 I (:355)
        GemmArgs a;
        memset(&a, 0, sizeof(a));
        a.A = W; a.B = XY; a.C = cy;
        a.M = m; a.N = m; a.K = n;
        a.lda = L.ld; a.ldb = L.ld; a.ldc = L.ldm;
        LCX_TRY(run_gemm(s, kLayoutKK, L.plan_mm, a, s->ptr(I_PART), (long long)m * L.ldm));
        diag_add_kernel<<<cdiv(m, 128), 128, 0, s->stream>>>(cy, L.ldm, m, 1.0);
        LAUNCHED(s);
    }
    syn_ry_kernel<<<dim3(cdiv(m, 128), m), 128, 0, s->stream>>>(cy, L.ldm, m, s->ptr(LCX_A_RY), s->ptr(LCX_A_YJ2),
                                                              s->ptr(I_SQRTY));
    LAUNCHED(s);
    syn_stage1_kernel<<<L.nstrips, dim3(kStripCols, kStripRows), 0, s->stream>>>(
        XY, s->ptr(I_SQRTY), s->ptr(LCX_A_RHO), s->ptr(LCX_A_INVRHO), rinv, s->ptr(LCX_A_SI), m, n, L.ld);
    LAUNCHED(s);
    {   // Qij = ry rinv (:361)
        GemmArgs a;
        memset(&a, 0, sizeof(a));
        a.A = s->ptr(LCX_A_RY); a.B = rinv; a.C = s->ptr(LCX_A_QIJ);
        a.M = m; a.N = n; a.K = m;
        a.lda = L.ldm; a.ldb = L.ld; a.ldc = L.ld;
        LCX_TRY(run_gemm(s, kLayoutKN, L.plan_mn, a, nullptr, 0));
    }
    syn_qi_kernel<<<L.nstrips, dim3(kStripCols, kStripRows), 0, s->stream>>>(rinv, s->ptr(LCX_A_QIJ), s->ptr(LCX_A_QISI2), m, n,
                                                                           L.ld);
    LAUNCHED(s);
    // X_i Z_j = solve(cy, XY^T)^T (:366) as cy^-1 XY
    LCX_TRY(run_inverse(s, cy, L.ldm, m, s->ptr(I_RYINV), L.ldm, s->ptr(I_AUG), (int*)s->ptr(I_STATUS)));
    {
        GemmArgs a;
        memset(&a, 0, sizeof(a));
        a.A = s->ptr(I_RYINV); a.B = XY; a.C = s->ptr(LCX_A_XZ);
        a.M = m; a.N = n; a.K = m;
        a.lda = L.ldm; a.ldb = L.ld; a.ldc = L.ld;
        LCX_TRY(run_gemm(s, kLayoutKN, L.plan_mn, a, nullptr, 0));
    }
    // sqrtY currently holds sqrt(Yj2); details_finish rewrites it with the same values
    LCX_TRY(details_tail(s, XY, s->ptr(LCX_A_YJ2)));
    syn_tc_kernel<<<1, 32, 0, s->stream>>>(s->ptr(LCX_A_SCALARS) + 4, s->ptr(LCX_A_SCALARS));
    LAUNCHED(s);
    LCX_CUDA(cudaGetLastError());
    LCX_TRY(read_mailbox(s));
    if (tc) *tc = s->mailbox[0];
    if (additivity) *additivity = s->mailbox[6];
    return 0;
}

extern "C" int lcx_update_syn(lcx_session* s, double eta, double* tc, double* additivity) {
    S_REQUIRE_BOUND(s);
    const Layout& L = s->L;
    const int m = s->m, n = s->n;
    double* W = s->ptr(LCX_A_W);
    double* Rm = s->ptr(I_T);
    double* H = s->ptr(LCX_A_CY);       // rebuilt by moments_syn below
    double* Sm = s->ptr(LCX_A_GRAD);
    syn_colscale_kernel<<<grid_mn(m, n), 256, 0, s->stream>>>(s->ptr(LCX_A_XZ), s->ptr(LCX_A_X2Y), Rm, m, n, L.ld);
    LAUNCHED(s);
    {   // H = (XZ^T / X2Y) XZ, diag -> 0 (:378-379)
        GemmArgs a;
        memset(&a, 0, sizeof(a));
        a.A = Rm; a.B = s->ptr(LCX_A_XZ); a.C = H;
        a.M = m; a.N = m; a.K = n;
        a.lda = L.ld; a.ldb = L.ld; a.ldc = L.ldm;
        LCX_TRY(run_square_gemm(s, a, 0.0, nullptr));
    }
    {   // S = H W (:381)
        GemmArgs a;
        memset(&a, 0, sizeof(a));
        a.A = H; a.B = W; a.C = Sm;
        a.M = m; a.N = n; a.K = m;
        a.lda = L.ldm; a.ldb = L.ld; a.ldc = L.ld;
        LCX_TRY(run_gemm(s, kLayoutKN, L.plan_mn, a, nullptr, 0));
    }
    syn_mix_kernel<<<grid_mn(m, n), 256, 0, s->stream>>>(W, Rm, Sm, eta, W, m, n, L.ld);
    LAUNCHED(s);
    LCX_CUDA(cudaGetLastError());
    return lcx_moments_syn(s, tc, additivity);
}

extern "C" int lcx_get_covariance(lcx_session* s, int synergy, double eps, const double* sd, int row0, int rows, double* out,
                                  long long ldc) {
    S_REQUIRE_BOUND(s);
    const Layout& L = s->L;
    const int m = s->m, n = s->n;
    LCX_REQUIRE(sd && out, "null argument");
    LCX_REQUIRE(row0 >= 0 && rows > 0 && row0 + rows <= n && row0 % 2 == 0, "bad row block (row0 must be even)");
    LCX_REQUIRE(ldc >= n && ldc % 2 == 0, "ldc must be even and >= n");
    const double* left;
    const double* right;
    if (!synergy) {
        double* z = s->ptr(I_T);
        cov_z_kernel<<<grid_mn(m, n), 256, 0, s->stream>>>(s->ptr(LCX_A_RHOINVRHO), s->ptr(LCX_A_SI), z, m, n, L.ld);
        LAUNCHED(s);
        left = z;
        right = z;
    } else {
        left = s->ptr(LCX_A_XZ);
        right = s->ptr(LCX_A_XY);
    }
    GemmPlan pl = plan_gemm(rows, n, m, kSMs, 1, false);
    GemmArgs a;
    memset(&a, 0, sizeof(a));
    a.A = left + row0; a.B = right; a.C = out;
    a.M = rows; a.N = n; a.K = m;
    a.lda = L.ld; a.ldb = L.ld; a.ldc = ldc;
    LCX_TRY(launch_gemm(kLayoutMN, pl, a, s->stream));
    LAUNCHED(s);
    cov_finish_kernel<<<dim3(cdiv(n, 256), rows), 256, 0, s->stream>>>(out, ldc, row0, rows, n,
                                                                     synergy ? 1.0 : (1.0 - eps * eps), sd);
    LAUNCHED(s);
    LCX_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int lcx_gemm_f64(lcx_session* s, int layout, int M, int N, int K, const double* a, long long lda, const double* b,
                            long long ldb, double* c, long long ldc, int trans_out, const double* cadd, int max_splits,
                            double* scratch, long long scratch_doubles) {
    LCX_REQUIRE(s && a && b && c, "null argument");
    LCX_REQUIRE(layout >= 0 && layout <= 2 && M > 0 && N > 0 && K > 0, "bad shape/layout");
    LCX_CUDA(cudaSetDevice(s->device));
    const long long out_rows = trans_out ? N : M;
    const long long out_count = out_rows * ldc;
    int ms = max_splits;
    if (scratch == nullptr || cadd != nullptr) ms = 1;
    if (ms > 1 && out_count > 0) ms = (int)min((long long)ms, scratch_doubles / out_count);
    if (ms < 1) ms = 1;
    GemmPlan pl = plan_gemm(M, N, K, kSMs, ms, ms > 1);
    GemmArgs g;
    memset(&g, 0, sizeof(g));
    g.A = a; g.B = b; g.C = c; g.Cadd = cadd;
    g.M = M; g.N = N; g.K = K;
    g.lda = lda; g.ldb = ldb; g.ldc = ldc;
    g.trans_out = trans_out;
    return run_gemm(s, (GemmLayout)layout, pl, g, scratch, out_count);
}

extern "C" int lcx_inverse(lcx_session* s, const double* a, long long lda, int m, double* out, long long ldo, double* aug) {
    LCX_REQUIRE(s && a && out && aug && m > 0, "bad argument");
    LCX_CUDA(cudaSetDevice(s->device));
    return run_inverse(s, a, lda, m, out, ldo, aug, (int*)(aug + 2LL * m * m));
}
