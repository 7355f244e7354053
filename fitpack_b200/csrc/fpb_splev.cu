// fpb_splev.cu -- batched B-spline evaluation (splev/fpbspl) for sm_100a.
//
// Reference: splev src/fitpack_core.F90:17826-17896, fpbspl 2387-2413.
//
// One warp per spline: the warp stages the spline's knots and coefficients in
// shared memory once (TMA bulk copy when the rows are 16-byte aligned), then
// streams the spline's m evaluation points through its 32 lanes with fully
// coalesced 256 B loads/stores.  Each lane finds its knot interval with a
// branch-free upper-bound binary search over the staged knots; the result is
// the history-independent interval the reference's bidirectional walk ends in
// (rightmost l in [k+1, n-k-1] with t(l) <= arg, core:17880-17887).
//
// mode FPB_SPLEV_EXACT: de Boor-Cox recurrence + (k+1)-term dot product in the
//   reference's operation order, no FMA: values bit-identical to the reference.
// mode FPB_SPLEV_FAST: the warp first converts the spline to one local Taylor
//   polynomial per knot interval (derivative pyramid + one basis evaluation per
//   interval, all in shared memory), then evaluates with Horner/FMA -- ~12
//   FP64 instructions per point instead of ~400, which is what makes the
//   kernel HBM-bound (16 B/point).  Same interval indices; values agree with
//   the exact mode to ~1e-15*max|c| (tests state 1e-12).
#include "fpb_common.cuh"
#include "fpb_kernels.h"

namespace fpb {

namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr int kSplevWarps = 8;

__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}

// One-shot TMA bulk copy global -> shared of `bytes` (multiple of 16, both
// addresses 16 B aligned), issued by one lane, completion on an mbarrier.
__device__ __forceinline__ void tma_load_1d(void *dst_smem, const void *src_gmem, uint32_t bytes,
                                            uint64_t *mbar)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
        :
        : "r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(mbar))
        : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t *mbar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(mbar)), "r"(count)
                 : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *mbar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(mbar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *mbar, uint32_t phase)
{
    uint32_t done = 0;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(smem_u32(mbar)), "r"(phase)
            : "memory");
    } while (!done);
}

// number of cells of the interval lookup grid for a spline with `nint` knot intervals
__host__ __device__ inline int splev_grid_cells(int nint)
{
    int g = 32;
    while (g < 2 * nint && g < 1024) g <<= 1;
    return g;
}

// per-warp shared memory in doubles: [t: ld][c: ld] (+ fast mode [dlev: K*ld][poly: K1*ld])
// followed by the lookup grid (ints, counted here in doubles)
__host__ __device__ inline int splev_smem_doubles(int mode, int k, int ld)
{
    const int grid = splev_grid_cells(ld - 2 * k - 1) / 2;
    return ((mode == 1) ? ld * (2 + k + (k + 1)) : ld * 2) + grid;
}

// 1/d to ~1 ulp without the IEEE division sequence (FAST mode only): hardware seed + two
// Newton steps; falls back to a true division outside the safe exponent range.
__device__ __forceinline__ double fast_rcp(double d)
{
    const double ad = fabs(d);
    if (!(ad > 1e-290 && ad < 1e290)) return 1.0 / d;
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
    double e = fma(-d, r, 1.0);
    r = fma(r, e, r);
    e = fma(-d, r, 1.0);
    r = fma(r, e, r);
    return r;
}

// interval search: l = k1 + #{ i in (k1, nk1] : t(i) <= arg }  (1-based), t non-decreasing.
// Fixed number of halving steps (`top` = largest power of two <= nk1-k1, uniform per spline),
// so the 32 lanes never diverge.
__device__ __forceinline__ int find_interval(const double *ts, int k1, int nk1, int top, double arg)
{
    int lo = k1;
    for (int step = top; step > 0; step >>= 1) {
        const int cand = lo + step;
        if (cand <= nk1 && ts[cand - 1] <= arg) lo = cand;
    }
    return lo;
}

// Interval of `arg` from the lookup grid + the reference's own bidirectional walk
// (core:17880-17887), which makes the result independent of the quality of the guess.
__device__ __forceinline__ int locate(const double *ts, const int *grid, int gcells, double tb,
                                      double inv_h, int k1, int nk1, double arg)
{
    int g = __double2int_rz((arg - tb) * inv_h);
    g = max(0, min(g, gcells - 1));
    int l = grid[g];
    while (l > k1 && arg < ts[l - 1]) --l;
    while (l < nk1 && arg >= ts[l]) ++l;
    return l;
}

template <int K>
__device__ __forceinline__ double eval_exact(const double *ts, const double *cs, int l, double arg)
{
    double tw[2 * K], h[K + 1];
#pragma unroll
    for (int j = 0; j < 2 * K; ++j) tw[j] = ts[l - K + j];  // t(l-K+1+j), 0-based index l-K+j
    bspl<K>(tw, arg, h);
    double sp = 0.0;
#pragma unroll
    for (int j = 0; j <= K; ++j) sp = add(sp, mul(cs[l - K - 1 + j], h[j]));  // c(l-K+j)
    return sp;
}

// Build the per-interval Taylor coefficients of spline (ts,cs):
//   s(x) = sum_r poly[(l-k1)*(K+1)+r] * (x - t(l))^r   on knot interval l.
// dlev[r*ld + i] (r = 1..K) receives the derivative pyramid
//   d^r_i = (K-r+1)/r * (d^{r-1}_i - d^{r-1}_{i-1}) / (t_{i+K-r+1} - t_i),   d^0 = c,
// i.e. the B-spline coefficients (degree K-r) of s^(r)/r!.  One de Boor-Cox recurrence
// per interval at x = t(l) then yields the basis of every degree 0..K on the way up.
template <int K>
__device__ void build_local_polys(const double *ts, const double *cs, int n, int ld, double *dlev,
                                  double *poly, int lane)
{
    constexpr int K1 = K + 1;
    const int nk1 = n - K1;
    const int nint = nk1 - K1 + 1;
    for (int r = 1; r <= K; ++r) {
        const double *prev = (r == 1) ? cs : dlev + (size_t)(r - 2) * ld;
        double *cur = dlev + (size_t)(r - 1) * ld;
        const double scale = (double)(K - r + 1) / (double)r;
        for (int i0 = lane; i0 < nk1; i0 += 32) {
            double v = 0.0;
            if (i0 >= r) {
                const double den = ts[i0 + K - r + 1] - ts[i0];
                if (fabs(den) >= DBL_MIN) v = scale * (prev[i0] - prev[i0 - 1]) * fast_rcp(den);
            }
            cur[i0] = v;
        }
        __syncwarp();
    }
    for (int li = lane; li < nint; li += 32) {
        const int l = K1 + li;       // 1-based interval index
        const double x = ts[l - 1];  // t(l)
        double h[K1];
        h[0] = 1.0;
        poly[li * K1 + K] = (K >= 1) ? dlev[(size_t)(K - 1) * ld + (l - 1)] : cs[l - 1];
#pragma unroll
        for (int j = 1; j <= K; ++j) {
            double hh[K];
#pragma unroll
            for (int i = 0; i < j; ++i) hh[i] = h[i];
            h[0] = 0.0;
#pragma unroll
            for (int i = 1; i <= j; ++i) {
                const double tli = ts[l + i - 1], tlj = ts[l + i - j - 1];
                const double d = tli - tlj;
                if (fabs(d) >= DBL_MIN) {
                    const double f = hh[i - 1] * fast_rcp(d);
                    h[i - 1] = fma(f, tli - x, h[i - 1]);
                    h[i] = f * (x - tlj);
                } else {
                    h[i] = 0.0;
                }
            }
            const int r = K - j;  // basis of degree j pairs with derivative level K-j
            const double *lev = (r == 0) ? cs : dlev + (size_t)(r - 1) * ld;
            double acc = 0.0;
#pragma unroll
            for (int q = 0; q <= j; ++q) acc = fma(lev[l - j - 1 + q], h[q], acc);
            poly[li * K1 + r] = acc;
        }
    }
    __syncwarp();
}

constexpr int kUnroll = 4;  // independent points per lane per iteration (memory-level parallelism)

// PLAIN: e == OUTSIDE_EXTRAPOLATE and no interval-index output -- no per-point policy code.
template <int K, int MODE, bool PLAIN>
__global__ void __launch_bounds__(kSplevWarps * 32, (MODE == 1) ? 4 : 2)
    splev_kernel(SplevParams p, int smem_ld)
{
    constexpr int K1 = K + 1;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    const int per_warp = splev_smem_doubles(MODE, K, smem_ld);
    uint64_t *mbars = reinterpret_cast<uint64_t *>(smem_raw);
    double *base = reinterpret_cast<double *>(smem_raw + 16 * kSplevWarps) + (size_t)warp * per_warp;
    double *ts = base, *cs = base + smem_ld, *dlev = base + 2 * smem_ld,
           *poly = base + (size_t)(2 + K) * smem_ld;
    int *grid = reinterpret_cast<int *>(base + ((MODE == 1) ? smem_ld * (2 + K + K1) : smem_ld * 2));
    uint64_t *mbar = mbars + warp * 2;
    if (lane == 0) {
        mbar_init(mbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncwarp();
    uint32_t phase = 0;

    const long long gwarp = (long long)blockIdx.x * wpb + warp;
    const long long nwarps = (long long)gridDim.x * wpb;
    int cached = -1;
    int n = 0, gcells = 32;
    double inv_h = 0.0;
    for (long long j = gwarp; j < p.nspl; j += nwarps) {
        const long long tj = p.nshared ? 0 : j;
        if (tj != cached) {
            n = p.n[tj];
            const double *tg = p.t + tj * p.ldt, *cg = p.c + tj * p.ldt;
            const uint32_t bytes = (uint32_t)(((n + 1) & ~1) * 8);
            const bool tma_ok = ((((uintptr_t)tg) | ((uintptr_t)cg)) & 15) == 0 &&
                                (((n + 1) & ~1) <= p.ldt);
            __syncwarp();  // previous spline's readers are done with ts/cs
            if (tma_ok) {
                if (lane == 0) {
                    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
                    mbar_expect_tx(mbar, 2 * bytes);
                    tma_load_1d(ts, tg, bytes, mbar);
                    tma_load_1d(cs, cg, bytes, mbar);
                }
                mbar_wait(mbar, phase);
                phase ^= 1;
            } else {
                for (int i = lane; i < n; i += 32) {
                    ts[i] = tg[i];
                    cs[i] = cg[i];
                }
            }
            __syncwarp();
            // lookup grid over [t(k+1), t(n-k)]: grid[g] = interval of the cell's left edge
            {
                const int nk1_ = n - K1;
                const int nint = nk1_ - K1 + 1;
                gcells = splev_grid_cells(nint);
                const double tb_ = ts[K1 - 1], te_ = ts[nk1_];
                const double span = te_ - tb_;
                inv_h = (span > 0.0) ? (double)gcells / span : 0.0;
                const double hcell = span / (double)gcells;
                int top = 1;
                while (top * 2 <= nk1_ - K1) top *= 2;
                if (nk1_ - K1 < 1) top = 0;
                for (int gI = lane; gI < gcells; gI += 32)
                    grid[gI] = find_interval(ts, K1, nk1_, top, tb_ + (double)gI * hcell);
            }
            if (MODE == 1) build_local_polys<K>(ts, cs, n, smem_ld, dlev, poly, lane);
            __syncwarp();
            cached = (int)tj;
        }
        const int nk1 = n - K1;
        const double tb = ts[K1 - 1], te = ts[nk1];
        const double *xg = p.x + j * p.m;
        double *yg = p.y + j * p.m;
        int *lg = p.lidx ? p.lidx + j * p.m : nullptr;
        bool out_of_range = false;
        for (int i0 = lane; i0 < p.m; i0 += 32 * kUnroll) {
            double arg[kUnroll];
            bool live[kUnroll], zero[kUnroll];
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) {
                const int i = i0 + 32 * u;
                live[u] = i < p.m;
                arg[u] = live[u] ? xg[i] : tb;
                zero[u] = false;
            }
            if (!PLAIN) {
#pragma unroll
                for (int u = 0; u < kUnroll; ++u) {
                    if (arg[u] < tb || arg[u] > te) {  // core:17864-17877
                        if (p.e == 1) zero[u] = true;
                        else if (p.e == 2) out_of_range = out_of_range || live[u];
                        else if (p.e == 3) arg[u] = fmax(fmin(arg[u], te), tb);
                    }
                }
            }
            int l[kUnroll];
#pragma unroll
            for (int u = 0; u < kUnroll; ++u)
                l[u] = locate(ts, grid, gcells, tb, inv_h, K1, nk1, arg[u]);
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) {
                double sp;
                if (MODE == 1) {
                    const double *pl = poly + (l[u] - K1) * K1;
                    const double uu = arg[u] - ts[l[u] - 1];
                    sp = pl[K];
#pragma unroll
                    for (int r = K - 1; r >= 0; --r) sp = fma(sp, uu, pl[r]);
                } else {
                    sp = eval_exact<K>(ts, cs, l[u], arg[u]);
                }
                const int i = i0 + 32 * u;
                if (live[u]) {
                    if (PLAIN) {
                        yg[i] = sp;
                    } else {
                        yg[i] = zero[u] ? 0.0 : sp;
                        if (lg) lg[i] = zero[u] ? 0 : l[u];
                    }
                }
            }
        }
        if (p.ier) {
            const bool any_out = __any_sync(FULL, out_of_range);
            if (lane == 0) p.ier[j] = any_out ? IER_RANGE : IER_OK;
        }
    }
}

}  // namespace

cudaError_t launch_splev(const SplevParams &p, int nsm, cudaStream_t stream)
{
    if (p.nspl <= 0) return cudaSuccess;
    const int smem_ld = (p.ldt + 1) & ~1;  // keep every per-warp region 16 B aligned
    const size_t per_warp = (size_t)splev_smem_doubles(p.mode, p.k, smem_ld) * sizeof(double);
    int wpb = kSplevWarps;
    while (wpb > 1 && 16 * kSplevWarps + per_warp * wpb > 100 * 1024) wpb >>= 1;
    const size_t smem = 16 * kSplevWarps + per_warp * wpb;
    if (smem > 220 * 1024) return cudaErrorInvalidValue;  // knot vectors beyond ~4000 entries
    long long blocks = ((long long)p.nspl + wpb - 1) / wpb;
    const long long resident = (long long)(220 * 1024 / smem);
    const long long max_blocks = (long long)nsm * (resident < 1 ? 1 : (resident > 8 ? 8 : resident));
    if (blocks > max_blocks) blocks = max_blocks;
#define FPB_LAUNCH_SPLEV(KK)                                                                       \
    do {                                                                                           \
        const bool plain = (p.e == 0 && p.lidx == nullptr);                                       \
        auto kern = (p.mode == 1) ? (plain ? splev_kernel<KK, 1, true> : splev_kernel<KK, 1, false>)\
                                  : (plain ? splev_kernel<KK, 0, true> : splev_kernel<KK, 0, false>);\
        cudaError_t e_ = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,   \
                                              (int)smem);                                          \
        if (e_ != cudaSuccess) return e_;                                                          \
        kern<<<(unsigned)blocks, wpb * 32, smem, stream>>>(p, smem_ld);                            \
    } while (0)
    switch (p.k) {
    case 1: FPB_LAUNCH_SPLEV(1); break;
    case 2: FPB_LAUNCH_SPLEV(2); break;
    case 3: FPB_LAUNCH_SPLEV(3); break;
    case 4: FPB_LAUNCH_SPLEV(4); break;
    case 5: FPB_LAUNCH_SPLEV(5); break;
    default: return cudaErrorInvalidValue;
    }
#undef FPB_LAUNCH_SPLEV
    return cudaGetLastError();
}

}  // namespace fpb
