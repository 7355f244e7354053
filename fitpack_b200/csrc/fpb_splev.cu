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
    while (g < 4 * nint && g < 1024) g <<= 1;
    return g;
}

// per-warp shared memory in doubles: two (t, c) buffers [2][2][ld] (the next spline's tables are
// prefetched while the current one is evaluated), + fast mode [dlev: K*ld][rc: K*ld]
// [rec: 8*(ld-2K-1)]
// followed by the lookup grid (ints, counted here in doubles)
__host__ __device__ inline int splev_smem_doubles(int mode, int k, int ld)
{
    const int grid = (splev_grid_cells(ld - 2 * k - 1) + 3) / 4;  // uint16 entries
    return ((mode == 1) ? ld * (4 + k + 8 + k) - 8 * (2 * k + 1) : ld * 4) + 2 * 256 /* x tiles */ + grid;
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
__device__ __forceinline__ int locate(const double *ts, const unsigned short *grid, int gcells, double tb,
                                      double inv_h, int k1, int nk1, double arg)
{
    int g = __double2int_rz((arg - tb) * inv_h);
    g = max(0, min(g, gcells - 1));
    int l = grid[g] & 0x7fff;
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

// Build one 64-byte record per knot interval l of spline (ts,cs):
//   rec[(l-k1)*8 + 0..1] = t(l), t(l+1);  rec[(l-k1)*8 + 2 + r] = a_r  with
//   s(x) = sum_r a_r * (x - t(l))^r   on interval l   (unused slots are zero).
// dlev[r*ld + i] (r = 1..K) receives the derivative pyramid
//   d^r_i = (K-r+1)/r * (d^{r-1}_i - d^{r-1}_{i-1}) / (t_{i+K-r+1} - t_i),   d^0 = c,
// i.e. the B-spline coefficients (degree K-r) of s^(r)/r!.  One de Boor-Cox recurrence
// per interval at x = t(l) then yields the basis of every degree 0..K on the way up.
template <int K>
__device__ void build_local_polys(const double *ts, const double *cs, int n, int ld, double *dlev,
                                  double *rec, double *rc, int lane)
{
    constexpr int K1 = K + 1;
    const int nk1 = n - K1;
    const int nint = nk1 - K1 + 1;
    // reciprocal knot spans rc[(j-1)*ld + a] = 1 / (t[a+j] - t[a]) (0-based a), j = 1..K; every
    // denominator of the derivative pyramid and of the basis recurrences is one of them.
    // Coincident knots give 0, which reproduces fpbspl's "skip the term" branch.
#pragma unroll
    for (int j = 1; j <= K; ++j) {
        for (int a = lane; a + j < n; a += 32) {
            const double den = ts[a + j] - ts[a];
            rc[(size_t)(j - 1) * ld + a] = (fabs(den) >= DBL_MIN) ? fast_rcp(den) : 0.0;
        }
    }
    __syncwarp();
    for (int r = 1; r <= K; ++r) {
        const double *prev = (r == 1) ? cs : dlev + (size_t)(r - 2) * ld;
        double *cur = dlev + (size_t)(r - 1) * ld;
        const double scale = (double)(K - r + 1) / (double)r;
        const double *rcj = rc + (size_t)(K - r) * ld;  // span K-r+1
        for (int i0 = lane; i0 < nk1; i0 += 32) {
            double v = 0.0;
            if (i0 >= r) v = scale * (prev[i0] - prev[i0 - 1]) * rcj[i0];
            cur[i0] = v;
        }
        __syncwarp();
    }
    for (int li = lane; li < nint; li += 32) {
        const int l = K1 + li;       // 1-based interval index
        const double x = ts[l - 1];  // t(l)
        double h[K1];
        h[0] = 1.0;
        double *rl = rec + (size_t)li * 8;
        rl[0] = x;
        rl[1] = ts[l];
#pragma unroll
        for (int r = K + 1; r < 6; ++r) rl[2 + r] = 0.0;
        rl[2 + K] = (K >= 1) ? dlev[(size_t)(K - 1) * ld + (l - 1)] : cs[l - 1];
#pragma unroll
        for (int j = 1; j <= K; ++j) {
            double hh[K];
#pragma unroll
            for (int i = 0; i < j; ++i) hh[i] = h[i];
            h[0] = 0.0;
            const double *rcj = rc + (size_t)(j - 1) * ld;
#pragma unroll
            for (int i = 1; i <= j; ++i) {
                const double tli = ts[l + i - 1], tlj = ts[l + i - j - 1];
                const double f = hh[i - 1] * rcj[l + i - j - 1];
                h[i - 1] = fma(f, tli - x, h[i - 1]);
                h[i] = f * (x - tlj);
            }
            const int r = K - j;  // basis of degree j pairs with derivative level K-j
            const double *lev = (r == 0) ? cs : dlev + (size_t)(r - 1) * ld;
            double acc = 0.0;
#pragma unroll
            for (int q = 0; q <= j; ++q) acc = fma(lev[l - j - 1 + q], h[q], acc);
            rl[2 + r] = acc;
        }
    }
    __syncwarp();
}

constexpr int kUnroll = 4;   // independent points per lane per group (instruction-level parallelism)
constexpr int kXTile = 256;  // evaluation points per TMA tile (2 KB), double buffered per warp

// Evaluate one group of kUnroll points per lane (indices ibase + 32*u) and store the results.
template <int K, int MODE, bool PLAIN, bool ALLLIVE>
__device__ __forceinline__ void eval_group(const SplevParams &p, const double *ts, const double *cs,
                                           const double *rec, const unsigned short *grid, int gcells, double tb,
                                           double te, double inv_h, int nk1, double (&arg)[kUnroll],
                                           const bool (&live)[kUnroll], double *yg, int *lg,
                                           int ibase, bool &out_of_range)
{
    constexpr int K1 = K + 1;
    bool zero[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) zero[u] = false;
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
    double sp[kUnroll];
    if (MODE == 1) {
        // FAST: one 64-byte record per interval {t(l), t(l+1), a0..a5}; the lookup grid lands in
        // the interval of the cell's left edge, a predicated single step covers one knot inside
        // the cell, the reference's own walk handles flagged cells (several knots) and edge rounding
        const double goff = -tb * inv_h;
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            int g = __double2int_rz(fma(arg[u], inv_h, goff));
            g = max(0, min(g, gcells - 1));
            l[u] = grid[g];
        }
        double tl[kUnroll];
        bool slow = false;
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            const bool cplx = (l[u] >> 15) != 0;
            const int l0 = l[u] & 0x7fff;
            const double2 tt = reinterpret_cast<const double2 *>(rec + (size_t)(l0 - K1) * 8)[0];
            const bool step = (arg[u] >= tt.y) && (l0 < nk1);
            l[u] = l0 + (step ? 1 : 0);
            tl[u] = step ? tt.y : tt.x;
            slow = slow || cplx || (arg[u] < tt.x && l0 > K1);
        }
        if (__any_sync(FULL, slow)) {
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) {
                int ll = l[u];
                while (ll > K1 && arg[u] < ts[ll - 1]) --ll;
                while (ll < nk1 && arg[u] >= ts[ll]) ++ll;
                l[u] = ll;
                tl[u] = ts[ll - 1];
            }
        }
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            const double uu = arg[u] - tl[u];
            const double2 *rp = reinterpret_cast<const double2 *>(rec + (size_t)(l[u] - K1) * 8);
            const double2 a01 = rp[1], a23 = rp[2], a45 = rp[3];
            const double a[6] = {a01.x, a01.y, a23.x, a23.y, a45.x, a45.y};
            double v = a[K];
#pragma unroll
            for (int r = K - 1; r >= 0; --r) v = fma(v, uu, a[r]);
            sp[u] = v;
        }
    } else {
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) l[u] = locate(ts, grid, gcells, tb, inv_h, K1, nk1, arg[u]);
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) sp[u] = eval_exact<K>(ts, cs, l[u], arg[u]);
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
        const int i = ibase + 32 * u;
        if (ALLLIVE || live[u]) {
            if (PLAIN) {
                yg[i] = sp[u];
            } else {
                yg[i] = zero[u] ? 0.0 : sp[u];
                if (lg) lg[i] = zero[u] ? 0 : l[u];
            }
        }
    }
}

// PLAIN: e == OUTSIDE_EXTRAPOLATE and no interval-index output -- no per-point policy code.
template <int K, int MODE, bool PLAIN>
__global__ void __launch_bounds__(kSplevWarps * 32, (MODE == 1) ? 3 : 2)
    splev_kernel(SplevParams p, int smem_ld)
{
    constexpr int K1 = K + 1;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    const int per_warp = splev_smem_doubles(MODE, K, smem_ld);
    uint64_t *mbars = reinterpret_cast<uint64_t *>(smem_raw);
    double *base = reinterpret_cast<double *>(smem_raw + 32 * kSplevWarps) + (size_t)warp * per_warp;
    auto tsb = [&](int b) { return base + (size_t)(2 * b) * smem_ld; };      // (t, c) buffer b
    auto csb = [&](int b) { return base + (size_t)(2 * b + 1) * smem_ld; };
    double *dlev = base + 4 * smem_ld;
    double *rc = base + (size_t)(4 + K) * smem_ld;
    double *rec = base + (size_t)(4 + 2 * K) * smem_ld;
    double *after = base + ((MODE == 1) ? smem_ld * (4 + K + 8 + K) - 8 * (2 * K + 1) : smem_ld * 4);
    double *ts = tsb(0), *cs = csb(0);
    int tbuf = 0;           // table buffer holding the current spline
    long long staged = -1;  // spline whose tables were prefetched into buffer 1-tbuf (TMA in flight)
    double *xs = after;                                         // [2][kXTile] evaluation points
    unsigned short *grid = reinterpret_cast<unsigned short *>(after + 2 * kXTile);  // lookup grid
    uint64_t *mbar = mbars + warp * 4;                          // [0] tables, [1..2] x tiles
    if (lane == 0) {
        mbar_init(mbar, 1);
        mbar_init(mbar + 1, 1);
        mbar_init(mbar + 2, 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncwarp();
    uint32_t phase = 0, xphase0 = 0, xphase1 = 0;

    const long long gwarp = (long long)blockIdx.x * wpb + warp;
    const long long nwarps = (long long)gridDim.x * wpb;
    int cached = -1;
    int n = 0, gcells = 32;
    double inv_h = 0.0;
    for (long long j = gwarp; j < p.nspl; j += nwarps) {
        const long long tj = p.nshared ? 0 : j;
        {
            // a knot count the row cannot hold, or too small for degree K, is invalid input
            // (splev itself would read out of bounds): flag the spline and skip it
            const int nj = p.n[tj];
            if (nj < 2 * K1 || nj > p.ldt) {
                if (p.ier && lane == 0) p.ier[j] = IER_INPUT;
                continue;
            }
        }
        const double *xg = p.x + j * p.m;
        // TMA tiles need 16-byte aligned rows of even length; otherwise plain loads
        // (EXACT mode is FP64-bound, not latency-bound: it keeps the register-pipelined loads)
        const bool xt_ok = (MODE == 1) && ((((uintptr_t)xg) & 15) == 0) && ((p.m & 1) == 0);
        const int ntile = (p.m + kXTile - 1) / kXTile;
        auto issue_tile = [&](int t) {
            if (lane == 0) {
                const int cnt = min(kXTile, p.m - t * kXTile);
                uint64_t *mb = mbar + 1 + (t & 1);
                asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
                mbar_expect_tx(mb, (uint32_t)cnt * 8u);
                tma_load_1d(xs + (t & 1) * kXTile, xg + (size_t)t * kXTile, (uint32_t)cnt * 8u, mb);
            }
        };
        if (xt_ok) {  // the points start streaming in while the spline's tables are built
            issue_tile(0);
            if (ntile > 1) issue_tile(1);
        }
        if (tj != cached) {
            n = p.n[tj];
            __syncwarp();  // previous spline's readers are done with their tables
            auto tables_tma_ok = [&](long long q, int nq) -> bool {
                const double *tg = p.t + q * p.ldt, *cg = p.c + q * p.ldt;
                return ((((uintptr_t)tg) | ((uintptr_t)cg)) & 15) == 0 && (((nq + 1) & ~1) <= p.ldt);
            };
            auto issue_tables = [&](long long q, int nq, int buf) {
                if (lane == 0) {
                    const uint32_t bytes = (uint32_t)(((nq + 1) & ~1) * 8);
                    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
                    mbar_expect_tx(mbar, 2 * bytes);
                    tma_load_1d(tsb(buf), p.t + q * p.ldt, bytes, mbar);
                    tma_load_1d(csb(buf), p.c + q * p.ldt, bytes, mbar);
                }
            };
            if (staged == tj) {  // prefetched while the previous spline was evaluated
                tbuf ^= 1;
                mbar_wait(mbar, phase);
                phase ^= 1;
            } else if (tables_tma_ok(tj, n)) {
                issue_tables(tj, n, tbuf);
                mbar_wait(mbar, phase);
                phase ^= 1;
            } else {
                const double *tg = p.t + tj * p.ldt, *cg = p.c + tj * p.ldt;
                for (int i = lane; i < n; i += 32) {
                    tsb(tbuf)[i] = tg[i];
                    csb(tbuf)[i] = cg[i];
                }
            }
            ts = tsb(tbuf);
            cs = csb(tbuf);
            staged = -1;
            // start fetching the next spline's tables into the other buffer
            {
                const long long jn = j + nwarps;
                if (!p.nshared && jn < p.nspl) {
                    const int nn = p.n[jn];
                    if (nn >= 2 * K1 && tables_tma_ok(jn, nn)) {
                        issue_tables(jn, nn, tbuf ^ 1);
                        staged = jn;
                    }
                }
            }
            __syncwarp();
            // lookup grid over [t(k+1), t(n-k)]: grid[g] = interval of the cell's left edge;
            // bit 15 flags a cell that may hold two or more knots (the single predicated step of
            // the evaluation loop cannot resolve it: the full walk is taken).  Edges are widened
            // by 1e-6 of a cell so that rounding in the cell index can never un-flag a cell.
            {
                const int nk1_ = n - K1;
                const int nint = nk1_ - K1 + 1;
                gcells = splev_grid_cells(nint);
                const double tb_ = ts[K1 - 1], te_ = ts[nk1_];
                const double span = te_ - tb_;
                inv_h = (span > 0.0) ? (double)gcells / span : 0.0;
                const double hcell = span / (double)gcells;
                // Each lane owns knot intervals l = k1+lane, k1+lane+32, ... and fills the cells whose
                // (widened) left edge falls into [t(l), t(l+1)): cell boundaries come from one
                // monotone function of the knot, so the ranges of neighbouring lanes tile the grid.
                auto bnd = [&](int l) -> int {  // first cell whose left edge is >= t(l)
                    if (l <= K1) return 0;
                    if (l > nk1_) return gcells;
                    const double v = (ts[l - 1] - tb_) * inv_h + 1e-6;
                    int c = __double2int_ru(v);
                    return max(0, min(c, gcells));
                };
                for (int li = lane; li < nint; li += 32) {
                    const int l = K1 + li;
                    const int g0 = bnd(l), g1 = bnd(l + 1);
                    const double t2 = (l + 2 <= nk1_) ? ts[l + 1] : 1.79e308;  // t(l+2), if stepping twice is possible
                    for (int gI = g0; gI < g1; ++gI) {
                        const double right = tb_ + ((double)(gI + 1) + 1e-6) * hcell;
                        grid[gI] = (unsigned short)(l | ((right >= t2) ? (1 << 15) : 0));
                    }
                }
            }
            if (MODE == 1) build_local_polys<K>(ts, cs, n, smem_ld, dlev, rec, rc, lane);
            __syncwarp();
            cached = (int)tj;
        }
        const int nk1 = n - K1;
        const double tb = ts[K1 - 1], te = ts[nk1];
        double *yg = p.y + j * p.m;
        int *lg = p.lidx ? p.lidx + j * p.m : nullptr;
        bool out_of_range = false;
        if (xt_ok) {
            for (int t = 0; t < ntile; ++t) {
                const int buf = t & 1;
                if (buf == 0) { mbar_wait(mbar + 1, xphase0); xphase0 ^= 1; }
                else { mbar_wait(mbar + 2, xphase1); xphase1 ^= 1; }
                const int cnt = min(kXTile, p.m - t * kXTile);
                const double *xt = xs + buf * kXTile;
                if (cnt == kXTile) {  // full tile: no per-point bounds predicates
#pragma unroll
                    for (int gq = 0; gq < kXTile / (32 * kUnroll); ++gq) {
                        const int il = gq * 32 * kUnroll + lane;
                        double arg[kUnroll];
                        bool live[kUnroll];
#pragma unroll
                        for (int u = 0; u < kUnroll; ++u) {
                            live[u] = true;
                            arg[u] = xt[il + 32 * u];
                        }
                        eval_group<K, MODE, PLAIN, true>(p, ts, cs, rec, grid, gcells, tb, te, inv_h,
                                                         nk1, arg, live, yg, lg, t * kXTile + il,
                                                         out_of_range);
                    }
                } else {
#pragma unroll
                    for (int gq = 0; gq < kXTile / (32 * kUnroll); ++gq) {
                        const int il = gq * 32 * kUnroll + lane;
                        if (il - lane < cnt) {  // uniform across the warp
                            double arg[kUnroll];
                            bool live[kUnroll];
#pragma unroll
                            for (int u = 0; u < kUnroll; ++u) {
                                live[u] = il + 32 * u < cnt;
                                arg[u] = live[u] ? xt[il + 32 * u] : tb;
                            }
                            eval_group<K, MODE, PLAIN, false>(p, ts, cs, rec, grid, gcells, tb, te,
                                                              inv_h, nk1, arg, live, yg, lg,
                                                              t * kXTile + il, out_of_range);
                        }
                    }
                }
                __syncwarp();  // every lane is done with this buffer
                if (t + 2 < ntile) issue_tile(t + 2);
            }
        } else {
            // software pipeline in registers: the loads of the next group are in flight while
            // the current group is evaluated
            double nxt[kUnroll];
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) {
                const int i = lane + 32 * u;
                nxt[u] = (i < p.m) ? xg[i] : tb;
            }
            for (int i0 = lane; i0 < p.m; i0 += 32 * kUnroll) {
                double arg[kUnroll];
                bool live[kUnroll];
#pragma unroll
                for (int u = 0; u < kUnroll; ++u) {
                    const int i = i0 + 32 * u;
                    live[u] = i < p.m;
                    arg[u] = nxt[u];
                    const int inx = i + 32 * kUnroll;
                    nxt[u] = (inx < p.m) ? xg[inx] : tb;
                }
                eval_group<K, MODE, PLAIN, false>(p, ts, cs, rec, grid, gcells, tb, te, inv_h, nk1,
                                                  arg, live, yg, lg, i0, out_of_range);
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
    while (wpb > 1 && 32 * kSplevWarps + per_warp * wpb > 100 * 1024) wpb >>= 1;
    const size_t smem = 32 * kSplevWarps + per_warp * wpb;
    if (smem > 220 * 1024) return cudaErrorInvalidValue;  // knot vectors beyond ~4000 entries
    long long blocks = ((long long)p.nspl + wpb - 1) / wpb;
    const long long resident = (long long)((228 * 1024) / (smem + 1024));  // 1 KB/CTA is reserved
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
