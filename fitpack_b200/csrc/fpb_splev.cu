// fpb_splev.cu -- batched B-spline evaluation (splev/fpbspl) for sm_100a.
//
// Reference: splev src/fitpack_core.F90:17826-17896, fpbspl 2387-2413.
//
// One warp per spline: the warp stages the spline's knots and coefficients in
// shared memory once (TMA bulk copy when the rows are 16-byte aligned), then
// streams the spline's m evaluation points through its 32 lanes with fully
// coalesced 256 B loads/stores.  Each lane finds its knot interval through a
// per-spline lookup grid; the result is always the history-independent interval
// the reference's bidirectional walk ends in (rightmost l in [k+1, n-k-1] with
// t(l) <= arg, core:17880-17887).
//
// mode FPB_SPLEV_EXACT (splev_kernel): de Boor-Cox recurrence + (k+1)-term dot
//   product in the reference's operation order, no FMA: values bit-identical to
//   the reference.  FP64-pipe bound (~400 FP64 slots per point).
// mode FPB_SPLEV_FAST (splev_fast_kernel): the warp first converts the spline to
//   one local Taylor polynomial per knot interval (each lane builds the records
//   of its intervals in registers), then evaluates with Horner/FMA.  The
//   evaluation points of the warp's whole spline sequence flow through a ring of
//   TMA tiles that runs ahead across spline boundaries, so neither the table
//   build nor the DRAM latency of the point stream is exposed.  Same interval
//   indices; values agree with the exact mode to ~1e-15*max|c| (tests: 1e-12).
#include <cstdlib>

#include "fpb_common.cuh"
#include "fpb_kernels.h"

namespace fpb {

namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr int kSplevWarps = 8;  // EXACT mode: warps per CTA (2 CTAs/SM); also sizes the mbarrier header
// FAST mode launch shape: warps per CTA x CTAs per SM (register budget 65536 / threads per SM)
#ifndef FPB_SPLEV_FAST_WARPS
#define FPB_SPLEV_FAST_WARPS 8
#endif
#ifndef FPB_SPLEV_FAST_BLOCKS
#define FPB_SPLEV_FAST_BLOCKS 2
#endif
constexpr int kFastWarps = FPB_SPLEV_FAST_WARPS, kFastBlocks = FPB_SPLEV_FAST_BLOCKS;

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

// FAST mode tables.  One record per knot interval, kRecD doubles apart (80 B: the STS.128 of a
// record written by 32 lanes are bank-conflict free per quarter warp; 64 B would be 16-way):
//   rec[0] = t(l) (expansion centre), rec[1 + r] = a_r, s(x) = sum_r a_r (x - t(l))^r on interval l.
// One 16-byte cell per lookup-grid cell: {t(l0+1) or +inf, byte offset of record l0, l0 | flag<<31}
// where l0 is the interval of the cell's (widened) left edge.
constexpr int kRecD = 10;
constexpr int kUnroll = 4;              // independent points per lane per group (instruction-level parallelism)
constexpr int kGroupPts = 32 * kUnroll;  // points one warp evaluates per group
#ifndef FPB_SPLEV_TILE_GROUPS
#define FPB_SPLEV_TILE_GROUPS 2
#endif
constexpr int kTileGroups = FPB_SPLEV_TILE_GROUPS;
constexpr int kTilePts = kGroupPts * kTileGroups;  // FAST: evaluation points per TMA tile (2 KB)
#ifndef FPB_SPLEV_NBUF
#define FPB_SPLEV_NBUF 2
#endif
constexpr int kNBuf = FPB_SPLEV_NBUF;   // FAST: tiles in flight per warp (ring, power of two)
constexpr int kMbarPerWarp = 16;        // mbarrier slots per warp: [0] tables, [1..kNBuf] tiles
static_assert(kNBuf + 1 <= kMbarPerWarp && (kNBuf & (kNBuf - 1)) == 0, "ring size");
static_assert(kFastWarps <= kSplevWarps, "mbarrier header is sized for kSplevWarps warps");

// per-warp shared memory in doubles: two (t, c) buffers [2][2][ld] (the next spline's tables are
// prefetched while the current one is evaluated); EXACT: + uint16 lookup grid; FAST: + interval
// records [kRecD*(ld-2K-1)], the 16-byte lookup cells and the ring of x tiles
// EXACT with tab != 0: + the span table of bspl_tab (k(k+1)/2 refined reciprocals per knot interval and one
// mask word per interval): fpbspl's divisions become three FP64 instructions each, same bits.
// `tab` = number of knot intervals the table holds (0: none); splines with more intervals use plain divisions.
__host__ __device__ inline int splev_smem_doubles(int mode, int k, int ld, int tab = 0)
{
    const int cells = splev_grid_cells(ld - 2 * k - 1);
    const int nintmax = (ld - 2 * k - 1 > 1) ? ld - 2 * k - 1 : 1;
    const int d = ld * 4 + ((mode == 1) ? ((kRecD * nintmax + 1) & ~1) + 2 * cells + kNBuf * kTilePts
                                        : (cells + 3) / 4 + (tab ? (k * (k + 1) / 2) * tab + (tab + 1) / 2 : 0));
    return (d + 1) & ~1;   // every warp's region starts 16-byte aligned (TMA destinations)
}

// 1/d by hardware seed + two Newton steps (FAST mode only; d must be a normal number well inside
// the exponent range -- the caller checks once per interval).
__device__ __forceinline__ double rcp_newton(double d)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
    double e = fma(-d, r, 1.0);
    r = fma(r, e, r);
    e = fma(-d, r, 1.0);
    r = fma(r, e, r);
    return r;
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

// EXACT mode with the span table: per knot interval the (refined) reciprocals of the K(K+1)/2 knot spans
// fpbspl divides by (fpb_common.cuh: the reciprocal-refinement half of the hardware's division sequence,
// done once per interval instead of once per point).  Each lane owns the intervals li = lane, lane+32, ...
template <int K>
__device__ __forceinline__ void build_spans(const double *ts, int n, double *spans, unsigned *spmask, int lane)
{
    constexpr int K1 = K + 1, NS = K * (K + 1) / 2;
    const int nint = n - K1 - K;
    for (int li = lane; li < nint; li += 32) {
        const int l = K1 + li;
        double tw[2 * K];
#pragma unroll
        for (int q = 0; q < 2 * K; ++q) tw[q] = ts[l - K + q];
        SpanTab<K, 1> sp;
        sp.rs = spans + (size_t)li * NS;
        span_init<K, 1>(sp, tw);
        spmask[li] = sp.special;
    }
}

template <int K>
__device__ __forceinline__ double eval_exact_tab(const double *ts, const double *cs, const double *spans,
                                                 const unsigned *spmask, int l, double arg)
{
    constexpr int NS = K * (K + 1) / 2;
    double tw[2 * K], h[K + 1];
#pragma unroll
    for (int j = 0; j < 2 * K; ++j) tw[j] = ts[l - K + j];
    SpanTab<K, 1> sp;
    sp.rs = const_cast<double *>(spans) + (size_t)(l - K - 1) * NS;
    sp.special = spmask[l - K - 1];
    bspl_tab<K, 1>(tw, sp, arg, h);
    double v = 0.0;
#pragma unroll
    for (int j = 0; j <= K; ++j) v = add(v, mul(cs[l - K - 1 + j], h[j]));
    return v;
}

// FAST mode: convert the spline (ts, cs) into one local Taylor polynomial per knot interval.
// Each lane owns the intervals l = k1+lane, k1+lane+32, ... and works entirely in registers on the
// 2K knots / K+1 coefficients around its interval (no cross-lane traffic, no scratch tables):
//   R[j][i]  = 1 / (t(l+i) - t(l+i-j)),  1 <= i <= j <= K -- the K(K+1)/2 spans that contain
//              [t(l), t(l+1)]; they are the denominators of BOTH recurrences below
//   D^r[e]   = (D^{r-1}[e] - D^{r-1}[e-1]) * R[K-r+1][e-r+1],  D^0[e] = c(l-K+e): the B-spline
//              coefficients of s^(r) / (K!/(K-r)!) restricted to the interval
//   h_j[0..j]= the degree-j basis at x = t(l) (de Boor-Cox, core:2387-2413)
//   a_r      = binom(K,r) * sum_q D^r[r+q] * h_{K-r}[q]   (= s^(r)(t(l)) / r!)
// Coincident knots give R = 0, which reproduces fpbspl's "skip the term" branch (core:2403).
template <int K>
__device__ void build_records(const double *ts, const double *cs, int n, double *rec, int lane)
{
    constexpr int K1 = K + 1;
    const int nk1 = n - K1;
    const int nint = nk1 - K;
    for (int li = lane; li < nint; li += 32) {
        const int l = K1 + li;  // 1-based interval index
        double tw[2 * K], D[K1];
#pragma unroll
        for (int q = 0; q < 2 * K; ++q) tw[q] = ts[l - K + q];  // tw[q] = t(l-K+1+q)
#pragma unroll
        for (int e = 0; e <= K; ++e) D[e] = cs[l - 1 - K + e];  // c(l-K+e)
        double R[K][K];
        // every span contains [t(l), t(l+1)] and lies inside [t(l-K+1), t(l+K)]
        const double dmin = tw[K] - tw[K - 1], dmax = tw[2 * K - 1] - tw[0];
        if (dmin > 1e-290 && dmax < 1e290) {
#pragma unroll
            for (int j = 1; j <= K; ++j)
#pragma unroll
                for (int i = 1; i <= j; ++i) R[j - 1][i - 1] = rcp_newton(tw[K - 1 + i] - tw[K - 1 + i - j]);
        } else {
#pragma unroll
            for (int j = 1; j <= K; ++j)
#pragma unroll
                for (int i = 1; i <= j; ++i) {
                    const double den = tw[K - 1 + i] - tw[K - 1 + i - j];
                    R[j - 1][i - 1] = (fabs(den) >= DBL_MIN) ? 1.0 / den : 0.0;
                }
        }
        const double x = tw[K - 1];
        double a[K1];
        // basis of degree j pairs with derivative level r = K-j: run the basis upwards and the
        // difference table downwards, meeting in the dot products
        double h[K1], lev[K1][K1];  // lev[r][e], e = r..K (fully unrolled: registers)
#pragma unroll
        for (int e = 0; e <= K; ++e) lev[0][e] = D[e];
#pragma unroll
        for (int r = 1; r <= K; ++r)
#pragma unroll
            for (int e = r; e <= K; ++e)
                lev[r][e] = (lev[r - 1][e] - lev[r - 1][e - 1]) * R[K - r][e - r];
        // x = t(l) is a knot: the last basis function of every degree >= 1 vanishes there, so
        // degree j has only j non-zero values h[0..j-1] and the i = j term of the recurrence drops
        h[0] = 1.0;
        a[K] = lev[K][K];
        double binom = 1.0;
#pragma unroll
        for (int j = 1; j <= K; ++j) {
            double hh[K];
#pragma unroll
            for (int i = 0; i < j; ++i) hh[i] = h[i];
            h[0] = 0.0;
#pragma unroll
            for (int i = 1; i <= ((j == 1) ? 1 : j - 1); ++i) {
                const double f = hh[i - 1] * R[j - 1][i - 1];
                h[i - 1] = fma(f, tw[K - 1 + i] - x, h[i - 1]);
                if (i < j) h[i] = f * (x - tw[K - 1 + i - j]);
            }
            const int r = K - j;
            double acc = 0.0;
#pragma unroll
            for (int q = 0; q < j; ++q) acc = fma(lev[r][r + q], h[q], acc);
            a[r] = acc;
        }
        // binom(K,r) for r = K-1 .. 0
#pragma unroll
        for (int r = K - 1; r >= 1; --r) {
            binom = binom * (double)(r + 1) / (double)(K - r);
            a[r] *= binom;
        }
        double2 *rl = reinterpret_cast<double2 *>(rec + (size_t)li * kRecD);
        double v[8];
        v[0] = x;
#pragma unroll
        for (int r = 0; r <= K; ++r) v[1 + r] = a[r];
#pragma unroll
        for (int r = K + 2; r < 8; ++r) v[r] = 0.0;
#pragma unroll
        for (int q = 0; q < (K + 3) / 2; ++q) rl[q] = make_double2(v[2 * q], v[2 * q + 1]);
    }
    __syncwarp();
}


// ---------------------------------------------------------------------------------------------
// shared pieces of the two kernels
// ---------------------------------------------------------------------------------------------

// out-of-support policy (core:17864-17877): returns true when the point's value is forced to 0
__device__ __forceinline__ bool apply_policy(int e, double tb, double te, double &arg, bool live,
                                             bool &out_of_range)
{
    if (arg < tb || arg > te) {
        if (e == 1) return true;
        if (e == 2) out_of_range = out_of_range || live;
        else if (e == 3) arg = fmax(fmin(arg, te), tb);
    }
    return false;
}

// Lookup-grid construction over [t(k+1), t(n-k)].  Each lane owns the knot intervals l = k1+lane,
// k1+lane+32, ... and fills the cells whose (widened) left edge falls into [t(l), t(l+1)): cell
// boundaries come from one monotone function of the knot, so the ranges of neighbouring lanes
// tile the grid.  A cell that may hold two or more knots is flagged (the single predicated step
// of the evaluation cannot resolve it: the reference's walk is taken).  Edges are widened by
// 1e-6 of a cell so that rounding in the cell index can never un-flag a cell.
template <int K, bool FASTCELLS>
__device__ __forceinline__ void build_grid(const double *ts, int n, void *grid, int &gcells,
                                           double &inv_h, int lane)
{
    constexpr int K1 = K + 1;
    const int nk1 = n - K1;
    const int nint = nk1 - K1 + 1;
    gcells = splev_grid_cells(nint);
    const double tb = ts[K1 - 1], te = ts[nk1];
    const double span = te - tb;
    inv_h = (span > 0.0) ? (double)gcells / span : 0.0;
    const double hcell = span / (double)gcells;
    auto bnd = [&](int l) -> int {  // first cell whose left edge is >= t(l)
        if (l <= K1) return 0;
        if (l > nk1) return gcells;
        const double v = (ts[l - 1] - tb) * inv_h + 1e-6;
        const int c = __double2int_ru(v);
        return max(0, min(c, gcells));
    };
    for (int li = lane; li < nint; li += 32) {
        const int l = K1 + li;
        const int g0 = bnd(l), g1 = bnd(l + 1);
        const double t2 = (l + 2 <= nk1) ? ts[l + 1] : 1.79e308;  // t(l+2), if stepping twice is possible
        const double t1 = (l < nk1) ? ts[l] : __longlong_as_double(0x7ff0000000000000LL);
        for (int gI = g0; gI < g1; ++gI) {
            const double right = tb + ((double)(gI + 1) + 1e-6) * hcell;
            const bool flag = right >= t2;
            if (FASTCELLS) {
                reinterpret_cast<double2 *>(grid)[gI] = make_double2(
                    t1, __hiloint2double(l | (flag ? (int)0x80000000 : 0), li * kRecD * 8));
            } else {
                reinterpret_cast<unsigned short *>(grid)[gI] = (unsigned short)(l | (flag ? (1 << 15) : 0));
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// EXACT mode
// ---------------------------------------------------------------------------------------------
// PLAIN: e == OUTSIDE_EXTRAPOLATE and no interval-index output -- no per-point policy code.
template <int K, bool PLAIN>
__global__ void __launch_bounds__(kSplevWarps * 32, 2) splev_kernel(SplevParams p, int smem_ld)
{
    constexpr int K1 = K + 1;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    const int per_warp = splev_smem_doubles(0, K, smem_ld, p.tab);
    uint64_t *mbars = reinterpret_cast<uint64_t *>(smem_raw);
    double *base = reinterpret_cast<double *>(smem_raw + 8 * kMbarPerWarp * kSplevWarps) + (size_t)warp * per_warp;
    auto tsb = [&](int b) { return base + (size_t)(2 * b) * smem_ld; };  // (t, c) buffer b
    auto csb = [&](int b) { return base + (size_t)(2 * b + 1) * smem_ld; };
    unsigned short *grid = reinterpret_cast<unsigned short *>(base + 4 * smem_ld);
    double *spans = base + 4 * smem_ld + (splev_grid_cells(smem_ld - 2 * K - 1) + 3) / 4;
    unsigned *spmask = reinterpret_cast<unsigned *>(spans + (size_t)(K * (K + 1) / 2) * p.tab);
    bool use_tab = false;   // per spline: its intervals fit the table
    double *ts = tsb(0), *cs = csb(0);
    int tbuf = 0;           // table buffer holding the current spline
    long long staged = -1;  // spline whose tables were prefetched into buffer 1-tbuf (TMA in flight)
    uint64_t *mbar = mbars + warp * kMbarPerWarp;
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
            {  // start fetching the next spline's tables into the other buffer
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
            build_grid<K, false>(ts, n, grid, gcells, inv_h, lane);
            use_tab = p.tab > 0 && (n - 2 * K - 1) <= p.tab;
            if (use_tab) build_spans<K>(ts, n, spans, spmask, lane);
            __syncwarp();
            cached = (int)tj;
        }
        const int nk1 = n - K1;
        const double tb = ts[K1 - 1], te = ts[nk1];
        double *yg = p.y + j * p.m;
        int *lg = p.lidx ? p.lidx + j * p.m : nullptr;
        bool out_of_range = false;
        // software pipeline in registers: the loads of the next group are in flight while the
        // current group is evaluated (the kernel is FP64-bound, not latency-bound)
        double nxt[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            const int i = lane + 32 * u;
            nxt[u] = (i < p.m) ? xg[i] : tb;
        }
        for (int i0 = lane; i0 < p.m; i0 += 32 * kUnroll) {
            double arg[kUnroll];
            bool live[kUnroll], zero[kUnroll];
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) {
                const int i = i0 + 32 * u;
                live[u] = i < p.m;
                arg[u] = nxt[u];
                const int inx = i + 32 * kUnroll;
                nxt[u] = (inx < p.m) ? xg[inx] : tb;
                zero[u] = PLAIN ? false : apply_policy(p.e, tb, te, arg[u], live[u], out_of_range);
            }
            int l[kUnroll];
            double sp[kUnroll];
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) l[u] = locate(ts, grid, gcells, tb, inv_h, K1, nk1, arg[u]);
#pragma unroll
            for (int u = 0; u < kUnroll; ++u)
                sp[u] = use_tab ? eval_exact_tab<K>(ts, cs, spans, spmask, l[u], arg[u]) : eval_exact<K>(ts, cs, l[u], arg[u]);
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) {
                const int i = i0 + 32 * u;
                if (live[u]) {
                    yg[i] = zero[u] ? 0.0 : sp[u];
                    if (!PLAIN && lg) lg[i] = zero[u] ? 0 : l[u];
                }
            }
        }
        if (p.ier) {
            const bool any_out = __any_sync(FULL, out_of_range);
            if (lane == 0) p.ier[j] = any_out ? IER_RANGE : IER_OK;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// FAST mode
// ---------------------------------------------------------------------------------------------
// Evaluate one group of kUnroll points per lane (indices ibase + 32*u) from the interval records
// and store the results.  The lookup cell gives the interval l0 of the cell's left edge, the next
// knot and the byte offset of l0's record; one compare steps over a knot inside the cell.  Cells
// that may hold two or more knots are flagged and take the reference's own walk
// (core:17880-17887).  (arg - tb) * inv_h is computed exactly as in the grid construction, with
// a relative error ~1e-16 << the 1e-6 cell widening used there.
template <int K, bool PLAIN, bool ALLLIVE>
__device__ __forceinline__ void fast_group(const SplevParams &p, const double *ts, const double *rec,
                                           const double2 *cells, int gcells, double tb, double te,
                                           double inv_h, int nk1, double (&arg)[kUnroll],
                                           const bool (&live)[kUnroll], double *yg, int *lg, int ibase,
                                           bool &out_of_range)
{
    constexpr int K1 = K + 1;
    bool zero[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u)
        zero[u] = PLAIN ? false : apply_policy(p.e, tb, te, arg[u], live[u], out_of_range);
    int off[kUnroll], lf[kUnroll];
    int anyflag = 0;
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
        int g = __double2int_rz((arg[u] - tb) * inv_h);
        g = max(0, min(g, gcells - 1));
        const double2 cv = cells[g];
        const bool step = arg[u] >= cv.x;
        off[u] = __double2loint(cv.y) + (step ? kRecD * 8 : 0);
        lf[u] = __double2hiint(cv.y) + (step ? 1 : 0);
        anyflag |= lf[u];
    }
    if (__any_sync(FULL, anyflag < 0)) {
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            if (lf[u] < 0) {
                int ll = lf[u] & 0x7fffffff;
                while (ll > K1 && arg[u] < ts[ll - 1]) --ll;
                while (ll < nk1 && arg[u] >= ts[ll]) ++ll;
                lf[u] = ll;
                off[u] = (ll - K1) * (kRecD * 8);
            }
        }
    }
    double sp[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
        const double2 *rp = reinterpret_cast<const double2 *>(reinterpret_cast<const char *>(rec) + off[u]);
        double a[8];
#pragma unroll
        for (int q = 0; q < (K + 3) / 2; ++q) {
            const double2 v = rp[q];
            a[2 * q] = v.x;
            a[2 * q + 1] = v.y;
        }
        const double uu = arg[u] - a[0];
        double v = a[1 + K];
#pragma unroll
        for (int r = K - 1; r >= 0; --r) v = fma(v, uu, a[1 + r]);
        sp[u] = v;
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
        const int i = ibase + 32 * u;
        if (ALLLIVE || live[u]) {
            if (PLAIN) {
                yg[(size_t)i * p.y_stride] = sp[u];
            } else {
                yg[(size_t)i * p.y_stride] = zero[u] ? 0.0 : sp[u];
                if (lg) lg[i] = zero[u] ? 0 : lf[u];
            }
        }
    }
}

template <int K, bool PLAIN>
__global__ void __launch_bounds__(kFastWarps * 32, kFastBlocks) splev_fast_kernel(SplevParams p, int smem_ld)
{
    constexpr int K1 = K + 1;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    const int per_warp = splev_smem_doubles(1, K, smem_ld);
    uint64_t *mbars = reinterpret_cast<uint64_t *>(smem_raw);
    double *base = reinterpret_cast<double *>(smem_raw + 8 * kMbarPerWarp * kSplevWarps) + (size_t)warp * per_warp;
    auto tsb = [&](int b) { return base + (size_t)(2 * b) * smem_ld; };  // (t, c) buffer b
    auto csb = [&](int b) { return base + (size_t)(2 * b + 1) * smem_ld; };
    double *rec = base + 4 * smem_ld;  // interval records
    const int nintmax = (smem_ld - 2 * K - 1 > 1) ? smem_ld - 2 * K - 1 : 1;
    double2 *cells = reinterpret_cast<double2 *>(rec + ((kRecD * nintmax + 1) & ~1));
    double *xs = reinterpret_cast<double *>(cells + splev_grid_cells(smem_ld - 2 * K - 1));  // [kNBuf][kTilePts]
    double *ts = tsb(0), *cs = csb(0);
    int tbuf = 0;           // table buffer holding the current spline
    long long staged = -1;  // spline whose tables were prefetched into buffer 1-tbuf (TMA in flight)
    uint64_t *mbar = mbars + warp * kMbarPerWarp;  // [0] tables, [1 + slot] x tiles
    if (lane == 0) {
#pragma unroll
        for (int b = 0; b <= kNBuf; ++b) mbar_init(mbar + b, 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncwarp();
    uint32_t phase = 0;

    const long long gwarp = (long long)blockIdx.x * wpb + warp;
    const long long nwarps = (long long)gridDim.x * wpb;

    // ---- the point stream: tiles of kTilePts points, spline after spline, kNBuf tiles ahead.
    // TMA needs 16-byte aligned rows of even length (uniform for the launch); otherwise the
    // points are loaded straight into registers.
    const bool xt_ok = ((((uintptr_t)p.x) & 15) == 0) && ((p.m & 1) == 0);
    const int ntile = (p.m + kTilePts - 1) / kTilePts;
    // producer cursor: where the next tile starts, how many points of that spline are left, and
    // how many splines of this warp's sequence have not been requested completely
    const double *psrc = p.x + gwarp * p.m;
    int prem = p.m;
    long long pleft = (gwarp < p.nspl) ? (p.nspl - gwarp + nwarps - 1) / nwarps : 0;
    const long long pjump = (nwarps - 1) * (long long)p.m;
    int slot = 0;          // ring slot of the next tile to consume == next slot to refill
    uint32_t xphase = 0;   // phase bit per slot
    auto produce = [&]() { // request the next tile of the stream into `slot`, then advance the ring
        if (pleft > 0) {
            const int cnt = min(kTilePts, prem);
            if (lane == 0) {
                uint64_t *mb = mbar + 1 + slot;
                asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
                mbar_expect_tx(mb, (uint32_t)cnt * 8u);
                tma_load_1d(xs + slot * kTilePts, psrc, (uint32_t)cnt * 8u, mb);
            }
            psrc += cnt;
            prem -= cnt;
            if (prem == 0) {
                psrc += pjump;
                prem = p.m;
                --pleft;
            }
        }
        slot = (slot + 1) & (kNBuf - 1);
    };
    auto consume_wait = [&]() {
        mbar_wait(mbar + 1 + slot, (xphase >> slot) & 1u);
        xphase ^= 1u << slot;
    };
    if (xt_ok) {
#pragma unroll
        for (int b = 0; b < kNBuf; ++b) produce();
    }

    const long long ldc = p.ldc ? p.ldc : p.ldt;
    auto coef = [&](long long q, int nq) { return p.c + q * ldc + (long long)p.c_mul * nq; };
    auto tables_tma_ok = [&](long long q, int nq) -> bool {
        const double *tg = p.t + q * p.ldt, *cg = coef(q, nq);
        return ((((uintptr_t)tg) | ((uintptr_t)cg)) & 15) == 0 && (((nq + 1) & ~1) <= p.ldt) &&
               ((long long)p.c_mul * nq + ((nq + 1) & ~1) <= ldc);
    };
    auto issue_tables = [&](long long q, int nq, int buf) {
        if (lane == 0) {
            const uint32_t bytes = (uint32_t)(((nq + 1) & ~1) * 8);
            asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
            mbar_expect_tx(mbar, 2 * bytes);
            tma_load_1d(tsb(buf), p.t + q * p.ldt, bytes, mbar);
            tma_load_1d(csb(buf), coef(q, nq), bytes, mbar);
        }
    };

    int cached = -1;
    int n = 0, gcells = 32;
    double inv_h = 0.0;
    // knot counts are read one spline ahead so that their DRAM latency is never on the critical path
    int n_cur = (gwarp < p.nspl) ? p.n[p.nshared ? 0 : gwarp] : 0;
    int n_nxt = (!p.nshared && gwarp + nwarps < p.nspl) ? p.n[gwarp + nwarps] : 0;
    for (long long j = gwarp; j < p.nspl; j += nwarps) {
        const long long tj = p.nshared ? 0 : j;
        const int nj = n_cur;
        const int n_fut = (!p.nshared && j + 2 * nwarps < p.nspl) ? p.n[j + 2 * nwarps] : 0;
        if (nj < 2 * K1 || nj > p.ldt) {
            // a knot count the row cannot hold, or too small for degree K, is invalid input
            // (splev itself would read out of bounds): flag the spline, drain its tiles, skip it
            if (p.ier && lane == 0) p.ier[j] = IER_INPUT;
            if (xt_ok) {
                for (int t = 0; t < ntile; ++t) {
                    consume_wait();
                    __syncwarp();
                    produce();
                }
            }
            if (!p.nshared) { n_cur = n_nxt; n_nxt = n_fut; }
            continue;
        }
        if (tj != cached) {
            n = nj;
            __syncwarp();  // previous spline's readers are done with their tables
            if (staged == tj) {  // prefetched while the previous spline was evaluated
                tbuf ^= 1;
                mbar_wait(mbar, phase);
                phase ^= 1;
            } else if (tables_tma_ok(tj, n)) {
                issue_tables(tj, n, tbuf);
                mbar_wait(mbar, phase);
                phase ^= 1;
            } else {
                const double *tg = p.t + tj * p.ldt, *cg = coef(tj, n);
                for (int i = lane; i < n; i += 32) {
                    tsb(tbuf)[i] = tg[i];
                    csb(tbuf)[i] = (i < n - K1) ? cg[i] : 0.0;   // only the n-k-1 coefficients are addressable
                }
            }
            ts = tsb(tbuf);
            cs = csb(tbuf);
            staged = -1;
            {  // start fetching the next spline's tables into the other buffer
                const long long jn = j + nwarps;
                if (!p.nshared && jn < p.nspl && n_nxt >= 2 * K1 && n_nxt <= p.ldt && tables_tma_ok(jn, n_nxt)) {
                    issue_tables(jn, n_nxt, tbuf ^ 1);
                    staged = jn;
                }
            }
            __syncwarp();
            build_grid<K, true>(ts, n, cells, gcells, inv_h, lane);
            build_records<K>(ts, cs, n, rec, lane);
            __syncwarp();
            cached = (int)tj;
        }
        const int nk1 = n - K1;
        const double tb = ts[K1 - 1], te = ts[nk1];
        double *yg = p.y + (j * p.m) * p.y_stride + p.y_off;
        int *lg = p.lidx ? p.lidx + j * p.m : nullptr;
        bool out_of_range = false;
        if (xt_ok) {
            for (int t = 0; t < ntile; ++t) {
                consume_wait();
                const double *xt = xs + slot * kTilePts;
                const int cnt = p.m - t * kTilePts;  // points left, this tile included
#pragma unroll
                for (int gq = 0; gq < kTileGroups; ++gq) {
                    const int gcnt = cnt - gq * kGroupPts;  // points left from this group on
                    double arg[kUnroll];
                    bool live[kUnroll];
#pragma unroll
                    for (int u = 0; u < kUnroll; ++u) {
                        live[u] = lane + 32 * u < gcnt;
                        arg[u] = xt[gq * kGroupPts + lane + 32 * u];  // stale beyond cnt: replaced below
                    }
                    if (gq == kTileGroups - 1) {
                        __syncwarp();  // every lane has read the tile: hand the slot back to the stream
                        produce();
                    }
                    const int ibase = t * kTilePts + gq * kGroupPts + lane;
                    if (gcnt >= kGroupPts) {
                        fast_group<K, PLAIN, true>(p, ts, rec, cells, gcells, tb, te, inv_h, nk1, arg, live,
                                                   yg, lg, ibase, out_of_range);
                    } else if (gcnt > 0) {
#pragma unroll
                        for (int u = 0; u < kUnroll; ++u)
                            if (!live[u]) arg[u] = tb;
                        fast_group<K, PLAIN, false>(p, ts, rec, cells, gcells, tb, te, inv_h, nk1, arg, live,
                                                    yg, lg, ibase, out_of_range);
                    }
                }
            }
        } else {
            // unaligned rows: software pipeline in registers, one group ahead
            const double *xg = p.x + j * p.m;
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
                fast_group<K, PLAIN, false>(p, ts, rec, cells, gcells, tb, te, inv_h, nk1, arg, live, yg, lg,
                                            i0, out_of_range);
            }
        }
        if (p.ier) {
            const bool any_out = __any_sync(FULL, out_of_range);
            if (lane == 0) p.ier[j] = any_out ? IER_RANGE : IER_OK;
        }
        if (!p.nshared) { n_cur = n_nxt; n_nxt = n_fut; }
    }
}

}  // namespace

cudaError_t launch_splev(const SplevParams &p_in, int nsm, cudaStream_t stream)
{
    SplevParams p = p_in;
    if (p.nspl <= 0) return cudaSuccess;
    const int smem_ld = (p.ldt + 1) & ~1;  // keep every per-warp region 16 B aligned
    const size_t header = 8 * kMbarPerWarp * kSplevWarps;
    // EXACT: a span table for up to 96 knot intervals per spline (6 to 12 KB per warp at k = 3..5, small
    // next to the knot/coefficient staging so that the resident warps stay); longer splines evaluate fpbspl
    // with plain divisions
    static const bool no_tab = std::getenv("FPB_SPLEV_NO_TAB") != nullptr;
    {
        const int nintmax = (smem_ld - 2 * p.k - 1 > 1) ? smem_ld - 2 * p.k - 1 : 1;
        p.tab = (p.mode == 0 && !no_tab) ? (nintmax < 96 ? nintmax : 96) : 0;
    }
    const int reg_blocks = (p.mode == 1) ? kFastBlocks : 2;  // CTAs/SM the register budget allows
    // keep reg_blocks CTAs resident if the tables allow it (227 KB per SM, 1 KB reserved per CTA)
    auto warps_for = [&](int tab) {
        const size_t pw = (size_t)splev_smem_doubles(p.mode, p.k, smem_ld, tab) * sizeof(double);
        int w = (p.mode == 1) ? kFastWarps : kSplevWarps;
        while (w > 1 && header + pw * w > (size_t)(227 * 1024) / reg_blocks - 1024) w >>= 1;
        return w;
    };
    if (p.tab && warps_for(p.tab) != warps_for(0)) p.tab = 0;   // never trade resident warps for the table
    const size_t per_warp = (size_t)splev_smem_doubles(p.mode, p.k, smem_ld, p.tab) * sizeof(double);
    int wpb = warps_for(p.tab);
    const size_t smem = header + per_warp * wpb;
    if (smem > 220 * 1024) return cudaErrorInvalidValue;  // knot vectors beyond ~4000 entries
    long long blocks = ((long long)p.nspl + wpb - 1) / wpb;
    const long long resident = (long long)((228 * 1024) / (smem + 1024));
    const long long max_blocks =
        (long long)nsm * (resident < 1 ? 1 : (resident > reg_blocks ? reg_blocks : resident));
    if (blocks > max_blocks) blocks = max_blocks;
#define FPB_LAUNCH_SPLEV(KK)                                                                       \
    do {                                                                                           \
        const bool plain = (p.e == 0 && p.lidx == nullptr);                                       \
        auto kern = (p.mode == 1) ? (plain ? splev_fast_kernel<KK, true> : splev_fast_kernel<KK, false>) \
                                  : (plain ? splev_kernel<KK, true> : splev_kernel<KK, false>);   \
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
