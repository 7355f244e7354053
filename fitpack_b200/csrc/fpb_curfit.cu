// fpb_curfit.cu -- batched smoothing-spline fit (curfit/fpcurf) for sm_100a.
//
// One thread per curve, 32 curves per warp, persistent warps pulling 32-curve
// tiles from an atomic queue.  The knot loop of fpcurf is cut at pass granularity:
// `curfit_pass_kernel` runs ONE least-squares pass (+ acceptance tests, residual
// split, knot insertion) for every curve in its input list and re-buckets the
// curves into "needs another pass" / "knots accepted: part 2" / done, so every
// lane of every warp always has a pass to do (the monolithic version lost half
// of its issue slots to lanes whose curve needed fewer knot sets, profiles/r01);
// `curfit_part2_kernel` then runs the smoothing-parameter iteration on the
// part-2 list sorted by knot count.  Follows, per curve, exactly the control flow
// and IEEE operation order of the reference:
//   curfit  src/fitpack_core.F90:1240-1299   (argument checks, fpchec)
//   fpcurf  src/fitpack_core.F90:4737-5087   (knot loop, LSQ, p-iteration)
//   fpbspl 2387-2413, fp_knot_interval 2432-2473, fp_rotate_row 5691-5717,
//   fp_rotate_shifted 5800-5820, fpgivs 5638-5658, fprota 12385-12401,
//   fpback 1808-1838, fpdisc 5453-5495, fpknot 8199-8275,
//   root_finding_iterate 11641-11691, fprati 11996-12025, fpchec 2507-2570.
//
// B200 mapping:
//  * inputs are read point-major (SoA): lane b of a warp reads x[i*ld+b], so
//    every pass over the m points is a stream of fully coalesced 256 B loads;
//  * the band matrix R is never materialised per point: a point in knot
//    interval l only touches R rows l-k..l, and x is sorted, so the k+1 active
//    rows (+ their right-hand sides) live in a per-thread circular window in SHARED
//    memory ([element][thread], conflict free; registers hold the 2k knots fpbspl needs
//    and the refined reciprocals of the knot spans); a row is written out exactly once
//    per pass, when the window moves past it;
//  * the basis values q(m,k+1) the reference stores in wrk are NOT stored: the residual
//    passes recompute them, with fpbspl's divisions reduced to the quotient step of the
//    hardware's own division sequence (bspl_tab, fpb_common.cuh) -- same bits, and two
//    thirds of the round-1 pipeline's DRAM traffic gone;
//  * fpgivs is one straight-line block of fast-path division / square-root sequences
//    (givens_fast) whenever its operands are in range;
//  * all per-knot state (t, c, z, R, fpint, nrdata, G, B) lives in a per-warp
//    workspace interleaved by lane ([element][lane]), so loops whose index is
//    uniform across the warp (back-substitution, p-iteration, fpknot, fpdisc)
//    are coalesced too; trip counts are made warp-uniform with a max-reduce and
//    per-lane predicates;
//  * the pass that accepts a curve's knot set hands its triangular system (R, z) to part 2
//    through a pool, so part 2 does not regenerate it;
//  * all decisions (knot insertion, the rational p-search, every exit) are taken
//    on the device; the host only reads the list lengths (every other pass, every pass
//    near the tail) to size the next launch and to start part 2 of the early finishers
//    beside the tail passes.  Per-curve ier is an output, a bad curve never aborts the batch.
//
// Parametric curves (parcur/fppara, core:15201-15285 / 8822-9177) run through the SAME pipeline:
// the kernels are templated on IDIM, the number of coordinate curves that share one knot vector
// and one band matrix (fp_rotate_row_vec core:5742-5771, fp_rotate_shifted_vec core:5845-5862:
// one Givens rotation, idim right-hand sides).  IDIM = 1, 2, 3 are compiled as separate
// translation units (fpb_curfit_d2.cu ... include this file with FPB_FIT_IDIM set); IDIM = 0 is
// the run-time fallback for idim = 4..10.  p.par selects parcur's argument checks and the few
// places where fppara's arithmetic differs from fpcurf's.
#include <algorithm>

#include "fpb_common.cuh"
#include "fpb_kernels.h"

#ifndef FPB_FIT_IDIM
#define FPB_FIT_IDIM 1
#endif

namespace fpb {

namespace {

enum : int { ST_LOOP = 0, ST_PART2 = 1, ST_DONE = 2 };

constexpr unsigned FULL = 0xffffffffu;

#define WS(off, i) wsd[(size_t)((off) + (i)) * 32]
// 1-based accessors in the reference's notation
#define T_(i) WS(p.off_t, (i) - 1)
#define C_(i) WS(p.off_c, (i) - 1)
#define Z_(i) WS(p.off_z, (i) - 1)
#define CD_(d, i) WS(p.off_c + (d) * p.nestp, (i) - 1) /* coordinate d */
#define ZD_(d, i) WS(p.off_z + (d) * p.nestp, (i) - 1)
#define FPI_(i) WS(p.off_fpint, (i) - 1)
#define A_(i, j) WS(p.off_a, ((i) - 1) * K1 + ((j) - 1))
#define G_(i, j) WS(p.off_g, ((i) - 1) * K2 + ((j) - 1))
#define B_(i, j) WS(p.off_b, ((i) - 1) * K2 + ((j) - 1))
#define Q_(it, j) WS(p.off_q, (it) * K1 + (j))  /* it 0-based, j 0-based */
#define NRD_(i) wsi[(size_t)((i) - 1) * 32]

// Build switches (A/B tuning, profiles/scripts/ab_fit.py): the defaults are the shipped configuration.
//  FPB_FIT_STOREQ = 1: the least-squares pass stores the basis values q(m,k+1) in the workspace and the
//  residual passes read them back (what the reference does, core:1295 / 4886); 0: the residual passes
//  recompute them with bspl_tab -- bit-identical, and 2/3 of the pipeline's DRAM traffic disappears.
#ifndef FPB_FIT_STOREQ
#define FPB_FIT_STOREQ 0
#endif
//  FPB_FIT_SPAN_SMEM = 1: the span reciprocals of bspl_tab live in shared memory (after the stream ring)
//  instead of registers.
#if FPB_FIT_SPAN_SMEM
#define FPB_SPAN_STRIDE kFitThreads
#else
#define FPB_SPAN_STRIDE 0
#endif

//  FPB_FIT_STREAM = 1: point streams through a cp.async ring in shared memory; 0: plain loads, one point of
//  register lookahead.   FPB_FIT_FASTDIV = 1: fpbspl through the span-reciprocal table (bspl_tab).
#ifndef FPB_FIT_FASTDIV
#define FPB_FIT_FASTDIV 1
#endif
//  FPB_FIT_GIVENS: 2 = givens_fast (one straight-line block of fast-path division / sqrt sequences),
//  1 = givens_shared (cos and sin share one reciprocal refinement), 0 = givens (plain IEEE operators).
#ifndef FPB_FIT_GIVENS
#define FPB_FIT_GIVENS 2
#endif
#if FPB_FIT_GIVENS == 2
#define FPB_GIVENS givens_fast
#elif FPB_FIT_GIVENS == 1
#define FPB_GIVENS givens_shared
#else
#define FPB_GIVENS givens
#endif
//  FPB_FIT_PFDIST: the plain-load point streams ask L2 for the rows they will read this many points later.
#ifndef FPB_FIT_PFDIST
#define FPB_FIT_PFDIST 0   // measured: per-lane L2 prefetches cost more than they hide (profiles/r02_ab.md)
#endif

// ---- point streams ---------------------------------------------------------------------------------
// x(i), y(i,:), w(i) of the thread's curve.  ASYNC: they flow through a ring of `depth` slots in shared
// memory filled by 8-byte cp.async copies issued `depth` points ahead, so the DRAM/L2 latency of the stream
// is covered without holding the values in registers.  Ring element e of slot s sits at
// ring[(s*E + e) * NT], E = 1 + nd + HAS_W doubles per point.  !ASYNC: plain loads, the next point is
// requested while the current one is processed.
template <int ND, bool HAS_W, bool ASYNC>
struct PointStream {
    double *ring;
    const double *xp, *yp, *wp;
    size_t x_si, y_si, y_sd, w_si;
    int nd, m, depth, E;
    double xn, yn[ND], wn;   // !ASYNC: the prefetched point
    __device__ __forceinline__ void issue(int it) const
    {
        const int ii = min(it, m - 1);
        double *slot = ring + (size_t)((it & (depth - 1)) * E) * kFitThreads;
        cp_async8(slot, xp + (size_t)ii * x_si);
#pragma unroll
        for (int d = 0; d < ND; ++d)
            if (d < nd) cp_async8(slot + (size_t)(1 + d) * kFitThreads, yp + (size_t)ii * y_si + (size_t)d * y_sd);
        if (HAS_W) cp_async8(slot + (size_t)(1 + nd) * kFitThreads, wp + (size_t)ii * w_si);
        cp_async_commit();
    }
    __device__ __forceinline__ void load(int it)
    {
        const int ii = min(it, m - 1);
        xn = xp[(size_t)ii * x_si];
#pragma unroll
        for (int d = 0; d < ND; ++d)
            if (d < nd) yn[d] = yp[(size_t)ii * y_si + (size_t)d * y_sd];
        wn = HAS_W ? wp[(size_t)ii * w_si] : 1.0;
        if (FPB_FIT_PFDIST > 0) {
            const int ip = min(it + FPB_FIT_PFDIST, m - 1);
            prefetch_l2(xp + (size_t)ip * x_si);
#pragma unroll
            for (int d = 0; d < ND; ++d)
                if (d < nd) prefetch_l2(yp + (size_t)ip * y_si + (size_t)d * y_sd);
            if (HAS_W) prefetch_l2(wp + (size_t)ip * w_si);
        }
    }
    __device__ __forceinline__ void start()
    {
        if (ASYNC) {
            for (int d = 0; d < depth; ++d) issue(d);
        } else {
            load(0);
        }
    }
    // point `it` (called for it = 0, 1, 2, ... in order); requests a later point
    __device__ __forceinline__ void fetch(int it, double &x, double (&y)[ND], double &w)
    {
        if (ASYNC) {
            // point `it` has landed once at most depth-1 younger groups are pending
            if (depth == 8) cp_async_wait<7>();
            else if (depth == 4) cp_async_wait<3>();
            else if (depth == 2) cp_async_wait<1>();
            else cp_async_wait<0>();
            const double *sl = ring + (size_t)((it & (depth - 1)) * E) * kFitThreads;
            x = sl[0];
#pragma unroll
            for (int d = 0; d < ND; ++d)
                if (d < nd) y[d] = sl[(size_t)(1 + d) * kFitThreads];
            w = HAS_W ? sl[(size_t)(1 + nd) * kFitThreads] : 1.0;
            issue(it + depth);
        } else {
            x = xn;
#pragma unroll
            for (int d = 0; d < ND; ++d)
                if (d < nd) y[d] = yn[d];
            w = wn;
            load(it + 1);
        }
    }
    __device__ __forceinline__ void drain() const
    {
        if (ASYNC) cp_async_wait<0>();
    }
};

// interpolation knots, core:4806-4816 / 4975-4985
template <int K>
__device__ __forceinline__ void interp_knots(const FitParams &p, double *wsd, const double *xp, int m)
{
    constexpr int K1 = K + 1, K2 = K + 2;
    const int mk1 = m - K1, k3 = K / 2;
    int i = K2, j = k3 + 2;
    for (int l = 1; l <= mk1; ++l) {
        const double xj = xp[(size_t)(j - 1) * p.x_si];
        double v;
        if (k3 * 2 != K) v = xj;
        else v = mul(add(xj, xp[(size_t)(j - 2) * p.x_si]), 0.5);
        T_(i) = v;
        ++i;
        ++j;
    }
}

// fpback, core:1808-1838.  Band matrix rows come from `offm` (row stride BAND),
// right-hand side from `offz`, result into the c array.  The row loop runs over
// a warp-uniform range so that the loads coalesce; `on` masks the lane.
template <int BAND>
__device__ __forceinline__ void back_subst(const FitParams &p, double *wsd, int offm, int offz,
                                           int offc, int nk1, bool on)
{
    const int hi = __reduce_max_sync(FULL, on ? nk1 : 0);
    double cw[BAND - 1];
#pragma unroll
    for (int q = 0; q < BAND - 1; ++q) cw[q] = 0.0;
    for (int i = hi; i >= 1; --i) {
        if (on && i <= nk1) {
            double store = WS(offz, i - 1);
            const int cnt = min(nk1 - i, BAND - 1);
#pragma unroll
            for (int q = 1; q <= BAND - 1; ++q)
                if (q <= cnt) store = sub(store, mul(cw[q - 1], WS(offm, (i - 1) * BAND + q)));
            const double ci = store / WS(offm, (i - 1) * BAND);
            WS(offc, i - 1) = ci;
#pragma unroll
            for (int q = BAND - 2; q >= 1; --q) cw[q] = cw[q - 1];
            cw[0] = ci;
        }
    }
}

// sum of squared residuals of the current c over all points (get_fp loop, core:5061-5069 / fppara
// core:9147-9159); when FPINT is set also the per-interval split of core:4942-4966 (fppara
// core:9023-9052).  The reference multiplies c with the basis values q(it,:) the least-squares pass
// stored; here they are recomputed with bspl_tab from the same knots in the same interval (the walk of
// fp_knot_interval, `lq`) -- the same bits -- while the coefficient window follows the reference's own,
// lagging interval counter `l` of this loop (at most one step per point).
template <int K, bool FPINT, bool HAS_W, int IDIM>
__device__ __forceinline__ double residual_pass(const FitParams &p, double *wsd, double *ring,
                                                const double *xp, const double *yp, const double *wp,
                                                size_t x_si, size_t y_si, size_t y_sd, size_t w_si, int m,
                                                int nk1, int nrint, bool on)
{
    constexpr int K1 = K + 1, K2 = K + 2, ND = IDIM > 0 ? IDIM : kMaxIdim;
    const int nd = IDIM > 0 ? IDIM : p.idim;
    PointStream<ND, HAS_W, FPB_FIT_STREAM != 0> ps{ring, xp, yp, wp, x_si, y_si, y_sd, w_si, nd, m, p.sdepth,
                                                   1 + nd + (HAS_W ? 1 : 0)};
    double cw[ND][K1];
    double fpart = 0.0, tl = 0.0;
    int l = K2, i = 1;
#pragma unroll
    for (int d = 0; d < ND; ++d) {
#pragma unroll
        for (int j = 0; j < K1; ++j) cw[d][j] = (on && d < nd) ? CD_(d, j + 1) : 0.0;
    }
    if (on) tl = T_(l);
#if !FPB_FIT_STOREQ
    double tw[2 * K];
    int lq = K1;
#pragma unroll
    for (int j = 0; j < 2 * K; ++j) tw[j] = on ? T_(lq - K + 1 + j) : 0.0;
    double tnx = (on && lq + K + 1 <= nk1 + K1) ? T_(lq + K + 1) : 0.0;  // knot that enters at the next advance
#if FPB_FIT_FASTDIV
    SpanTab<K, FPB_SPAN_STRIDE> sp;
    sp.rs = ring + (size_t)(ps.depth * ps.E) * kFitThreads;
    span_init<K, FPB_SPAN_STRIDE>(sp, tw);
#endif
#endif
    ps.start();
    for (int it = 0; it < m; ++it) {
        double xv, yv[ND], wv;
        ps.fetch(it, xv, yv, wv);
        if (on) {
            double qv[K1];
#if FPB_FIT_STOREQ
#pragma unroll
            for (int j = 0; j < K1; ++j) qv[j] = Q_(it, j);
#else
            while (lq < nk1 && xv >= tw[K]) {
#pragma unroll
                for (int j = 0; j < 2 * K - 1; ++j) tw[j] = tw[j + 1];
                ++lq;
                tw[2 * K - 1] = tnx;
                if (lq + K + 1 <= nk1 + K1) tnx = T_(lq + K + 1);
#if FPB_FIT_FASTDIV
                span_advance<K, FPB_SPAN_STRIDE>(sp, tw);
#endif
            }
#if FPB_FIT_FASTDIV
            bspl_tab<K, FPB_SPAN_STRIDE>(tw, sp, xv, qv);
#else
            bspl<K>(tw, xv, qv);
#endif
#endif
            bool newk = false;
            if (xv >= tl && l <= nk1) {
                newk = true;
                ++l;
#pragma unroll
                for (int d = 0; d < ND; ++d) {
                    if (d < nd) {
#pragma unroll
                        for (int j = 0; j < K; ++j) cw[d][j] = cw[d][j + 1];
                        cw[d][K] = CD_(d, l - 1);
                    }
                }
                tl = T_(l);
            }
            // fpcurf: (w*(s(x)-y))**2; fppara part 1: sum_d (w*(s_d(u)-x_d))**2 (identical
            // for one coordinate); fppara part 2: (sum_d (s_d(u)-x_d)**2) * w**2
            const bool late_w = HAS_W && !FPINT && p.par;
            double term = 0.0;
#pragma unroll
            for (int d = 0; d < ND; ++d) {
                if (d < nd) {
                    double fac = 0.0;
#pragma unroll
                    for (int j = 0; j < K1; ++j) fac = add(fac, mul(cw[d][j], qv[j]));
                    fac = sub(fac, yv[d]);
                    if (HAS_W && !late_w) fac = mul(wv, fac);
                    fac = mul(fac, fac);
                    term = (d == 0) ? fac : add(term, fac);
                }
            }
            if (late_w) term = mul(term, mul(wv, wv));
            fpart = add(fpart, term);
            if (FPINT && newk) {
                const double store = mul(term, 0.5);
                FPI_(i) = sub(fpart, store);
                ++i;
                fpart = store;
            }
        }
    }
    ps.drain();
    if (FPINT && on) FPI_(nrint) = fpart;
    return fpart;
}

// one least-squares pass (coefs loop core:4873-4896); returns fp = sum of squared transformed
// right-hand sides.  A point in knot interval l only touches band rows l-k..l and x is sorted,
// so only k+1 rows of R (+ right-hand sides) are live at any time.  They sit in a per-thread
// circular window in SHARED memory (column `rs` of the CTA's buffer, conflict-free
// [element][thread] layout): that keeps the kernel at 96 registers, i.e. 20 warps per SM
// instead of 16 -- the pass is bound by dependent FP64 issue latency, so warps are throughput.
// A row is written to the workspace exactly once per pass, when the window leaves it.
template <int K, bool HAS_W, int IDIM>
__device__ __forceinline__ double lsq_pass(const FitParams &p, double *wsd, double *rs, double *ring,
                                           const double *xp, const double *yp, const double *wp,
                                           size_t x_si, size_t y_si, size_t y_sd, size_t w_si, int m,
                                           int nk1, bool on)
{
    constexpr int K1 = K + 1, NT = kFitThreads, ND = IDIM > 0 ? IDIM : kMaxIdim;
    const int nd = IDIM > 0 ? IDIM : p.idim;
    PointStream<ND, HAS_W, FPB_FIT_STREAM != 0> ps{ring, xp, yp, wp, x_si, y_si, y_sd, w_si, nd, m, p.sdepth,
                                                   1 + nd + (HAS_W ? 1 : 0)};
#define RS_(slot, q) rs[((slot) * K1 + (q)) * NT]
#define ZS_(slot, d) rs[(K1 * K1 + (slot) * nd + (d)) * NT]
    double tw[2 * K], h[K1];
#pragma unroll
    for (int r = 0; r < K1; ++r) {
#pragma unroll
        for (int d = 0; d < ND; ++d)
            if (d < nd) ZS_(r, d) = 0.0;
#pragma unroll
        for (int q = 0; q < K1; ++q) RS_(r, q) = 0.0;
    }
    int l = K1, s0 = 0;  // s0: window slot holding band row l-K1+1
#pragma unroll
    for (int j = 0; j < 2 * K; ++j) tw[j] = on ? T_(l - K + 1 + j) : 0.0;
    double tnx = (on && l + K + 1 <= nk1 + K1) ? T_(l + K + 1) : 0.0;  // knot that enters at the next advance
#if FPB_FIT_FASTDIV
    SpanTab<K, FPB_SPAN_STRIDE> sp;
    sp.rs = ring + (size_t)(ps.depth * ps.E) * NT;
    span_init<K, FPB_SPAN_STRIDE>(sp, tw);
#endif
    double fp = 0.0;
    ps.start();
    for (int it = 0; it < m; ++it) {
        double xi, yi[ND], wi;
        ps.fetch(it, xi, yi, wi);
        // w == NULL means unit weights: y*1 and h*1 are exact, so the products are skipped
        if (HAS_W) {
#pragma unroll
            for (int d = 0; d < ND; ++d)
                if (d < nd) yi[d] = mul(yi[d], wi);
        }
        if (on) {
            // fp_knot_interval (core:2432-2473) == linear walk from the previous interval
            while (l < nk1 && xi >= tw[K]) {
                // window leaves row l-K1+1 (1-based): it is final for this pass; its slot is
                // recycled (zeroed) for the row that enters at the other end
                const int row = l - K1 + 1;
#pragma unroll
                for (int q = 0; q < K1; ++q) {
                    A_(row, q + 1) = RS_(s0, q);
                    RS_(s0, q) = 0.0;
                }
#pragma unroll
                for (int d = 0; d < ND; ++d) {
                    if (d < nd) {
                        ZD_(d, row) = ZS_(s0, d);
                        ZS_(s0, d) = 0.0;
                    }
                }
                s0 = (s0 + 1 == K1) ? 0 : s0 + 1;
#pragma unroll
                for (int j = 0; j < 2 * K - 1; ++j) tw[j] = tw[j + 1];
                ++l;
                tw[2 * K - 1] = tnx;
                if (l + K + 1 <= nk1 + K1) tnx = T_(l + K + 1);
#if FPB_FIT_FASTDIV
                span_advance<K, FPB_SPAN_STRIDE>(sp, tw);
#endif
            }
#if FPB_FIT_FASTDIV
            bspl_tab<K, FPB_SPAN_STRIDE>(tw, sp, xi, h);
#else
            bspl<K>(tw, xi, h);
#endif
#pragma unroll
            for (int j = 0; j < K1; ++j) {
#if FPB_FIT_STOREQ
                Q_(it, j) = h[j];
#endif
                if (HAS_W) h[j] = mul(wi, h[j]);
            }
            // fp_rotate_row / fp_rotate_row_vec (core:5691-5717 / 5742-5771): pivot i <-> band row
            // l-K1+1+i; one rotation serves all coordinates
            int sl_ = s0;
#pragma unroll
            for (int i = 0; i < K1; ++i) {
                const double piv = h[i];
                if (!fp_is_zero(piv)) {
                    double cs, sn, dg = RS_(sl_, 0);
                    FPB_GIVENS(piv, dg, cs, sn);
                    RS_(sl_, 0) = dg;
#pragma unroll
                    for (int d = 0; d < ND; ++d) {
                        if (d < nd) {
                            double z = ZS_(sl_, d);
                            rota(cs, sn, yi[d], z);
                            ZS_(sl_, d) = z;
                        }
                    }
#pragma unroll
                    for (int q = 1; q <= K - i; ++q) {
                        double r = RS_(sl_, q);
                        rota(cs, sn, h[i + q], r);
                        RS_(sl_, q) = r;
                    }
                }
                sl_ = (sl_ + 1 == K1) ? 0 : sl_ + 1;
            }
            // fp = fp + yi**2 (core:4895); fppara: fp + sum(xi**2) (core:8972)
            double sq = mul(yi[0], yi[0]);
#pragma unroll
            for (int d = 1; d < ND; ++d)
                if (d < nd) sq = add(sq, mul(yi[d], yi[d]));
            fp = add(fp, sq);
        }
    }
    ps.drain();
    if (on) {
        // rows never reached by the data stay zero, exactly as in the reference
        while (l < nk1) {
            const int row = l - K1 + 1;
#pragma unroll
            for (int q = 0; q < K1; ++q) {
                A_(row, q + 1) = RS_(s0, q);
                RS_(s0, q) = 0.0;
            }
#pragma unroll
            for (int d = 0; d < ND; ++d) {
                if (d < nd) {
                    ZD_(d, row) = ZS_(s0, d);
                    ZS_(s0, d) = 0.0;
                }
            }
            s0 = (s0 + 1 == K1) ? 0 : s0 + 1;
            ++l;
        }
        int sl_ = s0;
#pragma unroll
        for (int r = 0; r < K1; ++r) {
            const int row = l - K1 + 1 + r;
#pragma unroll
            for (int q = 0; q < K1; ++q) A_(row, q + 1) = RS_(sl_, q);
#pragma unroll
            for (int d = 0; d < ND; ++d)
                if (d < nd) ZD_(d, row) = ZS_(sl_, d);
            sl_ = (sl_ + 1 == K1) ? 0 : sl_ + 1;
        }
    }
#undef RS_
#undef ZS_
    return fp;
}

// fpchec, core:2507-2570 (per lane, iopt=-1 only)
template <int K>
__device__ int check_knots(const FitParams &p, double *wsd, const double *xp, int m, int n)
{
    constexpr int K1 = K + 1, K2 = K + 2;
    const int nk1 = n - K1, nk2 = nk1 + 1;
    auto X = [&](int i) { return xp[(size_t)(i - 1) * p.x_si]; };
    if (nk1 < K1 || nk1 > m) return IER_INPUT;
    int j = n;
    for (int i = 1; i <= K; ++i) {
        if (T_(i) > T_(i + 1)) return IER_INPUT;
        if (T_(j) < T_(j - 1)) return IER_INPUT;
        --j;
    }
    for (int i = K2; i <= nk2; ++i)
        if (T_(i) <= T_(i - 1)) return IER_INPUT;
    if (X(1) < T_(K1) || X(m) > T_(nk2)) return IER_INPUT;
    if (X(1) >= T_(K2) || X(m) <= T_(nk1)) return IER_INPUT;
    int i = 1, l = K2;
    const int nk3 = nk1 - 1;
    for (j = 2; j <= nk3; ++j) {
        const double tj = T_(j);
        ++l;
        const double tl = T_(l);
        while (i < m && X(i) <= tj) {
            ++i;
            if (i >= m) return IER_INPUT;
        }
        if (X(i) >= tl) return IER_INPUT;
    }
    return IER_OK;
}

// warp-aggregated append of curve id `b` to a device list
__device__ __forceinline__ void list_append(int *list, int *count, bool cond, int b, int lane)
{
    const unsigned mask = __ballot_sync(FULL, cond);
    if (mask == 0u) return;
    const int leader = __ffs(mask) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(count, __popc(mask));
    base = __shfl_sync(FULL, base, leader);
    if (cond) list[base + __popc(mask & ((1u << lane) - 1u))] = b;
}

// ---------------------------------------------------------------------------------------
// ONE iteration of fpcurf's knot loop (main_loop, core:4847-4998) for the 32 curves of a
// tile.  first != 0: also curfit's argument checks (core:1265-1285) and fpcurf's prologue
// (core:4761-4843).  Returns the curve's state after the pass: ST_LOOP (another knot set is
// needed), ST_PART2 (knots accepted, smoothing iteration follows), ST_DONE (result written).
// Between passes a curve lives in global memory only: n, t (in the caller's output arrays),
// nrdata(1:nrint) and the scalars fpold, fp0, nplus, iter.
// ---------------------------------------------------------------------------------------
template <int K, bool HAS_W, int IDIM>
__device__ int pass_tile(const FitParams &p, int b, bool valid, bool first,
                         double *__restrict__ wsd, int *__restrict__ wsi, double *rs, double *ring)
{
    constexpr int K1 = K + 1, K2 = K + 2;
    const int nd = IDIM > 0 ? IDIM : p.idim;
    const int m = p.m, nest = p.nest, iopt = p.iopt;
    const int nmax = m + K1, nmin = 2 * K1;
    const int bb = valid ? b : 0;
    const double *xp = p.x + (size_t)bb * p.x_sb;
    const double *yp = p.y + (size_t)bb * p.y_sb;
    const double *wp = HAS_W ? p.w + (size_t)bb * p.w_sb : nullptr;

    int state = valid ? ST_LOOP : ST_DONE;
    int ier = IER_OK, n = 0, nplus = 0, iter = 0;
    double fp = 0.0, fpold = 0.0, fp0 = 0.0, fpms = DBL_MAX;
    bool write_fit = false;  // outputs beyond ier are only written once fpcurf ran

    const double s = p.s[(size_t)bb * p.s_stride];
    const double acc = mul(kTol, s);
    double xb, xe;
    if (p.xb) {
        xb = p.xb[(size_t)bb * p.xbe_stride];
        xe = p.xe[(size_t)bb * p.xbe_stride];
    } else {
        xb = xp[0];
        xe = xp[(size_t)(m - 1) * p.x_si];
    }

    if (first) {
        // ---- curfit argument checks, core:1265-1285 (k, iopt, m, nest are call-level and were
        //      checked on the host) ----
        bool bad = false;
        double prev = xp[0];
        if (xb > prev) bad = true;
        if (!p.par) {
            for (int it = 1; it < m; ++it) {  // any(x(1:m-1) > x(2:m))
                const double cur = xp[(size_t)it * p.x_si];
                if (prev > cur) bad = true;
                prev = cur;
            }
        } else {
            // parcur (core:15255-15256): strictly increasing parameters, positive weights
            if (HAS_W && wp[0] <= 0.0) bad = true;
            for (int it = 1; it < m; ++it) {
                const double cur = xp[(size_t)it * p.x_si];
                if (prev >= cur) bad = true;
                if (HAS_W && wp[(size_t)it * p.w_si] <= 0.0) bad = true;
                prev = cur;
            }
        }
        if (xe < prev) bad = true;
        if (state == ST_LOOP) {
            if (iopt >= 0) {
                if (s < 0.0 || (fp_is_zero(s) && nest < (m + K1))) bad = true;
            } else {
                n = p.n[bb];
                if (n < nmin || n > nest) bad = true;
            }
            if (bad) {
                ier = IER_INPUT;
                state = ST_DONE;
            }
        }
        // knots / continuation state from the caller
        if (state == ST_LOOP && iopt != 0) {
            if (iopt > 0) n = p.n[bb];
            const bool have = (iopt < 0) || (n >= nmin && n <= nest && n <= nmax);
            if (iopt > 0 && !have) n = nmin;  // garbage n on a continuation: treat as a fresh start
            if (have) {
                for (int i = 1; i <= n; ++i) T_(i) = p.t[(size_t)bb * nest + (i - 1)];
                if (iopt > 0 && p.fpint && p.nrdata) {
                    for (int i = 1; i <= n; ++i) {
                        FPI_(i) = p.fpint[(size_t)bb * nest + (i - 1)];
                        NRD_(i) = p.nrdata[(size_t)bb * nest + (i - 1)];
                    }
                }
            }
            if (iopt < 0) {
                int j = n;
                for (int i = 1; i <= K1; ++i) {
                    T_(i) = xb;
                    T_(j) = xe;
                    --j;
                }
                // the clamped end knots are visible to the caller even if fpchec fails (core:1278-1284)
                j = n;
                for (int i = 1; i <= K1; ++i) {
                    p.t[(size_t)bb * nest + (i - 1)] = xb;
                    p.t[(size_t)bb * nest + (j - 1)] = xe;
                    --j;
                }
                const int chk = check_knots<K>(p, wsd, xp, m, n);
                if (chk != 0) {
                    ier = chk;
                    state = ST_DONE;
                }
            }
        }
        // ---- fpcurf prologue, core:4761-4843 ----
        if (state == ST_LOOP) {
            write_fit = true;
            if (iopt >= 0) {
                if (s <= 0.0) {
                    n = nmax;
                    if (nmax > nest) {
                        ier = IER_STORAGE;
                        state = ST_DONE;
                    } else {
                        interp_knots<K>(p, wsd, xp, m);
                    }
                } else if (iopt != 0 && n != nmin) {
                    fp0 = FPI_(n);
                    fpold = FPI_(n - 1);
                    nplus = NRD_(n);
                    if (fp0 <= s) {
                        n = nmin;
                        fpold = 0.0;
                        nplus = 0;
                        NRD_(1) = m - 2;
                    }
                } else {
                    n = nmin;
                    fpold = 0.0;
                    nplus = 0;
                    NRD_(1) = m - 2;
                }
            }
        }
    } else if (valid) {
        // ---- resume: reload the curve's state ----
        write_fit = true;
        n = p.n[b];
        ier = p.ier[b];
        nplus = p.st_nplus[b];
        iter = p.st_iter[b];
        fpold = p.st_fpold[b];
        fp0 = p.st_fp0[b];
        for (int i = 1; i <= n; ++i) T_(i) = p.t[(size_t)b * nest + (i - 1)];
        const int nri = n - nmin + 1;
        for (int i = 1; i <= nri; ++i) NRD_(i) = p.st_nrd[(size_t)b * p.nestp + (i - 1)];
    }

    // ---- one iteration of main_loop, core:4847-4998 ----
    int nk1 = n - K1, nrint = 0;
    bool ran_pass = false;
    {
        bool on = (state == ST_LOOP);
        if (on && !(iter <= m)) {
            state = ST_PART2;
            on = false;
        }
        if (on) {
            ++iter;
            if (n == nmin) ier = IER_LSQ;
            nrint = n - nmin + 1;
            nk1 = n - K1;
            for (int i = 1; i <= K1; ++i) T_(i) = xb;
            for (int i = nk1 + 1; i <= n; ++i) T_(i) = xe;
            ran_pass = true;
        }
        const double fp_pass = lsq_pass<K, HAS_W, IDIM>(p, wsd, rs, ring, xp, yp, wp, (size_t)p.x_si,
                                                        (size_t)p.y_si, (size_t)p.y_sd, (size_t)p.w_si,
                                                        m, nk1, on);
        if (on) {
            fp = fp_pass;
            if (ier == IER_LSQ) fp0 = fp;
            FPI_(n - 1) = fpold;
            FPI_(n) = fp0;
            NRD_(n) = nplus;
        }
        for (int d = 0; d < nd; ++d)
            back_subst<K1>(p, wsd, p.off_a, p.off_z + d * p.nestp, p.off_c + d * p.nestp, nk1, on);
        if (on) {
            if (iopt < 0) {
                state = ST_DONE;
            } else {
                fpms = sub(fp, s);
                if (fabs(fpms) < acc) state = ST_DONE;
                else if (fpms < 0.0) state = ST_PART2;
                else if (n == nmax) { ier = IER_INTERP; state = ST_DONE; }
                else if (n == nest) { ier = IER_STORAGE; state = ST_DONE; }
            }
        }
        on = on && (state == ST_LOOP);
        if (on) {  // number of knots to add, core:4927-4935
            if (ier == IER_OK) {
                int npl1 = nplus * 2;
                const double rn = (double)nplus;
                if (sub(fpold, fp) > acc) npl1 = x86_int(mul(rn, fpms) / sub(fpold, fp));
                nplus = min(nplus * 2, max(max(npl1, nplus / 2), 1));
            } else {
                nplus = 1;
                ier = IER_OK;
            }
            fpold = fp;
        }
        if (__any_sync(FULL, on))
            (void)residual_pass<K, true, HAS_W, IDIM>(p, wsd, ring, xp, yp, wp, (size_t)p.x_si,
                                                      (size_t)p.y_si, (size_t)p.y_sd, (size_t)p.w_si, m,
                                                      nk1, nrint, on);
        // add_new_knots, core:4968-4996 with fpknot core:8199-8275 inlined
        {
            const int maxplus = __reduce_max_sync(FULL, on ? nplus : 0);
            bool adding = on;
            for (int lp = 1; lp <= maxplus; ++lp) {
                if (!__any_sync(FULL, adding && lp <= nplus)) break;
                const bool act = adding && lp <= nplus;
                const int hi = __reduce_max_sync(FULL, act ? nrint : 0);
                double fpmax = 0.0;
                int number = 0, maxpt = 0, maxbeg = 0, jbegin = 1;
                // loads of 4 intervals are issued together (independent of the running maximum)
                for (int j0 = 1; j0 <= hi; j0 += 4) {
                    int jp[4];
                    double fv[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const bool in = act && (j0 + u) <= nrint;
                        jp[u] = in ? NRD_(j0 + u) : 0;
                        fv[u] = in ? FPI_(j0 + u) : 0.0;
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        if (act && (j0 + u) <= nrint) {
                            const int jpoint = jp[u];
                            if (jpoint != 0 && (number == 0 || fv[u] > fpmax)) {
                                fpmax = fv[u];
                                number = j0 + u;
                                maxpt = jpoint;
                                maxbeg = jbegin;
                            }
                            jbegin = jbegin + jpoint + 1;
                        }
                    }
                }
                const bool ins = act && number != 0;
                if (act && number == 0) adding = false;  // nothing left to refine: later calls are no-ops too
                const int ihalf = maxpt / 2 + 1;
                const int nrx = maxbeg + ihalf;
                const int next = number + 1;
                // shift right by one, descending in chunks of 4: all loads of a chunk precede its
                // stores, and a chunk only overwrites slots that later (lower) chunks never read
                for (int j0 = hi; j0 >= 1; j0 -= 4) {
                    double fv[4], tv[4];
                    int nv[4];
                    bool mv[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int jj = j0 - u;
                        mv[u] = ins && jj >= next && jj <= nrint && jj >= 1;
                        fv[u] = mv[u] ? FPI_(jj) : 0.0;
                        nv[u] = mv[u] ? NRD_(jj) : 0;
                        tv[u] = mv[u] ? T_(jj + K) : 0.0;
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int jj = j0 - u;
                        if (mv[u]) {
                            FPI_(jj + 1) = fv[u];
                            NRD_(jj + 1) = nv[u];
                            T_(jj + K + 1) = tv[u];
                        }
                    }
                }
                if (ins) {
                    NRD_(number) = ihalf - 1;
                    NRD_(next) = maxpt - ihalf;
                    const double am = (double)maxpt;
                    double an = (double)(ihalf - 1);
                    FPI_(number) = mul(fpmax, an) / am;
                    an = (double)(maxpt - ihalf);
                    FPI_(next) = mul(fpmax, an) / am;
                    T_(next + K) = xp[(size_t)(nrx - 1) * p.x_si];
                    ++n;
                    ++nrint;
                    if (n == nmax) {
                        interp_knots<K>(p, wsd, xp, m);
                        iter = 0;
                        adding = false;
                    } else if (n == nest) {
                        adding = false;
                    }
                }
            }
        }
    }
    if (state == ST_PART2 && ier == IER_LSQ) state = ST_DONE;  // core:5002

    // ---- write back ----
    if (valid) {
        int passes = 0;
        if (p.stats && write_fit) {
            passes = (first ? 0 : p.stats[2 * (size_t)b]) + (ran_pass ? 1 : 0);
            p.stats[2 * (size_t)b] = passes;
            if (first) p.stats[2 * (size_t)b + 1] = 0;
        }
        p.ier[b] = ier;
        if (write_fit) {
            p.n[b] = n;
            for (int i = 1; i <= n; ++i) p.t[(size_t)b * nest + (i - 1)] = T_(i);
            if (state == ST_DONE) {
                if (!(ier == IER_STORAGE && iopt >= 0 && s <= 0.0)) p.fp[b] = fp;
                // coordinate d's coefficients start at c(d*n) (fppara core:8981-8985)
                const int nk = n - K1;
                for (int d = 0; d < nd; ++d)
                    for (int i = 1; i <= nk; ++i)
                        p.c[(size_t)b * p.c_sb + (size_t)d * n + (i - 1)] = CD_(d, i);
            }
            if (state != ST_LOOP) {
                // knots are final: the continuation state the caller may keep (iopt=1 next time)
                if (p.fpint && p.nrdata && iopt >= 0) {
                    for (int i = 1; i <= n; ++i) {
                        p.fpint[(size_t)b * nest + (i - 1)] = FPI_(i);
                        p.nrdata[(size_t)b * nest + (i - 1)] = NRD_(i);
                    }
                }
                if (state == ST_PART2) {
                    p.st_fp0[b] = fp0;
                    // hand the triangular system of this (accepted) knot set to part 2
                    long long off = -1;
                    const int rowlen = K1 + nd;
                    if (p.az_pool && ran_pass) {
                        const unsigned long long need = (unsigned long long)(n - K1) * rowlen;
                        const unsigned long long at = atomicAdd(p.az_cursor, need);
                        if (at + need <= (unsigned long long)p.az_cap) off = (long long)at;
                    }
                    if (p.az_off) p.az_off[b] = off;
                    if (off >= 0) {
                        double *dst = p.az_pool + off;
                        const int nk = n - K1;
                        for (int i = 1; i <= nk; ++i) {
#pragma unroll
                            for (int q = 0; q < K1; ++q) dst[q] = A_(i, q + 1);
                            for (int d = 0; d < nd; ++d) dst[K1 + d] = ZD_(d, i);
                            dst += rowlen;
                        }
                        p.st_fpold[b] = fp;   // fp of the least-squares spline on these knots
                    }
                }
            } else {
                p.st_nplus[b] = nplus;
                p.st_iter[b] = iter;
                p.st_fpold[b] = fpold;
                p.st_fp0[b] = fp0;
                const int nri = n - nmin + 1;
                for (int i = 1; i <= nri; ++i) p.st_nrd[(size_t)b * p.nestp + (i - 1)] = NRD_(i);
            }
        }
    }
    return valid ? state : ST_DONE;
}

// ---------------------------------------------------------------------------------------
// Part 2 of fpcurf (core:5004-5084) for the 32 curves of a tile: the knots are final; find
// the smoothing parameter p with f(p)=s.  The triangularised system (R, z) and the basis
// values q of the accepted knot set are regenerated by one LSQ pass (bit-identical to the
// last pass of part 1), which keeps no per-curve matrix state alive between kernels.
// ---------------------------------------------------------------------------------------
template <int K, bool HAS_W, int IDIM>
__device__ void part2_tile(const FitParams &p, int b, bool valid, double *__restrict__ wsd,
                           int *__restrict__ wsi, double *rs, double *ring)
{
    constexpr int K1 = K + 1, K2 = K + 2, ND = IDIM > 0 ? IDIM : kMaxIdim;
    const int nd = IDIM > 0 ? IDIM : p.idim;
    (void)wsi;
    const int m = p.m, nest = p.nest;
    const int nmin = 2 * K1;
    const int bb = valid ? b : 0;
    const double *xp = p.x + (size_t)bb * p.x_sb;
    const double *yp = p.y + (size_t)bb * p.y_sb;
    const double *wp = HAS_W ? p.w + (size_t)bb * p.w_sb : nullptr;
    const double s = p.s[(size_t)bb * p.s_stride];
    const double acc = mul(kTol, s);
    int state = valid ? ST_PART2 : ST_DONE;
    int ier = IER_OK, n = nmin, p_iters = 0;
    double fp0 = 0.0;
    if (valid) {
        n = p.n[b];
        fp0 = p.st_fp0[b];
        for (int i = 1; i <= n; ++i) T_(i) = p.t[(size_t)b * nest + (i - 1)];
    }
    const int nk1 = n - K1;
    const int n8 = n - nmin;
    const bool on2 = valid;
    // The part-2 list is sorted by knot count, so the lanes of a warp hold unrelated curves and
    // their x/y reads are scattered (one 32 B sector per 8 B).  Gather the curve's data ONCE into
    // the warp's interleaved workspace; the LSQ pass and every f(p) evaluation then stream it
    // with fully coalesced loads.
    {
        constexpr int U = 4;
        for (int it0 = 0; it0 < m; it0 += U) {
            double xv[U], yv[U][ND], wv[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int it = min(it0 + u, m - 1);
                xv[u] = xp[(size_t)it * p.x_si];
#pragma unroll
                for (int d = 0; d < ND; ++d)
                    if (d < nd) yv[u][d] = yp[(size_t)it * p.y_si + (size_t)d * p.y_sd];
                wv[u] = HAS_W ? wp[(size_t)it * p.w_si] : 1.0;
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (it0 + u < m) {
                    WS(p.off_x, it0 + u) = xv[u];
#pragma unroll
                    for (int d = 0; d < ND; ++d)
                        if (d < nd) WS(p.off_y + d * m, it0 + u) = yv[u][d];
                    if (HAS_W) WS(p.off_w, it0 + u) = wv[u];
                }
            }
        }
    }
    const double *xg = &WS(p.off_x, 0), *yg = &WS(p.off_y, 0);
    const double *wg = HAS_W ? &WS(p.off_w, 0) : nullptr;
    const size_t yg_sd = (size_t)m * 32;
    // (R, z) of the accepted knot set: normally handed over by the pass that accepted it (the same values
    // the pass below would regenerate); curves without stored rows (pool exhausted, or part 2 reached without
    // a pass in the last launch) run the least-squares pass again
#if FPB_FIT_STOREQ
    const long long az = -1;   // the stored basis values come from this pass
#else
    const long long az = (valid && p.az_off) ? p.az_off[b] : -1;
#endif
    double fp = 0.0;
    if (__any_sync(FULL, on2 && az < 0)) {
        const double fpl = lsq_pass<K, HAS_W, IDIM>(p, wsd, rs, ring, xg, yg, wg, 32, 32, yg_sd, 32, m, nk1,
                                                    on2 && az < 0);
        if (az < 0) fp = fpl;
    }
    if (on2 && az >= 0) {
        const double *src = p.az_pool + az;
        const int rowlen = K1 + nd;
        for (int i = 1; i <= nk1; ++i) {
#pragma unroll
            for (int q = 0; q < K1; ++q) A_(i, q + 1) = src[q];
            for (int d = 0; d < nd; ++d) ZD_(d, i) = src[K1 + d];
            src += rowlen;
        }
        fp = p.st_fpold[b];
    }
    double fpms = sub(fp, s);
    // fpdisc, core:5453-5495
    {
        const int hi = __reduce_max_sync(FULL, on2 ? nk1 : 0);
        double fac = 0.0;
        if (on2) fac = (double)(nk1 - K) / sub(T_(nk1 + 1), T_(K1));
        for (int l = K2; l <= hi; ++l) {
            if (on2 && l <= nk1) {
                double hd[2 * K1];
                const double tl = T_(l);
#pragma unroll
                for (int j = 1; j <= K1; ++j) {
                    hd[j - 1] = sub(tl, T_(l + j - K2));
                    hd[j + K1 - 1] = sub(tl, T_(l + j));
                }
                const int lmk = l - K1;
#pragma unroll
                for (int j = 1; j <= K2; ++j) {
                    double prod = hd[j - 1];
#pragma unroll
                    for (int i = 1; i <= K; ++i) prod = mul(mul(prod, hd[j - 1 + i]), fac);
                    const int lp = lmk + j - 1;
                    B_(lmk, j) = sub(T_(lp + K1), T_(lp)) / prod;
                }
            }
        }
    }
    double p1 = 0.0, f1 = sub(fp0, s), p3 = -1.0, f3 = fpms, p2 = 0.0, f2 = 0.0, pp = 0.0;
    bool check1 = false, check3 = false;
    {
        const int hi = __reduce_max_sync(FULL, on2 ? nk1 : 0);
        for (int i = 1; i <= hi; ++i)
            if (on2 && i <= nk1) pp = add(pp, A_(i, 1));
        // fpcurf: p = nk1/sum(a(:,1)) (core:5031-5033); fppara: p = sum(a(:,1))/nk1 (core:9119)
        if (on2) pp = p.par ? pp / (double)nk1 : (double)nk1 / pp;
    }
    int it2 = 0;
    while (__any_sync(FULL, state == ST_PART2)) {
        bool on = (state == ST_PART2);
        if (on && !(it2 < kMaxIt)) {
            ier = IER_MAXIT;
            state = ST_DONE;
            on = false;
        }
        double pinv = 0.0;
        if (on) {
            ++it2;
            ++p_iters;
            pinv = 1.0 / pp;
        }
        const int hi = __reduce_max_sync(FULL, on ? nk1 : 0);
        const int hi8 = __reduce_max_sync(FULL, on ? n8 : 0);
        // The smoothing rows b(it,:)/p are rotated into g TWO AT A TIME (fp_rotate_shifted / _vec,
        // core:5800-5820 / 5845-5862): band row j passes through row `it`'s rotation and, one step
        // later and still in registers, through row `it+1`'s, and is stored once.  Row it+1 at band
        // row j needs only what row `it` left in band row j and its own state from band row j-1, so
        // the arithmetic and its order are exactly the reference's; the band matrix, which part 2
        // re-reads from DRAM (its working set plus the point streams exceed L2: 82 % of the HBM peak,
        // profiles/r01g_kernels.md), is read and written half as often.
        // g := a, g(:,k2) := 0, c := z (core:5047-5049) is fused into the first pair's sweep, which
        // visits every row 1..nk1: it reads (a, z) instead of (g, c) and stores every row it passes,
        // rotated or not (n8 >= 1 for every curve that reaches part 2).
        // One rotation step of a smoothing row against band row (g, c) with `noff` columns to its right.
        auto sweep_step = [&](double (&h)[K2], double (&y)[ND], double (&g)[K2], double (&c)[ND], int noff) {
            const double piv = h[0];
            if (p.par || !fp_is_zero(piv)) {  // the vector variant has no zero-pivot skip
                double cs, sn;
                FPB_GIVENS(piv, g[0], cs, sn);
#pragma unroll
                for (int d = 0; d < ND; ++d)
                    if (d < nd) rota(cs, sn, y[d], c[d]);
#pragma unroll
                for (int q = 1; q <= K1; ++q)
                    if (q <= noff) rota(cs, sn, h[q], g[q]);
            }
#pragma unroll
            for (int q = 0; q < K1; ++q) h[q] = h[q + 1];
            h[K1] = 0.0;
        };
        for (int it = 1; it <= hi8; it += 2) {
            const bool a_on = on && it <= n8, b_on = on && (it + 1) <= n8;
            const bool first_pair = (it == 1);
            double hA[K2], hB[K2], yA[ND], yB[ND], gp[K2], cp[ND];
#pragma unroll
            for (int q = 0; q < K2; ++q) {
                hA[q] = a_on ? mul(B_(it, q + 1), pinv) : 0.0;
                hB[q] = b_on ? mul(B_(it + 1, q + 1), pinv) : 0.0;
                gp[q] = 0.0;
            }
#pragma unroll
            for (int d = 0; d < ND; ++d) yA[d] = yB[d] = cp[d] = 0.0;
            for (int j = it; j <= hi + 1; ++j) {
                // band row j is requested first; row it+1's step on band row j-1 (in registers) covers
                // the latency
                const bool a_act = a_on && j <= nk1;
                double ga[K2], ca[ND];
#pragma unroll
                for (int q = 0; q < K2; ++q)
                    ga[q] = a_act ? (first_pair ? (q < K1 ? A_(j, q + 1) : 0.0) : G_(j, q + 1)) : 0.0;
#pragma unroll
                for (int d = 0; d < ND; ++d)
                    ca[d] = (a_act && d < nd) ? (first_pair ? ZD_(d, j) : CD_(d, j)) : 0.0;
                const int jb = j - 1;
                if (a_on && jb >= it && jb <= nk1) {
                    if (b_on && jb >= it + 1) sweep_step(hB, yB, gp, cp, min(nk1 - jb, K1));
#pragma unroll
                    for (int q = 0; q < K2; ++q) G_(jb, q + 1) = gp[q];
#pragma unroll
                    for (int d = 0; d < ND; ++d)
                        if (d < nd) CD_(d, jb) = cp[d];
                }
                if (a_act) sweep_step(hA, yA, ga, ca, min(nk1 - j, K1));
#pragma unroll
                for (int q = 0; q < K2; ++q) gp[q] = ga[q];
#pragma unroll
                for (int d = 0; d < ND; ++d) cp[d] = ca[d];
            }
        }
        for (int d = 0; d < nd; ++d)
            back_subst<K2>(p, wsd, p.off_g, p.off_c + d * p.nestp, p.off_c + d * p.nestp, nk1, on);
        const double fpn = residual_pass<K, false, HAS_W, IDIM>(p, wsd, ring, xg, yg, wg, 32, 32, yg_sd, 32, m,
                                                                nk1, 0, on);
        if (on) {
            fp = fpn;
            fpms = sub(fp, s);
            if (fabs(fpms) < acc) {
                state = ST_DONE;
            } else {
                // root_finding_iterate core:11641-11691 + fprati core:11996-12025
                const double con1 = 0.1, con9 = 0.9, con4 = 0.04;
                bool ret = false, ok = true;
                p2 = pp;
                f2 = fpms;
                if (!check3) {
                    if (sub(f2, f3) > acc) {
                        check3 = f2 < 0.0;
                    } else {
                        p3 = p2;
                        f3 = f2;
                        pp = mul(pp, con4);
                        if (pp <= p1) pp = add(mul(p1, con9), mul(p2, con1));
                        ret = true;
                    }
                }
                if (!ret && !check1) {
                    if (sub(f1, f2) > acc) {
                        check1 = f2 > 0.0;
                    } else {
                        p1 = p2;
                        f1 = f2;
                        pp = pp / con4;
                        if (p3 >= 0.0 && pp >= p3) pp = add(mul(p2, con1), mul(p3, con9));
                        ret = true;
                    }
                }
                if (!ret) {
                    if (f2 >= f1 || f2 <= f3) {
                        ok = false;
                    } else {
                        if (p3 > 0.0) {
                            const double h1 = mul(f1, sub(f2, f3));
                            const double h2 = mul(f2, sub(f3, f1));
                            const double h3 = mul(f3, sub(f1, f2));
                            const double num = add(add(mul(mul(p1, p2), h3), mul(mul(p2, p3), h1)),
                                                   mul(mul(p3, p1), h2));
                            const double den = add(add(mul(p1, h1), mul(p2, h2)), mul(p3, h3));
                            pp = -num / den;
                        } else {
                            const double num = sub(mul(mul(p1, sub(f1, f3)), f2),
                                                   mul(mul(p2, sub(f2, f3)), f1));
                            pp = num / mul(sub(f1, f2), f3);
                        }
                        if (f2 >= 0.0) { p1 = p2; f1 = f2; } else { p3 = p2; f3 = f2; }
                    }
                }
                if (!ok) {
                    ier = IER_S_SMALL;
                    state = ST_DONE;
                }
            }
        }
    }
    if (valid) {
        p.ier[b] = ier;
        p.fp[b] = fp;
        for (int d = 0; d < nd; ++d)
            for (int i = 1; i <= nk1; ++i)
                p.c[(size_t)b * p.c_sb + (size_t)d * n + (i - 1)] = CD_(d, i);
        if (p.stats) p.stats[2 * (size_t)b + 1] = p_iters;
    }
}

// persistent warps; tiles of 32 list entries are handed out through an atomic counter
// STAGED: list entries may carry the "no pass yet" flag (chunks of a host batch joining a running pipeline); a
// separate instantiation so that the device-resident path keeps its kernel-uniform `first`.
template <int K, bool HAS_W, int IDIM, bool STAGED>
__global__ void __launch_bounds__(kFitThreads, fit_blocks_per_sm(K, IDIM))
    curfit_pass_kernel(FitParams p, const int *__restrict__ in_list, const int *in_count,
                       int *out_list, int *out_count, int *p2_list, int *p2_count, int first)
{
    extern __shared__ double fit_smem[];
    double *rs = fit_smem + threadIdx.x;
    double *ring = rs + (size_t)((K + 1) * (K + 1) + (K + 1) * (IDIM > 0 ? IDIM : p.idim)) * kFitThreads;  // after the band window
    const int lane = threadIdx.x & 31;
    const int warp_slot = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    double *wsd = p.ws + (size_t)warp_slot * p.ws_per_lane * 32 + lane;
    int *wsi = p.wsi + (size_t)warp_slot * p.nestp * 32 + lane;
    const int count = first ? p.nbatch : *in_count;
    for (;;) {
        unsigned tile = 0;
        if (lane == 0) tile = atomicAdd(p.counter, 1u);
        tile = __shfl_sync(FULL, tile, 0);
        const long long base = (long long)tile * 32;
        if (base >= count) break;
        const bool valid = base + lane < count;
        int b = 0;
        if (valid) b = first ? (int)(base + lane) : in_list[base + lane];
        // a list entry with the top bit set is a curve that has not had its first pass yet (chunks of a
        // host batch join the running pipeline as they arrive, launch_fit_admit)
        bool fresh = first != 0;
        if (STAGED) {
            fresh = fresh || (valid && b < 0);
            b &= 0x7fffffff;
        }
        const int st = pass_tile<K, HAS_W, IDIM>(p, b, valid, fresh, wsd, wsi, rs, ring);
        __syncwarp();
        list_append(out_list, out_count, valid && st == ST_LOOP, b, lane);
        const bool to2 = valid && st == ST_PART2;
        list_append(p2_list, p2_count, to2, b, lane);
        if (to2) atomicAdd(&p.hist[min(p.n[b], p.nestp)], 1);
    }
}

template <int K, bool HAS_W, int IDIM>
__global__ void __launch_bounds__(kFitThreads, fit_p2_blocks_per_sm(K, IDIM))
    curfit_part2_kernel(FitParams p, const int *__restrict__ list, const int *count_ptr)
{
    extern __shared__ double fit_smem[];
    double *rs = fit_smem + threadIdx.x;
    double *ring = rs + (size_t)((K + 1) * (K + 1) + (K + 1) * (IDIM > 0 ? IDIM : p.idim)) * kFitThreads;  // after the band window
    const int lane = threadIdx.x & 31;
    const int warp_slot = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    double *wsd = p.ws + (size_t)warp_slot * p.ws_per_lane * 32 + lane;
    int *wsi = p.wsi + (size_t)warp_slot * p.nestp * 32 + lane;
    const int count = *count_ptr;
    for (;;) {
        unsigned tile = 0;
        if (lane == 0) tile = atomicAdd(p.counter, 1u);
        tile = __shfl_sync(FULL, tile, 0);
        const long long base = (long long)tile * 32;
        if (base >= count) break;
        const bool valid = base + lane < count;
        const int b = valid ? list[base + lane] : 0;
        part2_tile<K, HAS_W, IDIM>(p, b, valid, wsd, wsi, rs, ring);
        __syncwarp();
    }
}

template <int K, bool HAS_W>
cudaError_t launch_pass_kw(const FitParams &p, int grid, cudaStream_t st, const int *in_list,
                           const int *in_count, int *out_list, int *out_count, int *p2_list,
                           int *p2_count, int first)
{
    constexpr int D = FPB_FIT_IDIM;
    const unsigned smem = fit_smem_bytes_stream(p.k, p.idim, HAS_W, p.sdepth);
    auto kern = p.staged ? curfit_pass_kernel<K, HAS_W, D, true> : curfit_pass_kernel<K, HAS_W, D, false>;
    if (smem > 48u * 1024u) {
        const cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    kern<<<grid, kFitThreads, smem, st>>>(p, in_list, in_count, out_list, out_count, p2_list, p2_count, first);
    return cudaGetLastError();
}

template <int K, bool HAS_W>
cudaError_t launch_part2_kw(const FitParams &p, int grid, cudaStream_t st, const int *list, const int *count)
{
    constexpr int D = FPB_FIT_IDIM;
    const unsigned smem = fit_smem_bytes_stream(p.k, p.idim, HAS_W, p.sdepth);
    auto kern = curfit_part2_kernel<K, HAS_W, D>;
    if (smem > 48u * 1024u) {
        const cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    kern<<<grid, kFitThreads, smem, st>>>(p, list, count);
    return cudaGetLastError();
}

#if FPB_FIT_IDIM == 1
// counting sort of the part-2 list by knot count, largest first: curves of one warp then
// do the same amount of work per p-iteration, and the longest curves start first.
__global__ void p2_offsets_kernel(int *hist, int nbins)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    int run = 0;
    for (int n = nbins - 1; n >= 0; --n) {
        const int c = hist[n];
        hist[n] = run;
        run += c;
    }
}

__global__ void p2_scatter_kernel(const int *list, const int *count_ptr, const int *n_arr,
                                  int *offsets, int nbins, int *sorted)
{
    const int count = *count_ptr;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) {
        const int b = list[i];
        const int pos = atomicAdd(&offsets[min(n_arr[b], nbins - 1)], 1);
        sorted[pos] = b;
    }
}

// curves c0 .. c0+nc-1 join the pipeline: appended to `list` behind its `base` entries, flagged "no pass yet"
__global__ void fit_admit_kernel(int *list, int *count, int base, int c0, int nc)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nc; i += gridDim.x * blockDim.x)
        list[base + i] = (c0 + i) | (int)0x80000000;
    if (blockIdx.x == 0 && threadIdx.x == 0) *count = base + nc;
}

// chord-length parameters of parcur (ipar=0, core:15238-15253): u(1)=0, u(i)=u(i-1)+norm2(dx),
// normalised to [0,1] with u(m)=1 exactly.  NORM2 is evaluated as gfortran does (scaled form).
// A curve of zero total length keeps its un-normalised u (all zero) and is rejected with ier=10
// by the strict-monotonicity check of the first pass, as in the reference.
__global__ void parcur_param_kernel(int nbatch, int m, int idim, const double *x, long long x_sb,
                                    long long x_si, long long x_sd, double *u, long long u_sb,
                                    long long u_si)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nbatch) return;
    const double *xp = x + (size_t)b * x_sb;
    double *up = u + (size_t)b * u_sb;
    double prev[kMaxIdim], acc = 0.0;
    for (int d = 0; d < idim; ++d) prev[d] = xp[(size_t)d * x_sd];
    up[0] = 0.0;
    for (int i = 1; i < m; ++i) {
        double result = 0.0, scale = 1.0;
        for (int d = 0; d < idim; ++d) {
            const double cur = xp[(size_t)i * x_si + (size_t)d * x_sd];
            const double df = sub(cur, prev[d]);
            prev[d] = cur;
            if (df != 0.0) {
                const double a = fabs(df);
                if (scale < a) {
                    const double val = scale / a;
                    result = add(1.0, mul(mul(result, val), val));
                    scale = a;
                } else {
                    const double val = a / scale;
                    result = add(result, mul(val, val));
                }
            }
        }
        acc = add(acc, mul(scale, sqrt(result)));
        up[(size_t)i * u_si] = acc;
    }
    if (acc <= 0.0) return;
    for (int i = 1; i < m; ++i) up[(size_t)i * u_si] = up[(size_t)i * u_si] / acc;
    up[(size_t)(m - 1) * u_si] = 1.0;
}
#endif

}  // namespace

// per-IDIM entry points (one translation unit each)
#define FPB_CAT2(a, b) a##b
#define FPB_CAT(a, b) FPB_CAT2(a, b)
cudaError_t FPB_CAT(launch_curfit_pass_d, FPB_FIT_IDIM)(const FitParams &p, int grid, cudaStream_t st,
                                                       const int *in_list, const int *in_count,
                                                       int *out_list, int *out_count, int *p2_list,
                                                       int *p2_count, int first)
{
#define FPB_CASE(KK)                                                                                       \
    case KK:                                                                                               \
        return p.w ? launch_pass_kw<KK, true>(p, grid, st, in_list, in_count, out_list, out_count, p2_list, \
                                              p2_count, first)                                             \
                   : launch_pass_kw<KK, false>(p, grid, st, in_list, in_count, out_list, out_count,        \
                                               p2_list, p2_count, first);
    switch (p.k) {
        FPB_CASE(1) FPB_CASE(2) FPB_CASE(3) FPB_CASE(4) FPB_CASE(5)
    default: return cudaErrorInvalidValue;
    }
#undef FPB_CASE
}

cudaError_t FPB_CAT(launch_curfit_part2_d, FPB_FIT_IDIM)(const FitParams &p, int grid, cudaStream_t st,
                                                        const int *list, const int *count)
{
#define FPB_CASE(KK)                                                              \
    case KK:                                                                      \
        return p.w ? launch_part2_kw<KK, true>(p, grid, st, list, count)          \
                   : launch_part2_kw<KK, false>(p, grid, st, list, count);
    switch (p.k) {
        FPB_CASE(1) FPB_CASE(2) FPB_CASE(3) FPB_CASE(4) FPB_CASE(5)
    default: return cudaErrorInvalidValue;
    }
#undef FPB_CASE
}

#if FPB_FIT_IDIM == 1
FitLayout fit_layout(int m, int k, int nest, int idim, int periodic)
{
    const int k1 = k + 1, k2 = k + 2;
    FitLayout L;
    const int ncap = periodic ? m + 2 * k : m + k1;  // knots a curve can reach (interpolation)
    int nestp = nest < ncap ? nest : ncap;
    if (nestp < 2 * k1) nestp = 2 * k1;
    L.nestp = nestp;
    int off = 0;
    L.off_t = off; off += nestp;
    L.off_c = off; off += nestp * idim;
    L.off_z = off; off += nestp * idim;
    L.off_fpint = off; off += nestp;
    L.off_a = off; off += nestp * k1;
    L.off_g = off; off += nestp * k2;
    L.off_b = off; off += nestp * k2;
    L.off_q = off; off += (periodic || FPB_FIT_STOREQ) ? m * k1 : 0;  // stored basis values: percur only
    L.off_x = off; off += m;   // part 2: the curve's own data, gathered once
    L.off_y = off; off += m * idim;
    L.off_w = off; off += m;
    L.off_a2 = off; off += periodic ? nestp * k : 0;
    L.off_g2 = off; off += periodic ? nestp * k1 : 0;
    L.ws_per_lane = off;
    return L;
}

int fit_max_resident_warps(int k, int idim, int nsm)
{
    const int c = idim_class(idim);
    return nsm * (fit_blocks_per_sm(k, c) > fit_p2_blocks_per_sm(k, c) ? fit_blocks_per_sm(k, c) : fit_p2_blocks_per_sm(k, c)) *
           (kFitThreads / 32);
}

int fit_p2_resident_warps(int k, int idim, int nsm)
{
    return nsm * fit_p2_blocks_per_sm(k, idim_class(idim)) * (kFitThreads / 32);
}

cudaError_t launch_curfit_pass(const FitParams &p, int grid_blocks, cudaStream_t stream,
                               const int *in_list, const int *in_count, int *out_list,
                               int *out_count, int *p2_list, int *p2_count, int first)
{
    if (p.periodic)
        return launch_percur_pass(p, grid_blocks, stream, in_list, in_count, out_list, out_count, p2_list, p2_count, first);
    switch (idim_class(p.idim)) {
    case 1: return launch_curfit_pass_d1(p, grid_blocks, stream, in_list, in_count, out_list, out_count, p2_list, p2_count, first);
    case 2: return launch_curfit_pass_d2(p, grid_blocks, stream, in_list, in_count, out_list, out_count, p2_list, p2_count, first);
    case 3: return launch_curfit_pass_d3(p, grid_blocks, stream, in_list, in_count, out_list, out_count, p2_list, p2_count, first);
    default: return launch_curfit_pass_d0(p, grid_blocks, stream, in_list, in_count, out_list, out_count, p2_list, p2_count, first);
    }
}

cudaError_t launch_curfit_part2(const FitParams &p, int grid_blocks, cudaStream_t stream,
                                const int *list, const int *count)
{
    if (p.periodic) return launch_percur_part2(p, grid_blocks, stream, list, count);
    switch (idim_class(p.idim)) {
    case 1: return launch_curfit_part2_d1(p, grid_blocks, stream, list, count);
    case 2: return launch_curfit_part2_d2(p, grid_blocks, stream, list, count);
    case 3: return launch_curfit_part2_d3(p, grid_blocks, stream, list, count);
    default: return launch_curfit_part2_d0(p, grid_blocks, stream, list, count);
    }
}

cudaError_t launch_fit_admit(int *list, int *count, int base, int c0, int nc, cudaStream_t stream)
{
    if (nc <= 0) return cudaSuccess;
    fit_admit_kernel<<<std::min((nc + 255) / 256, 1024), 256, 0, stream>>>(list, count, base, c0, nc);
    return cudaGetLastError();
}

cudaError_t launch_p2_sort(const int *list, const int *count, const int *n_arr, int *hist,
                           int nbins, int *sorted, int nsm, cudaStream_t stream)
{
    p2_offsets_kernel<<<1, 32, 0, stream>>>(hist, nbins);
    p2_scatter_kernel<<<nsm * 4, 256, 0, stream>>>(list, count, n_arr, hist, nbins, sorted);
    return cudaGetLastError();
}

cudaError_t launch_parcur_param(int nbatch, int m, int idim, const double *x, long long x_sb,
                                long long x_si, long long x_sd, double *u, long long u_sb,
                                long long u_si, cudaStream_t stream)
{
    parcur_param_kernel<<<(nbatch + 127) / 128, 128, 0, stream>>>(nbatch, m, idim, x, x_sb, x_si, x_sd, u,
                                                                 u_sb, u_si);
    return cudaGetLastError();
}
#endif

}  // namespace fpb
