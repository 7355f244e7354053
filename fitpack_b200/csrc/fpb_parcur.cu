// fpb_parcur.cu -- parametric curves: batched parcur (smoothing fit of idim coordinate curves
// sharing one knot vector) and curev (evaluation), SURVEY 8f rank 2.
//
//   parcur  src/fitpack_core.F90:15201-15285 (argument checks, chord-length parameters)
//   fppara  src/fitpack_core.F90:8822-9177   (knot loop, vector-RHS LSQ, p-iteration)
//   curev   src/fitpack_core.F90:1126-1182
//   C binding: fp_parcur_c / fp_curev_c, src/fitpack_core_c.f90:90-100, 178-184
//
// The fit itself is the pass pipeline of fpb_curfit.cu instantiated for IDIM coordinates (one
// Givens rotation, idim right-hand sides); this file holds the evaluation kernel and the entry
// points.  curev computes the k+1 basis values of a point ONCE and reuses them for all idim
// coordinates (the reference does the same); values are bit-identical to the reference's
// operation order (no FMA).
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../include/fitpack_b200.h"
#include "fpb_common.cuh"
#include "fpb_engine.h"

namespace fpb {

namespace {

constexpr unsigned FULLW = 0xffffffffu;

// One warp per curve, lanes over the evaluation points (32 consecutive points per step).  The
// curve's knots and the coefficients of all coordinates are staged once in the warp's slice of
// shared memory; x is written in the reference's layout x(idim,m), i.e. idim consecutive doubles per
// point.  u is non-decreasing (checked first: any(u(2:m) < u(1:m-1)) -> ier = 10, nothing written,
// core:1145), so the knot interval of a point is found by a short linear walk that starts from the
// interval the previous 32-point step ended in -- the same interval fp_knot_interval returns.
constexpr int kCurevWarps = 8;

template <int K>
__global__ void __launch_bounds__(kCurevWarps * 32)
    curev_kernel(int nspl, int idim, const double *__restrict__ t, const int *__restrict__ n,
                 const double *__restrict__ c, int ldt, long long ldc, const double *__restrict__ u,
                 double *__restrict__ x, int m, int *__restrict__ ier, int smem_ld, int tab)
{
    constexpr int K1 = K + 1, NS = K * (K + 1) / 2;
    extern __shared__ double curev_smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    // per warp: knots, idim coefficient rows and (tab) the span table of bspl_tab: NS refined reciprocals
    // per knot interval + one mask word per interval (fpb_common.cuh) -- same bits, a third of the FP64 work
    // `tab` = knot intervals the table holds (0: none); curves with more intervals use plain divisions
    const size_t per_warp = (size_t)smem_ld * (1 + idim) + (tab ? (size_t)NS * tab + (tab + 1) / 2 : 0);
    double *ts = curev_smem + (size_t)wib * per_warp;
    double *cs = ts + smem_ld;
    double *spans = ts + (size_t)smem_ld * (1 + idim);
    unsigned *spmask = reinterpret_cast<unsigned *>(spans + (size_t)NS * tab);
    const int wpb = blockDim.x >> 5;  // warps per CTA: kCurevWarps unless the tables of a long curve need the room
    const long long warp0 = (long long)blockIdx.x * wpb + wib;
    const long long nwarps = (long long)gridDim.x * wpb;
    for (long long j = warp0; j < nspl; j += nwarps) {
        const int nj = n[j];
        if (nj < 2 * K1 || nj > ldt) {  // not a spline of degree K
            if (lane == 0) ier[j] = IER_INPUT;
            continue;
        }
        const double *tj = t + j * ldt, *cj = c + j * ldc, *uj = u + j * (long long)m;
        double *xj = x + j * (long long)m * idim;
        bool bad = false;
        for (int i = lane + 1; i < m; i += 32) bad = bad || (uj[i] < uj[i - 1]);
        if (__any_sync(FULLW, bad)) {
            if (lane == 0) ier[j] = IER_INPUT;
            continue;
        }
        if (lane == 0) ier[j] = IER_OK;
        const int nk1 = nj - K1;
        __syncwarp();
        for (int i = lane; i < nj; i += 32) ts[i] = tj[i];
        for (int d = 0; d < idim; ++d)
            for (int i = lane; i < nk1; i += 32) cs[d * smem_ld + i] = cj[(long long)d * nj + i];
        __syncwarp();
        const bool use_tab = tab > 0 && nk1 - K <= tab;
        if (use_tab) {
            for (int li = lane; li < nk1 - K; li += 32) {   // interval l = K1 + li
                double twi[2 * K];
#pragma unroll
                for (int q = 0; q < 2 * K; ++q) twi[q] = ts[li + 1 + q];   // t(l-K+1+q), 0-based l-K+q = li+1+q
                SpanTab<K, 1> sp;
                sp.rs = spans + (size_t)li * NS;
                span_init<K, 1>(sp, twi);
                spmask[li] = sp.special;
            }
            __syncwarp();
        }
        const double tb = ts[K1 - 1], te = ts[nk1];
        int lbase = K1;  // interval (1-based) the previous step's last point fell in
        for (int i0 = 0; i0 < m; i0 += 32) {
            const int i = i0 + lane;
            const bool act = i < m;
            double arg = act ? uj[i] : te;
            if (arg < tb) arg = tb;  // min(max(u,tb),te)
            if (arg > te) arg = te;
            int l = lbase;
            while (l < nk1 && arg >= ts[l]) ++l;  // ts[l] = t(l+1)
            lbase = __shfl_sync(FULLW, l, min(31, m - 1 - i0));
            if (act) {
                double tw[2 * K], h[K1];
#pragma unroll
                for (int q = 0; q < 2 * K; ++q) tw[q] = ts[l - K + q];  // t(l-K+1+q)
                if (use_tab) {
                    SpanTab<K, 1> sp;
                    sp.rs = spans + (size_t)(l - K1) * NS;
                    sp.special = spmask[l - K1];
                    bspl_tab<K, 1>(tw, sp, arg, h);
                } else {
                    bspl<K>(tw, arg, h);
                }
                const double *cl = cs + (l - K1);
                for (int d = 0; d < idim; ++d) {
                    double sum = 0.0;
#pragma unroll
                    for (int q = 0; q < K1; ++q) sum = add(sum, mul(h[q], cl[q]));
                    xj[(long long)i * idim + d] = sum;
                    cl += smem_ld;
                }
            }
        }
    }
}

// fast path: curves whose parameters are not non-decreasing (ier = 10, nothing written, core:1145) get a zero
// knot count, which makes the evaluation kernel flag and skip them
__global__ void curev_check_kernel(int nspl, int m, const double *__restrict__ u, const int *__restrict__ n,
                                   int *__restrict__ n_eff)
{
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long j = warp; j < nspl; j += nwarps) {
        const double *uj = u + j * (long long)m;
        bool bad = false;
        for (int i = lane + 1; i < m; i += 32) bad = bad || (uj[i] < uj[i - 1]);
        bad = __any_sync(FULLW, bad);
        if (lane == 0) n_eff[j] = bad ? 0 : n[j];
    }
}

template <int K>
cudaError_t launch_curev_k(int nspl, int idim, const double *t, const int *n, const double *c, int ldt,
                           long long ldc, const double *u, double *x, int m, int *ier, int nsm, cudaStream_t st)
{
    const int smem_ld = (ldt + 1) & ~1;
    // 8 warps per CTA; curves with many knots x coordinates get fewer warps so that their tables still
    // fit the 200 KB a CTA may claim (one warp: n*(1+idim) <= 25 600 doubles).
    const int nintmax = smem_ld > 2 * K + 2 ? smem_ld - 2 * K - 1 : 1;
    const size_t base_d = (size_t)smem_ld * (1 + idim);
    static const bool no_tab = std::getenv("FPB_CUREV_NO_TAB") != nullptr;
    // span table for up to 96 knot intervals per curve -- only where it does not cost resident CTAs (measured on
    // the bench shape, ld = 260: the table halves the residency and evaluation drops from 40 to 34 G points/s)
    int tab = no_tab ? 0 : (nintmax < 96 ? nintmax : 96);
    auto resident = [&](size_t doubles) {
        const size_t b = doubles * sizeof(double) * kCurevWarps;
        const int r = b > 0 ? (int)((size_t)200 * 1024 / b) : 8;
        return r > 8 ? 8 : r;
    };
    if (tab && resident(base_d + (size_t)(K * (K + 1) / 2) * tab + (tab + 1) / 2) != resident(base_d)) tab = 0;
    const size_t per_warp = (base_d + (tab ? (size_t)(K * (K + 1) / 2) * tab + (tab + 1) / 2 : 0)) * sizeof(double);
    int warps = kCurevWarps;
    size_t smem = (size_t)warps * per_warp;
    while (smem > 200u * 1024u && warps > 1) {
        warps >>= 1;
        smem = (size_t)warps * per_warp;
    }
    if (smem > 200u * 1024u) {
        set_last_error("curev: knot/coefficient tables of one curve exceed 200 KB of shared memory (n*(1+idim) > 25600)");
        return cudaErrorInvalidValue;
    }
    auto kern = curev_kernel<K>;
    if (smem > 48u * 1024u) {
        const cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    int per_sm = smem > 0 ? (int)((size_t)200 * 1024 / smem) : 8;
    per_sm = per_sm < 1 ? 1 : (per_sm > 8 ? 8 : per_sm);
    long long blocks = ((long long)nspl + warps - 1) / warps;
    blocks = std::min<long long>(blocks, (long long)nsm * per_sm);
    kern<<<(int)blocks, warps * 32, smem, st>>>(nspl, idim, t, n, c, ldt, ldc, u, x, m, ier, smem_ld, tab);
    return cudaGetLastError();
}

cudaError_t launch_curev(int nspl, int idim, const double *t, const int *n, const double *c, int ldt,
                         long long ldc, int k, const double *u, double *x, int m, int *ier, int nsm,
                         cudaStream_t st)
{
    switch (k) {
    case 1: return launch_curev_k<1>(nspl, idim, t, n, c, ldt, ldc, u, x, m, ier, nsm, st);
    case 2: return launch_curev_k<2>(nspl, idim, t, n, c, ldt, ldc, u, x, m, ier, nsm, st);
    case 3: return launch_curev_k<3>(nspl, idim, t, n, c, ldt, ldc, u, x, m, ier, nsm, st);
    case 4: return launch_curev_k<4>(nspl, idim, t, n, c, ldt, ldc, u, x, m, ier, nsm, st);
    case 5: return launch_curev_k<5>(nspl, idim, t, n, c, ldt, ldc, u, x, m, ier, nsm, st);
    default: return cudaErrorInvalidValue;
    }
}

struct Guard {
    int prev = -1;
    bool ok = false;
    explicit Guard(int dev)
    {
        ok = cudaGetDevice(&prev) == cudaSuccess && cudaSetDevice(dev) == cudaSuccess;
    }
    ~Guard()
    {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

}  // namespace
}  // namespace fpb

using namespace fpb;

extern "C" {

int32_t fp_parcur_batch_dev_c(fpb_engine *eng, FP_SIZE nbatch, FP_SIZE iopt, FP_SIZE ipar, FP_SIZE idim,
                              FP_SIZE m, FP_REAL *u, int64_t u_sb, int64_t u_si, const FP_REAL *x,
                              int64_t x_sb, int64_t x_si, int64_t x_sd, const FP_REAL *w, int64_t w_sb,
                              int64_t w_si, const FP_REAL *ub, const FP_REAL *ue, FP_SIZE ube_stride,
                              FP_SIZE k, const FP_REAL *s, FP_SIZE s_stride, FP_SIZE nest, FP_SIZE *n,
                              FP_REAL *t, FP_REAL *c, FP_REAL *fp, FP_FLAG *ier, FP_REAL *fpint,
                              FP_SIZE *nrdata, FP_SIZE *stats)
{
    if (!eng) {
        set_last_error("null engine");
        return FPB_ERR_ARG;
    }
    if (nbatch <= 0) return FPB_STATUS_OK;
    if (!u || !x || !s || !n || !t || !c || !fp || !ier || (ub && !ue) || (!ub && ue)) {
        set_last_error("fp_parcur_batch_dev_c: null required pointer");
        return FPB_ERR_ARG;
    }
    if (iopt == 1 && (!fpint || !nrdata)) {
        set_last_error("fp_parcur_batch_dev_c: iopt=1 needs the fpint/nrdata state arrays");
        return FPB_ERR_ARG;
    }
    Guard g(eng->device);
    if (!g.ok) return FPB_ERR_CUDA;
    // call-level part of parcur's data check (core:15221-15232): every curve gets ier=10
    if (iopt < -1 || iopt > 1 || ipar < 0 || ipar > 1 || idim <= 0 || idim > kMaxIdim || k <= 0 || k > 5 ||
        m < k + 1 || nest < 2 * (k + 1)) {
        if (!cuda_ok(launch_fill_i32(ier, FPB_INPUT_ERROR, nbatch, eng->stream), "fill ier"))
            return FPB_ERR_CUDA;
        eng->launches++;
        return FPB_STATUS_OK;
    }
    const bool chord = (ipar == 0 && iopt <= 0);
    if (chord) {  // u(i) = normalised cumulative chord length, ub = 0, ue = 1 (core:15238-15253)
        if (!cuda_ok(launch_parcur_param(nbatch, m, idim, x, x_sb, x_si, x_sd, u, u_sb, u_si, eng->stream),
                     "parcur parameter kernel"))
            return FPB_ERR_CUDA;
        eng->launches++;
    }
    FitParams p;
    std::memset(&p, 0, sizeof p);
    p.nbatch = nbatch; p.iopt = iopt; p.m = m; p.k = k; p.nest = nest;
    p.idim = idim; p.par = 1;
    p.x = u; p.x_sb = u_sb; p.x_si = u_si;
    p.y = x; p.y_sb = x_sb; p.y_si = x_si; p.y_sd = x_sd;
    p.w = w; p.w_sb = w_sb; p.w_si = w_si;
    // chord-length parameters run over [0,1] exactly (u(1)=0, u(m)=1): the data range
    p.xb = chord ? nullptr : ub; p.xe = chord ? nullptr : ue; p.xbe_stride = ube_stride;
    p.s = s; p.s_stride = s_stride;
    p.n = n; p.t = t; p.c = c; p.fp = fp; p.ier = ier;
    p.fpint = fpint; p.nrdata = nrdata; p.stats = stats;
    p.c_sb = (long long)nest * idim;
    return run_fit_pipeline(eng, p);
}

int32_t fp_curev_batch_dev_c(fpb_engine *eng, FP_SIZE nspl, FP_SIZE idim, const FP_REAL *t,
                             const FP_SIZE *n, const FP_REAL *c, FP_SIZE ldt, int64_t ldc, FP_SIZE k,
                             const FP_REAL *u, FP_REAL *x, FP_SIZE m, FP_FLAG *ier)
{
    if (!eng || !t || !n || !c || !u || !x || !ier) {
        set_last_error("fp_curev_batch_dev_c: null required pointer");
        return FPB_ERR_ARG;
    }
    if (nspl <= 0) return FPB_STATUS_OK;
    Guard g(eng->device);
    if (!g.ok) return FPB_ERR_CUDA;
    if (m < 1 || idim < 1 || idim > kMaxIdim || k < 1 || k > 5) {  // core:1142
        if (!cuda_ok(launch_fill_i32(ier, FPB_INPUT_ERROR, nspl, eng->stream), "fill ier")) return FPB_ERR_CUDA;
        eng->launches++;
        return FPB_STATUS_OK;
    }
    if (eng->curev_fast) {
        // FAST: every coordinate is a spline on the shared knots -- one pass of the fast splev kernel per
        // coordinate (coefficients at c + d*n, results interleaved into x(idim,m)); arguments are clamped to
        // [t(k+1), t(n-k)] as curev does (core:1163-1165 = splev's e = 3)
        if (!eng->curev_n.reserve((size_t)nspl * sizeof(int))) return FPB_ERR_ALLOC;
        int *n_eff = eng->curev_n.as<int>();
        long long blocks = ((long long)nspl + 7) / 8;
        blocks = std::min<long long>(blocks, (long long)eng->nsm * 8);
        curev_check_kernel<<<(int)blocks, 256, 0, eng->stream>>>(nspl, m, u, n, n_eff);
        if (!cuda_ok(cudaGetLastError(), "curev check kernel")) return FPB_ERR_CUDA;
        eng->launches++;
        for (int d = 0; d < idim; ++d) {
            SplevParams sp;
            sp.nspl = nspl; sp.ldt = ldt; sp.k = k; sp.m = m; sp.e = 3; sp.mode = 1; sp.nshared = 0;
            sp.t = t; sp.c = c; sp.x = u; sp.n = n_eff; sp.y = x; sp.ier = ier; sp.lidx = nullptr;
            sp.ldc = ldc; sp.c_mul = d; sp.y_stride = idim; sp.y_off = d;
            if (!cuda_ok(launch_splev(sp, eng->nsm, eng->stream), "curev fast kernel")) return FPB_ERR_CUDA;
            eng->launches++;
        }
        return FPB_STATUS_OK;
    }
    if (!cuda_ok(launch_curev(nspl, idim, t, n, c, ldt, ldc, k, u, x, m, ier, eng->nsm, eng->stream),
                 "curev kernel"))
        return FPB_ERR_CUDA;
    eng->launches++;
    return FPB_STATUS_OK;
}

// Host-pointer batch: x [nbatch][m][idim] (= Fortran x(idim,m,nbatch)), u [nbatch][m] in/out,
// w [nbatch][m] or NULL (unit weights), ub/ue [nbatch] in/out, t [nbatch][nest] (in for iopt != 0),
// c [nbatch][nest*idim], fpint/nrdata [nbatch][nest] (the wrk(1:nest)/iwrk state of the
// reference; may be NULL when iopt = 0 and no continuation is wanted).
int32_t fp_parcur_batch_c(FP_SIZE nbatch, FP_SIZE iopt, FP_SIZE ipar, FP_SIZE idim, FP_SIZE m, FP_REAL *u,
                          const FP_REAL *x, const FP_REAL *w, FP_REAL *ub, FP_REAL *ue, FP_SIZE k,
                          FP_REAL s, FP_SIZE nest, FP_SIZE *n, FP_REAL *t, FP_REAL *c, FP_REAL *fp,
                          FP_FLAG *ier, FP_REAL *fpint, FP_SIZE *nrdata, FP_SIZE ngpus)
{
    if (nbatch <= 0) return FPB_STATUS_OK;
    if (!u || !x || !ub || !ue || !n || !t || !c || !fp || !ier) {
        set_last_error("fp_parcur_batch_c: null required pointer");
        return FPB_ERR_ARG;
    }
    if (iopt < -1 || iopt > 1 || ipar < 0 || ipar > 1 || idim <= 0 || idim > kMaxIdim || k <= 0 || k > 5 ||
        m < k + 1 || nest < 2 * (k + 1)) {
        for (FP_SIZE b = 0; b < nbatch; ++b) ier[b] = FPB_INPUT_ERROR;
        return FPB_STATUS_OK;
    }
    if (iopt == 1 && (!fpint || !nrdata)) {
        set_last_error("fp_parcur_batch_c: iopt=1 needs the fpint/nrdata state arrays");
        return FPB_ERR_ARG;
    }
    // contiguous slices of the batch per GPU, one host thread + engine per device (no collective)
    return shard_units_over_gpus(nbatch, ngpus, 128, [&](fpb_engine *eng, long long s0, long long s1) -> int32_t {
    {   // a few curves (fp_parcur_c: one): the warp-per-curve kernel
        int32_t status = FPB_STATUS_OK;
        const size_t md0 = (size_t)m * idim;
        if (parcur_solo_host(eng, s1 - s0, iopt, ipar, idim, m, u + (size_t)s0 * m, x + (size_t)s0 * md0,
                             w ? w + (size_t)s0 * m : nullptr, ub + s0, ue + s0, k, s, nest, n + s0, t + (size_t)s0 * nest,
                             c + (size_t)s0 * nest * idim, fp + s0, ier + s0, fpint ? fpint + (size_t)s0 * nest : nullptr,
                             nrdata ? nrdata + (size_t)s0 * nest : nullptr, &status))
            return status;
    }
    Guard g(eng->device);
    if (!g.ok) return FPB_ERR_CUDA;
    cudaStream_t st = eng->stream;
    const bool chord = (ipar == 0 && iopt <= 0);
    const size_t per_curve = (size_t)m * (idim + 2) * 16 + (size_t)nest * (idim + 3) * 8 + 64;
    const long long chunk = std::max<long long>(128, (long long)(((size_t)6 << 30) / per_curve) / 128 * 128);
    enum { XC, XP, UC, UP, WC, WP, SS, BE, NN, TT, CC, FP_IER };
    DevBuf *B = eng->scratch;
    for (long long b0 = s0; b0 < s1; b0 += chunk) {
        const long long nb = std::min<long long>(chunk, s1 - b0);
        const long long ld = (nb + 31) / 32 * 32;  // point-major leading dimension
        const size_t md = (size_t)m * idim;
        // TT holds t [nb][nest] doubles, fpint [nb][nest] doubles and nrdata [nb][nest] ints
        if (!B[XC].reserve(nb * md * 8) || !B[XP].reserve(ld * md * 8) || !B[UC].reserve((size_t)nb * m * 8) ||
            !B[UP].reserve((size_t)ld * m * 8) || !B[WC].reserve((size_t)nb * m * 8) ||
            !B[WP].reserve((size_t)ld * m * 8) || !B[SS].reserve(64) || !B[BE].reserve((size_t)nb * 16) ||
            !B[NN].reserve((size_t)nb * 8) || !B[TT].reserve((size_t)nb * nest * 20) ||
            !B[CC].reserve((size_t)nb * nest * idim * 8) || !B[FP_IER].reserve((size_t)nb * 16))
            return FPB_ERR_ALLOC;
        double *d_xc = B[XC].as<double>(), *d_xp = B[XP].as<double>(), *d_uc = B[UC].as<double>(),
               *d_up = B[UP].as<double>(), *d_wc = B[WC].as<double>(), *d_wp = B[WP].as<double>(),
               *d_s = B[SS].as<double>(), *d_ub = B[BE].as<double>(), *d_ue = d_ub + nb,
               *d_t = B[TT].as<double>(), *d_fpint = d_t + (size_t)nb * nest, *d_c = B[CC].as<double>(),
               *d_fp = B[FP_IER].as<double>();
        int *d_n = B[NN].as<int>(), *d_ier = reinterpret_cast<int *>(d_fp + nb),
            *d_nrd = reinterpret_cast<int *>(d_fpint + (size_t)nb * nest);
        const bool keep = fpint && nrdata;
        bool ok = cuda_ok(cudaMemcpyAsync(d_xc, x + (size_t)b0 * md, nb * md * 8, cudaMemcpyHostToDevice, st), "H2D x");
        if (!chord)
            ok = ok && cuda_ok(cudaMemcpyAsync(d_uc, u + (size_t)b0 * m, (size_t)nb * m * 8, cudaMemcpyHostToDevice, st), "H2D u");
        if (w) ok = ok && cuda_ok(cudaMemcpyAsync(d_wc, w + (size_t)b0 * m, (size_t)nb * m * 8, cudaMemcpyHostToDevice, st), "H2D w");
        ok = ok && cuda_ok(cudaMemcpyAsync(d_s, &s, 8, cudaMemcpyHostToDevice, st), "H2D s");
        if (!chord) {
            ok = ok && cuda_ok(cudaMemcpyAsync(d_ub, ub + b0, (size_t)nb * 8, cudaMemcpyHostToDevice, st), "H2D ub");
            ok = ok && cuda_ok(cudaMemcpyAsync(d_ue, ue + b0, (size_t)nb * 8, cudaMemcpyHostToDevice, st), "H2D ue");
        }
        // n, t, c are inputs for iopt != 0 -- and for every iopt the caller's values must survive a curve
        // rejected with ier=10 (the pass kernel writes nothing but ier for it, the reference leaves them
        // untouched, core:15255-15262), so they always round-trip through the device
        ok = ok && cuda_ok(cudaMemcpyAsync(d_n, n + b0, (size_t)nb * 4, cudaMemcpyHostToDevice, st), "H2D n");
        ok = ok && cuda_ok(cudaMemcpyAsync(d_t, t + (size_t)b0 * nest, (size_t)nb * nest * 8, cudaMemcpyHostToDevice, st), "H2D t");
        ok = ok && cuda_ok(cudaMemcpyAsync(d_c, c + (size_t)b0 * nest * idim, (size_t)nb * nest * idim * 8, cudaMemcpyHostToDevice, st), "H2D c");
        if (iopt == 1) {
            ok = ok && cuda_ok(cudaMemcpyAsync(d_fpint, fpint + (size_t)b0 * nest, (size_t)nb * nest * 8, cudaMemcpyHostToDevice, st), "H2D fpint");
            ok = ok && cuda_ok(cudaMemcpyAsync(d_nrd, nrdata + (size_t)b0 * nest, (size_t)nb * nest * 4, cudaMemcpyHostToDevice, st), "H2D nrdata");
        }
        if (!ok) return FPB_ERR_CUDA;
        // curve-major -> point-major: element (i,d) of curve b at d_xp[(i*idim+d)*ld + b]
        ok = cuda_ok(launch_transpose(d_xc, (long long)md, d_xp, ld, nb, (long long)md, st), "transpose x");
        if (!chord) ok = ok && cuda_ok(launch_transpose(d_uc, m, d_up, ld, nb, m, st), "transpose u");
        if (w) ok = ok && cuda_ok(launch_transpose(d_wc, m, d_wp, ld, nb, m, st), "transpose w");
        if (!ok) return FPB_ERR_CUDA;
        eng->launches += 1 + (chord ? 0 : 1) + (w ? 1 : 0);
        const int32_t rc = fp_parcur_batch_dev_c(eng, (FP_SIZE)nb, iopt, ipar, idim, m, d_up, 1, ld, d_xp, 1,
                                                 (int64_t)idim * ld, ld, w ? d_wp : nullptr, 1, ld,
                                                 chord ? nullptr : d_ub, chord ? nullptr : d_ue, 1, k, d_s, 0,
                                                 nest, d_n, d_t, d_c, d_fp, d_ier, keep ? d_fpint : nullptr,
                                                 keep ? d_nrd : nullptr, nullptr);
        if (rc != FPB_STATUS_OK) return rc;
        if (chord) {  // the parameters are an output
            ok = cuda_ok(launch_transpose(d_up, ld, d_uc, m, m, nb, st), "transpose u back");
            eng->launches++;
            ok = ok && cuda_ok(cudaMemcpyAsync(u + (size_t)b0 * m, d_uc, (size_t)nb * m * 8, cudaMemcpyDeviceToHost, st), "D2H u");
            if (!ok) return FPB_ERR_CUDA;
        }
        ok = cuda_ok(cudaMemcpyAsync(n + b0, d_n, (size_t)nb * 4, cudaMemcpyDeviceToHost, st), "D2H n");
        ok = ok && cuda_ok(cudaMemcpyAsync(t + (size_t)b0 * nest, d_t, (size_t)nb * nest * 8, cudaMemcpyDeviceToHost, st), "D2H t");
        ok = ok && cuda_ok(cudaMemcpyAsync(c + (size_t)b0 * nest * idim, d_c, (size_t)nb * nest * idim * 8, cudaMemcpyDeviceToHost, st), "D2H c");
        std::vector<double> hfp((size_t)nb);
        std::vector<int> hier((size_t)nb);
        ok = ok && cuda_ok(cudaMemcpyAsync(hfp.data(), d_fp, (size_t)nb * 8, cudaMemcpyDeviceToHost, st), "D2H fp");
        ok = ok && cuda_ok(cudaMemcpyAsync(hier.data(), d_ier, (size_t)nb * 4, cudaMemcpyDeviceToHost, st), "D2H ier");
        if (keep && iopt >= 0) {
            ok = ok && cuda_ok(cudaMemcpyAsync(fpint + (size_t)b0 * nest, d_fpint, (size_t)nb * nest * 8, cudaMemcpyDeviceToHost, st), "D2H fpint");
            ok = ok && cuda_ok(cudaMemcpyAsync(nrdata + (size_t)b0 * nest, d_nrd, (size_t)nb * nest * 4, cudaMemcpyDeviceToHost, st), "D2H nrdata");
        }
        ok = ok && cuda_ok(cudaStreamSynchronize(st), "sync");
        if (!ok) return FPB_ERR_CUDA;
        for (long long b = 0; b < nb; ++b) {
            ier[b0 + b] = hier[(size_t)b];
            // fp is an output only where fppara ran (the reference leaves it untouched on ier=10)
            if (hier[(size_t)b] != FPB_INPUT_ERROR) fp[b0 + b] = hfp[(size_t)b];
            if (chord) {
                ub[b0 + b] = 0.0;
                ue[b0 + b] = 1.0;
            }
        }
    }
    return FPB_STATUS_OK;
    });
}

// ---- the reference's own flat entry points: a batch of one curve ----

void fp_parcur_c(FP_SIZE iopt, FP_SIZE ipar, FP_SIZE idim, FP_SIZE m, FP_REAL *u, FP_SIZE mx,
                 const FP_REAL *x, FP_REAL *w, FP_REAL *ub, FP_REAL *ue, FP_SIZE k, FP_REAL *s,
                 FP_SIZE nest, FP_SIZE *n, FP_REAL *t, FP_SIZE nc, FP_REAL *c, FP_REAL *fp, FP_REAL *wrk,
                 FP_SIZE lwrk, FP_SIZE *iwrk, FP_FLAG *ier)
{
    // data check that does not need the data, core:15221-15236
    *ier = FPB_INPUT_ERROR;
    if (iopt < -1 || iopt > 1) return;
    if (ipar < 0 || ipar > 1) return;
    if (idim <= 0 || idim > kMaxIdim) return;
    if (k <= 0 || k > 5) return;
    if (m < k + 1 || nest < 2 * (k + 1)) return;
    if (mx < m * idim || nc < nest * idim) return;
    if (lwrk < m * (k + 1) + nest * (6 + idim + 3 * k)) return;
    // wrk(1:nest) is fpint and iwrk(1:nest) is nrdata (core:15275-15283): the iopt=1 state
    FP_FLAG ier1 = FPB_INPUT_ERROR;
    const int32_t rc = fp_parcur_batch_c(1, iopt, ipar, idim, m, u, x, w, ub, ue, k, *s, nest, n, t, c, fp,
                                         &ier1, wrk, iwrk, 1);
    *ier = (rc != FPB_STATUS_OK) ? rc : ier1;
}

void fp_curev_c(FP_SIZE idim, const FP_REAL *t, FP_SIZE n, const FP_REAL *c, FP_SIZE nc, FP_SIZE k,
                const FP_REAL *u, FP_SIZE m, FP_REAL *x, FP_SIZE mx, FP_FLAG *ier)
{
    *ier = FPB_INPUT_ERROR;
    if (m < 1) return;                 // core:1142
    if (mx < m * idim) return;         // core:1148
    if (idim < 1 || idim > kMaxIdim || k < 1 || k > 5 || n < 2 * (k + 1) || nc < n * (idim - 1) + n - k - 1) return;
    for (FP_SIZE i = 1; i < m; ++i)    // any(u(2:m) < u(1:m-1)), core:1145 (checked here: the points are
        if (u[i] < u[i - 1]) return;   // spread over many warps below, each sees only its own chunk)
    HostApiLock api_lock(-1);  // per-device serialisation (fpb_engine.h)
    fpb_engine *eng = default_engine(-1);
    if (!eng) {
        *ier = FPB_ERR_CUDA;
        return;
    }
    Guard g(eng->device);
    cudaStream_t st = eng->stream;
    DevBuf *B = eng->scratch;
    // one curve, its m points spread over the device: chunks of `chunk` points, every chunk a "curve"
    // of the batch kernel with its own copy of the (small) knot and coefficient tables
    const int chunk = m > 8192 ? 2048 : m;
    const long long nchunks = ((long long)m + chunk - 1) / chunk;
    const size_t padded = (size_t)nchunks * chunk;
    const size_t ncu = (size_t)n * (idim - 1) + (size_t)(n - k - 1);  // coefficients actually read
    const size_t ldc = (ncu + 1) & ~(size_t)1;
    const int ldt = (n + 1) & ~1;
    std::vector<double> ht((size_t)nchunks * ldt, 0.0), hc((size_t)nchunks * ldc, 0.0);
    std::vector<int> hn((size_t)nchunks, n);
    for (long long q = 0; q < nchunks; ++q) {
        std::memcpy(&ht[(size_t)q * ldt], t, (size_t)n * 8);
        std::memcpy(&hc[(size_t)q * ldc], c, ncu * 8);
    }
    if (!B[0].reserve(ht.size() * 8) || !B[1].reserve(hc.size() * 8) || !B[2].reserve((size_t)nchunks * 8) ||
        !B[3].reserve(padded * 8) || !B[4].reserve(padded * idim * 8)) {
        *ier = FPB_ERR_ALLOC;
        return;
    }
    int *d_n = B[2].as<int>(), *d_ier = d_n + nchunks;
    bool ok = cuda_ok(cudaMemcpyAsync(B[0].ptr, ht.data(), ht.size() * 8, cudaMemcpyHostToDevice, st), "H2D");
    ok = ok && cuda_ok(cudaMemcpyAsync(B[1].ptr, hc.data(), hc.size() * 8, cudaMemcpyHostToDevice, st), "H2D");
    ok = ok && cuda_ok(cudaMemcpyAsync(d_n, hn.data(), (size_t)nchunks * 4, cudaMemcpyHostToDevice, st), "H2D");
    ok = ok && cuda_ok(cudaMemcpyAsync(B[3].ptr, u, (size_t)m * 8, cudaMemcpyHostToDevice, st), "H2D");
    std::vector<double> pad;
    if (ok && padded > (size_t)m) {  // the tail chunk repeats the last parameter (keeps it monotone)
        pad.assign(padded - m, u[m - 1]);
        ok = cuda_ok(cudaMemcpyAsync(B[3].as<double>() + m, pad.data(), pad.size() * 8, cudaMemcpyHostToDevice, st),
                     "H2D pad");
    }
    ok = ok && cuda_ok(cudaStreamSynchronize(st), "sync");  // the host vectors go out of use here
    if (!ok) {
        *ier = FPB_ERR_CUDA;
        return;
    }
    const int32_t rc = fp_curev_batch_dev_c(eng, (FP_SIZE)nchunks, idim, B[0].as<double>(), d_n, B[1].as<double>(),
                                            ldt, (int64_t)ldc, k, B[3].as<double>(), B[4].as<double>(), chunk, d_ier);
    if (rc != FPB_STATUS_OK) {
        *ier = rc;
        return;
    }
    std::vector<int> hier((size_t)nchunks, FPB_INPUT_ERROR);
    ok = cuda_ok(cudaMemcpyAsync(hier.data(), d_ier, (size_t)nchunks * 4, cudaMemcpyDeviceToHost, st), "D2H") &&
         cuda_ok(cudaStreamSynchronize(st), "sync");
    FP_FLAG h = FPB_OK;
    for (long long q = 0; ok && q < nchunks; ++q)
        if (hier[(size_t)q] != FPB_OK) h = hier[(size_t)q];
    if (ok && h == FPB_OK)
        ok = cuda_ok(cudaMemcpy(x, B[4].ptr, (size_t)m * idim * 8, cudaMemcpyDeviceToHost), "D2H x");
    *ier = ok ? h : FPB_ERR_CUDA;
}

}  // extern "C"
