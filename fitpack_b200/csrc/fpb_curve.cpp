// fpb_curve.cpp -- host-side mirror of the reference's `fitpack_curve` type and of its
// opaque-handle C-ABI `fitpack_curve_c`.
//
// Reference: type fitpack_fitter  src/fitpack_fitters.f90:37-77, 153-164
//            type fitpack_curve   src/fitpack_curves.f90:54-131
//              new_points :215-263, new_fit :168-179, fit :431-488, eval :273-393,
//              interpolate :396-408, least_squares :411-428, destroy :182-201
//            C handle             src/fitpack_curves_c.f90:56-396
//            get_smoothing        src/fitpack_core.F90:351-370
//            fitpack_argsort      src/fitpack_core.F90:18486-18590
//
// The reference's host code is Modern Fortran; this image has no Fortran compiler, so
// the host layer above the C-ABI is written in C++ (the Fortran iso_c_binding interface
// module that binds the same symbols is shipped as source under fortran/).  All numerical
// work goes through fp_curfit_c / fp_splev_c, i.e. the CUDA kernels, as a batch of one.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <new>
#include <vector>

#include "../../include/fitpack_b200.h"

namespace {

constexpr int kMaxK = 5;  // MAX_K, fitpack_curves.f90:38

bool fp_equal(double a, double b)  // equal(), core:18644-18647 (== |a-b| < tiny for binary64)
{
    return std::fabs(a - b) < std::numeric_limits<double>::min();
}

// toBeSwapped(a,b,orEqual) for an ascending sort, core:18575-18583
bool to_be_swapped(double a, double b, bool or_equal)
{
    bool r = a > b;
    if (or_equal && fp_equal(a, b)) r = true;
    return r;
}

// fitpack_quicksort with index tracking, core:18504-18573: interchange sort up to 8
// elements, otherwise partition around the middle element.  `list`/`idx` are 1-based views.
void quicksort(double *list, int *idx, int n)
{
    double *L = list - 1;
    int *I = idx - 1;
    if (n <= 8) {
        for (int i = 1; i <= n - 1; ++i)
            for (int j = i + 1; j <= n; ++j)
                if (to_be_swapped(L[i], L[j], false)) {
                    std::swap(L[i], L[j]);
                    std::swap(I[i], I[j]);
                }
        return;
    }
    const double chosen = L[n / 2];
    int i = 0, j = n + 1;
    for (;;) {
        do ++i; while (!(to_be_swapped(L[i], chosen, true) || i == n));
        do --j; while (!(to_be_swapped(chosen, L[j], true) || j == 1));
        if (i < j) {
            std::swap(L[i], L[j]);
            std::swap(I[i], I[j]);
        } else {
            if (i == j) ++i;
            break;
        }
    }
    if (1 < j) quicksort(list, idx, j);
    if (i < n) quicksort(list + (i - 1), idx + (i - 1), n - i + 1);
}

std::vector<int> argsort(const double *x, int n)
{
    std::vector<double> copy(x, x + n);
    std::vector<int> idx(n);
    for (int i = 0; i < n; ++i) idx[i] = i;
    if (n > 1) quicksort(copy.data(), idx.data(), n);
    return idx;
}

// state of type(fitpack_curve), defaults as in the reference's declarations
struct Curve {
    // fitpack_fitter
    int iopt = FPB_IOPT_NEW_SMOOTHING;
    double smoothing = 1000.0;
    double fp = 0.0;
    std::vector<double> c;
    int lwrk = 0;
    std::vector<double> wrk;
    int liwrk = 0;
    std::vector<int> iwrk;
    // fitpack_curve
    int m = 0;
    std::vector<double> x, y, sp, w;
    int order = 3;
    double xleft = 0.0, xright = 0.0;
    int nest = 0;
    int bc = FPB_OUTSIDE_NEAREST_BND;
    int knots = 0;
    std::vector<double> t;
    std::vector<double> wrk_fou;  // (nest,2) workspace of fourier_coefficients, :71, :259

    void destroy()  // fitpack_curves.f90:182-201 + fitter_destroy_base
    {
        *this = Curve();
    }

    void new_points(int npts, const double *xi, const double *yi, const double *wi)
    {
        destroy();
        m = npts > 0 ? npts : 0;
        const std::vector<int> isort = argsort(xi, m);
        x.resize(m);
        y.resize(m);
        w.resize(m);
        for (int i = 0; i < m; ++i) {
            x[i] = xi[isort[i]];
            y[i] = yi[isort[i]];
            w[i] = wi ? wi[isort[i]] : 1.0;
        }
        sp.assign(m, 0.0);
        if (m > 0) {
            xleft = xi[0];
            xright = xi[0];
            for (int i = 1; i < m; ++i) {
                if (xi[i] < xleft) xleft = xi[i];
                if (xi[i] > xright) xright = xi[i];
            }
        }
        iopt = FPB_IOPT_NEW_SMOOTHING;
        // nest = max(SAFE*2*MAX_K+2, ceiling(1.4*m)) with a default-REAL (binary32) product
        const float est = 1.4f * static_cast<float>(m);
        nest = std::max(10 * 2 * kMaxK + 2, static_cast<int>(std::ceil(est)));
        liwrk = nest;
        iwrk.assign(nest, 0);
        t.assign(nest, 0.0);
        c.assign(nest, 0.0);
        lwrk = m * (kMaxK + 1) + nest * (8 + 5 * kMaxK);
        wrk.assign(lwrk, 0.0);
        wrk_fou.assign(static_cast<size_t>(nest) * 2, 0.0);
    }

    // curve_fit_automatic_knots, fitpack_curves.f90:431-488
    int fit(const double *user_smoothing, const int *user_order, bool keep_knots)
    {
        // get_smoothing, core:351-370
        int nit;
        double smooth_now[3];
        if (user_smoothing) {
            nit = 1;
            smooth_now[0] = smooth_now[1] = smooth_now[2] = *user_smoothing;
        } else if (smoothing < 1000.0) {
            nit = 1;
            smooth_now[0] = smooth_now[1] = smooth_now[2] = smoothing;
        } else {
            nit = 3;
            smooth_now[0] = 1000.0;
            smooth_now[1] = 60.0;
            smooth_now[2] = 30.0;
        }
        if (!keep_knots && iopt == FPB_IOPT_OLD_FIT) iopt = FPB_IOPT_NEW_SMOOTHING;
        if (user_order) order = *user_order;
        int ierr = FPB_INPUT_ERROR;
        for (int loop = 0; loop < nit; ++loop) {
            smoothing = smooth_now[loop];
            fp_curfit_c(iopt, m, x.data(), y.data(), w.data(), xleft, xright, order, smoothing, nest,
                        &knots, t.data(), c.data(), &fp, wrk.data(), lwrk, iwrk.data(), &ierr);
            if (FITPACK_SUCCESS_c(ierr)) iopt = FPB_IOPT_OLD_FIT;
        }
        return ierr;
    }

    int interpolate(const int *user_order, bool reset_knots)  // :396-408
    {
        if (reset_knots) iopt = FPB_IOPT_NEW_SMOOTHING;
        const double zero = 0.0;
        return fit(&zero, user_order, !reset_knots);
    }

    int least_squares(const double *user_smoothing, bool reset_knots)  // :411-428
    {
        if (reset_knots) {
            iopt = FPB_IOPT_NEW_SMOOTHING;
            const int ierr = fit(user_smoothing, nullptr, false);
            if (!FITPACK_SUCCESS_c(ierr)) return ierr;
        }
        iopt = FPB_IOPT_NEW_LEASTSQUARES;
        return fit(nullptr, nullptr, false);
    }

    int eval(int npts, const double *xe, double *ye) const  // curve_eval_many :304-347
    {
        int ier = FPB_INPUT_ERROR;
        if (knots < 2 * (order + 1) || static_cast<int>(t.size()) < knots) return ier;
        fp_splev_c(t.data(), knots, c.data(), order, xe, ye, npts, bc, &ier);
        return ier;
    }
};

Curve *get(const fitpack_curve_c *h) { return h ? static_cast<Curve *>(h->cptr) : nullptr; }

// fitpack_curve_c_pointer, fitpack_curves_c.f90:79-98: every call lazily allocates
Curve *ensure(fitpack_curve_c *h)
{
    if (!h) return nullptr;
    if (!h->cptr) h->cptr = new (std::nothrow) Curve();
    return static_cast<Curve *>(h->cptr);
}


// ---- comm wire format (fitter_core_comm_* src/fitpack_fitters.f90:105-147, curve_comm_*
// src/fitpack_curves.f90:821-895, FP_*_COMM_* core:18654-18956): a flat array of doubles;
// scalars stored as doubles; every array = header (one double per dimension holding the
// lower/upper bound as two int32, or -99999999 twice when not allocated) + raw payload
// (int32 payloads padded to a whole double).  An empty vector here = "not allocated".
constexpr int32_t kNotAlloc = -99999999;  // FP_NOT_ALLOC, core:34

size_t comm_size_r(const std::vector<double> &a, int rank = 1) { return rank + a.size(); }
size_t comm_size_i(const std::vector<int> &a) { return 1 + (a.size() + 1) / 2; }

double *pack_bounds(double *buf, int32_t lo, int32_t hi)
{
    const int32_t b[2] = {lo, hi};
    std::memcpy(buf, b, sizeof b);
    return buf + 1;
}

double *pack_r(double *buf, const std::vector<double> &a)
{
    if (a.empty()) return pack_bounds(buf, kNotAlloc, kNotAlloc);
    buf = pack_bounds(buf, 1, static_cast<int32_t>(a.size()));
    std::memcpy(buf, a.data(), a.size() * sizeof(double));
    return buf + a.size();
}

double *pack_r2(double *buf, const std::vector<double> &a, int rows)  // (rows, size/rows), column-major
{
    if (a.empty() || rows <= 0) return pack_bounds(pack_bounds(buf, kNotAlloc, kNotAlloc), kNotAlloc, kNotAlloc);
    buf = pack_bounds(buf, 1, rows);
    buf = pack_bounds(buf, 1, static_cast<int32_t>(a.size() / rows));
    std::memcpy(buf, a.data(), a.size() * sizeof(double));
    return buf + a.size();
}

double *pack_i(double *buf, const std::vector<int> &a)
{
    if (a.empty()) return pack_bounds(buf, kNotAlloc, kNotAlloc);
    buf = pack_bounds(buf, 1, static_cast<int32_t>(a.size()));
    const size_t nd = (a.size() + 1) / 2;
    std::memset(buf, 0, nd * sizeof(double));
    std::memcpy(buf, a.data(), a.size() * sizeof(int));
    return buf + nd;
}

const double *expand_bounds(const double *buf, int32_t &lo, int32_t &hi)
{
    int32_t b[2];
    std::memcpy(b, buf, sizeof b);
    lo = b[0];
    hi = b[1];
    return buf + 1;
}

const double *expand_r(const double *buf, std::vector<double> &a)
{
    int32_t lo, hi;
    buf = expand_bounds(buf, lo, hi);
    a.clear();
    if (lo == kNotAlloc || hi == kNotAlloc) return buf;
    const size_t n = hi >= lo ? static_cast<size_t>(hi - lo + 1) : 0;
    a.assign(buf, buf + n);
    return buf + n;
}

const double *expand_r2(const double *buf, std::vector<double> &a)
{
    int32_t lo1, hi1, lo2, hi2;
    buf = expand_bounds(buf, lo1, hi1);
    buf = expand_bounds(buf, lo2, hi2);
    a.clear();
    if (lo1 == kNotAlloc || hi1 == kNotAlloc || lo2 == kNotAlloc || hi2 == kNotAlloc) return buf;
    const size_t n = static_cast<size_t>(std::max(hi1 - lo1 + 1, 0)) * static_cast<size_t>(std::max(hi2 - lo2 + 1, 0));
    a.assign(buf, buf + n);
    return buf + n;
}

const double *expand_i(const double *buf, std::vector<int> &a)
{
    int32_t lo, hi;
    buf = expand_bounds(buf, lo, hi);
    a.clear();
    if (lo == kNotAlloc || hi == kNotAlloc) return buf;
    const size_t n = hi >= lo ? static_cast<size_t>(hi - lo + 1) : 0;
    a.resize(n);
    std::memcpy(a.data(), buf, n * sizeof(int));
    return buf + (n + 1) / 2;
}

size_t curve_comm_size(const Curve &cv)
{
    return 5 + comm_size_r(cv.c) + comm_size_r(cv.wrk) + comm_size_i(cv.iwrk) + 7 + comm_size_r(cv.x) +
           comm_size_r(cv.y) + comm_size_r(cv.sp) + comm_size_r(cv.w) + comm_size_r(cv.t) +
           comm_size_r(cv.wrk_fou, 2);
}

void curve_comm_pack(const Curve &cv, double *buf)
{
    *buf++ = static_cast<double>(cv.iopt);
    *buf++ = cv.smoothing;
    *buf++ = cv.fp;
    *buf++ = static_cast<double>(cv.lwrk);
    *buf++ = static_cast<double>(cv.liwrk);
    buf = pack_r(buf, cv.c);
    buf = pack_r(buf, cv.wrk);
    buf = pack_i(buf, cv.iwrk);
    *buf++ = static_cast<double>(cv.m);
    *buf++ = static_cast<double>(cv.order);
    *buf++ = static_cast<double>(cv.knots);
    *buf++ = static_cast<double>(cv.bc);
    *buf++ = static_cast<double>(cv.nest);
    *buf++ = cv.xleft;
    *buf++ = cv.xright;
    buf = pack_r(buf, cv.x);
    buf = pack_r(buf, cv.y);
    buf = pack_r(buf, cv.sp);
    buf = pack_r(buf, cv.w);
    buf = pack_r(buf, cv.t);
    pack_r2(buf, cv.wrk_fou, cv.nest);
}

void curve_comm_expand(Curve &cv, const double *buf)
{
    cv.iopt = static_cast<int>(std::lround(*buf++));
    cv.smoothing = *buf++;
    cv.fp = *buf++;
    cv.lwrk = static_cast<int>(std::lround(*buf++));
    cv.liwrk = static_cast<int>(std::lround(*buf++));
    buf = expand_r(buf, cv.c);
    buf = expand_r(buf, cv.wrk);
    buf = expand_i(buf, cv.iwrk);
    cv.m = static_cast<int>(std::lround(*buf++));
    cv.order = static_cast<int>(std::lround(*buf++));
    cv.knots = static_cast<int>(std::lround(*buf++));
    cv.bc = static_cast<int>(std::lround(*buf++));
    cv.nest = static_cast<int>(std::lround(*buf++));
    cv.xleft = *buf++;
    cv.xright = *buf++;
    buf = expand_r(buf, cv.x);
    buf = expand_r(buf, cv.y);
    buf = expand_r(buf, cv.sp);
    buf = expand_r(buf, cv.w);
    buf = expand_r(buf, cv.t);
    expand_r2(buf, cv.wrk_fou);
}

}  // namespace

extern "C" {

// comm_size / comm_pack / comm_expand of type(fitpack_curve) (src/fitpack_curves.f90:821-895): the
// buffers are interchangeable with the ones the reference's Fortran objects produce
FP_SIZE fitpack_curve_c_comm_size(fitpack_curve_c *self)
{
    Curve *cv = ensure(self);
    return cv ? static_cast<FP_SIZE>(curve_comm_size(*cv)) : 0;
}

void fitpack_curve_c_comm_pack(fitpack_curve_c *self, FP_REAL *buffer)
{
    Curve *cv = ensure(self);
    if (cv && buffer) curve_comm_pack(*cv, buffer);
}

void fitpack_curve_c_comm_expand(fitpack_curve_c *self, const FP_REAL *buffer)
{
    Curve *cv = ensure(self);
    if (cv && buffer) curve_comm_expand(*cv, buffer);
}

void fitpack_curve_c_allocate(fitpack_curve_c *self) { (void)ensure(self); }

void fitpack_curve_c_destroy(fitpack_curve_c *self)
{
    if (!self) return;
    delete get(self);
    self->cptr = nullptr;
}

void fitpack_curve_c_copy(fitpack_curve_c *self, const fitpack_curve_c *other)
{
    const Curve *src = get(other);
    if (src) {
        Curve *dst = ensure(self);
        if (dst && dst != src) *dst = *src;
    } else {
        fitpack_curve_c_destroy(self);
    }
}

void fitpack_curve_c_move_alloc(fitpack_curve_c *to, fitpack_curve_c *from)
{
    if (!to || !from) return;
    fitpack_curve_c_destroy(to);
    to->cptr = from->cptr;
    from->cptr = nullptr;
}

void fitpack_curve_c_new_points(fitpack_curve_c *self, FP_SIZE npts, FP_REAL *x, FP_REAL *y,
                                FP_REAL *w)
{
    Curve *cv = ensure(self);
    if (cv) cv->new_points(npts, x, y, w);
}

FP_FLAG fitpack_curve_c_new_fit(fitpack_curve_c *self, FP_SIZE npts, FP_REAL *x, FP_REAL *y,
                                FP_REAL *w, FP_REAL *smoothing)
{
    Curve *cv = ensure(self);
    if (!cv) return FPB_INPUT_ERROR;
    cv->new_points(npts, x, y, w);
    return cv->fit(smoothing, nullptr, false);
}

FP_FLAG fitpack_curve_c_fit(fitpack_curve_c *self, FP_REAL *smoothing, FP_SIZE *order)
{
    Curve *cv = ensure(self);
    return cv ? cv->fit(smoothing, order, false) : FPB_INPUT_ERROR;
}

FP_FLAG fitpack_curve_c_interpolating(fitpack_curve_c *self, FP_SIZE *order)
{
    Curve *cv = ensure(self);
    return cv ? cv->interpolate(order, true) : FPB_INPUT_ERROR;
}

FP_FLAG fitpack_curve_c_least_squares(fitpack_curve_c *self, FP_REAL *smoothing,
                                      FP_BOOL *reset_knots)
{
    Curve *cv = ensure(self);
    return cv ? cv->least_squares(smoothing, reset_knots ? *reset_knots : false) : FPB_INPUT_ERROR;
}

FP_REAL fitpack_curve_c_eval_one(fitpack_curve_c *self, FP_REAL x, FP_FLAG *ierr)
{
    Curve *cv = ensure(self);
    double y = 0.0;
    const int ier = cv ? cv->eval(1, &x, &y) : FPB_INPUT_ERROR;
    if (ierr) *ierr = ier;
    return y;
}

void fitpack_curve_c_eval_many(fitpack_curve_c *self, FP_SIZE npts, FP_REAL *x, FP_REAL *y,
                               FP_FLAG *ierr)
{
    Curve *cv = ensure(self);
    const int ier = cv ? cv->eval(npts, x, y) : FPB_INPUT_ERROR;
    if (ierr) *ierr = ier;
}

// curve_derivative / curve_derivatives, fitpack_curves.f90:499-532, 606-618 (C handle :294-307):
// order < 0 is taken as 0; the curve's bc is passed as splder's extrapolation flag
FP_REAL fitpack_curve_c_derivative(fitpack_curve_c *self, FP_REAL x, FP_SIZE order, FP_SIZE *ierr)
{
    Curve *cv = ensure(self);
    double y = 0.0;
    int ier = FPB_INPUT_ERROR;
    if (cv && cv->knots >= 2 * (cv->order + 1)) {
        std::vector<double> wk(cv->knots);
        fp_splder_c(cv->t.data(), cv->knots, cv->c.data(), cv->order, order > 0 ? order : 0, &x, &y, 1,
                    cv->bc, wk.data(), &ier);
    }
    if (ierr) *ierr = ier;
    return y;
}

// curve_all_derivatives, fitpack_curves.f90:543-565 (C handle :310-323): ddx(1:order+1)
FP_FLAG fitpack_curve_c_all_derivatives(fitpack_curve_c *self, FP_REAL x, FP_REAL *ddx)
{
    Curve *cv = ensure(self);
    int ier = FPB_INPUT_ERROR;
    if (cv && ddx && cv->knots >= 2 * (cv->order + 1))
        fp_spalde_c(cv->t.data(), cv->knots, cv->c.data(), cv->order + 1, x, ddx, &ier);
    return ier;
}

// integral, fitpack_curves.f90 (splint over [from, to]; C handle :263-274)
FP_REAL fitpack_curve_c_integral(fitpack_curve_c *self, FP_REAL from, FP_REAL to)
{
    Curve *cv = ensure(self);
    if (!cv || cv->knots < 2 * (cv->order + 1)) return 0.0;
    std::vector<double> wk(cv->knots);
    return fp_splint_c(cv->t.data(), cv->knots, cv->c.data(), cv->order, from, to, wk.data());
}

// fourier_coefficients, fitpack_curves.f90:650-668 (C handle src/fitpack_curves_c.f90:276-291): fourco on the curve's
// knots and coefficients with wrk_fou(:,1), wrk_fou(:,2) as workspace; A = sine integrals, B = cosine integrals
void fitpack_curve_c_fourier(fitpack_curve_c *self, FP_SIZE nparm, const FP_REAL *alpha, FP_REAL *A, FP_REAL *B,
                             FP_FLAG *ierr)
{
    Curve *cv = ensure(self);
    FP_FLAG ier = FPB_INPUT_ERROR;
    if (cv) {
        if (cv->wrk_fou.size() < static_cast<size_t>(cv->nest) * 2) cv->wrk_fou.assign(static_cast<size_t>(cv->nest) * 2, 0.0);
        if (cv->knots >= 1 && cv->knots <= cv->nest && (nparm <= 0 || (alpha && A && B)))
            fp_fourco_c(cv->t.data(), cv->knots, cv->c.data(), alpha, nparm, A, B, cv->wrk_fou.data(),
                        cv->wrk_fou.data() + cv->nest, &ier);
    }
    if (ierr) *ierr = ier;
}

FP_REAL fitpack_curve_c_smoothing(fitpack_curve_c *self)
{
    Curve *cv = ensure(self);
    return cv ? cv->smoothing : 0.0;
}

FP_REAL fitpack_curve_c_mse(fitpack_curve_c *self)
{
    Curve *cv = ensure(self);
    return cv ? cv->fp : 0.0;
}

FP_SIZE fitpack_curve_c_degree(fitpack_curve_c *self)
{
    Curve *cv = ensure(self);
    return cv ? cv->order : 0;
}

void fitpack_curve_c_set_bc(fitpack_curve_c *self, FP_FLAG bc)
{
    Curve *cv = ensure(self);
    if (!cv) return;
    // invalid values are silently ignored, fitpack_curves_c.f90:374-378
    if (bc != FPB_OUTSIDE_EXTRAPOLATE && bc != FPB_OUTSIDE_ZERO && bc != FPB_OUTSIDE_NOT_ALLOWED &&
        bc != FPB_OUTSIDE_NEAREST_BND)
        return;
    cv->bc = bc;
}

FP_FLAG fitpack_curve_c_get_bc(fitpack_curve_c *self)
{
    Curve *cv = ensure(self);
    return cv ? cv->bc : FPB_OUTSIDE_NEAREST_BND;
}

FP_SIZE fitpack_curve_c_knots(fitpack_curve_c *self)
{
    Curve *cv = ensure(self);
    return cv ? cv->knots : 0;
}

FP_FLAG fitpack_curve_c_iopt(fitpack_curve_c *self)
{
    Curve *cv = ensure(self);
    return cv ? cv->iopt : 0;
}

FP_SIZE fitpack_curve_c_get_t(fitpack_curve_c *self, FP_SIZE nmax, FP_REAL *t)
{
    Curve *cv = ensure(self);
    if (!cv || !t) return 0;
    const int n = std::min(nmax, cv->knots);
    if (n > 0) std::memcpy(t, cv->t.data(), sizeof(double) * n);
    return n > 0 ? n : 0;
}

FP_SIZE fitpack_curve_c_get_c(fitpack_curve_c *self, FP_SIZE nmax, FP_REAL *c)
{
    Curve *cv = ensure(self);
    if (!cv || !c) return 0;
    const int n = std::min(nmax, cv->knots - cv->order - 1);
    if (n > 0) std::memcpy(c, cv->c.data(), sizeof(double) * n);
    return n > 0 ? n : 0;
}

FP_SIZE fitpack_curve_c_set_knots(fitpack_curve_c *self, FP_SIZE n, const FP_REAL *t)
{
    Curve *cv = ensure(self);
    if (!cv || !t || n < 0 || n > cv->nest) return 0;
    std::memcpy(cv->t.data(), t, sizeof(double) * n);
    cv->knots = n;
    return n;
}

}  // extern "C"
