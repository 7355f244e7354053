// fpb_calculus.cu -- spline calculus on a fitted curve: splder, spalde, splint,
// the routines behind fitpack_curve_c_derivative / _all_derivatives / _integral.
//
// Reference: splder core:17699-17802, spalde core:16865-16895, fpader core:1546-1603,
// splint core:17919-17939, fpintb core:8058-8166; C binding src/fitpack_core_c.f90:157-175,
// 207-213 ("core" = src/fitpack_core.F90).
//
// The derivative of order nu of a degree-k spline is the degree k-nu spline on the knots
// t(nu+1 : n-nu) whose coefficients come from de Boor's difference recurrence (core:17736-17753);
// a one-thread kernel runs that recurrence (O(n*nu), serial) and the EVALUATION then goes through
// the same splev kernel as fp_splev_c (identical arithmetic: fpbspl of degree k-nu + dot product),
// so many-point derivative evaluation is as parallel as splev.  spalde / splint are O(k^2) / O(n)
// scalar recurrences per call: one-thread kernels, same IEEE operations in the same order (no FMA).
#include "fpb_common.cuh"
#include "fpb_engine.h"

#include <cstring>
#include <vector>

#include "../../include/fitpack_b200.h"

namespace fpb {

namespace {

// fp_knot_interval, core:2432-2473: smallest l >= l_start with x < t(l+1), capped at l_max
__device__ __forceinline__ int knot_interval(const double *t, double x, int l_start, int l_max)
{
    int l = l_start;
    while (l < l_max && x >= t[l]) ++l;
    return l;
}

// derivative coefficients, core:17736-17753
__global__ void deriv_coef_kernel(const double *__restrict__ t, int n, const double *__restrict__ c,
                                  int k, int nu, double *__restrict__ wrk)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const int k1 = k + 1, nk1 = n - k1;
    auto T = [&](int i) { return t[i - 1]; };
    for (int i = 0; i < nk1; ++i) wrk[i] = c[i];
    int l = 1, nk2 = nk1;
    for (int j = 1; j <= nu; ++j) {
        const double ak = (double)(k1 - j);
        nk2 = nk2 - 1;
        int l1 = l;
        for (int i = 1; i <= nk2; ++i) {
            l1 = l1 + 1;
            const int l2 = l1 + k1 - j;
            const double fac = sub(T(l2), T(l1));
            if (fac > 0.0) wrk[i - 1] = mul(ak, sub(wrk[i], wrk[i - 1])) / fac;
        }
        l = l + 1;
    }
}

// nu == k: the derivative is piecewise constant, y = wrk(l-k) (core:17793-17796)
__global__ void flat_deriv_kernel(const double *__restrict__ t, int n, const double *__restrict__ wrk,
                                  int k, const double *__restrict__ x, double *__restrict__ y, int m, int e)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const int k1 = k + 1, nk1 = n - k1;
    const double tb = t[k1 - 1], te = t[nk1];
    const double arg = x[i];
    if ((arg < tb || arg > te) && e == 1) {
        y[i] = 0.0;
        return;
    }
    int l = k1;  // rightmost l in [k1, nk1] with t(l) <= arg (floor k1)
    while (l < nk1 && arg >= t[l]) ++l;
    y[i] = wrk[l - k - 1];
}

// spalde + fpader, core:16865-16895, 1546-1603
__global__ void spalde_kernel(const double *__restrict__ t, int n, const double *__restrict__ c, int k1,
                              double x, double *__restrict__ d, int *__restrict__ ier)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    auto T = [&](int i) { return t[i - 1]; };
    const int nk1 = n - k1;
    *ier = IER_INPUT;
    if (x < T(k1) || x > T(nk1 + 1)) return;
    const int l = knot_interval(t, x, k1, nk1);
    if (T(l) >= T(l + 1)) return;
    *ier = IER_OK;
    double h[8], dd[8];
    const int lk = l - k1;
    for (int i = 1; i <= k1; ++i) h[i] = c[i + lk - 1];
    int kj = k1;
    double fac = 1.0;
    for (int j = 1; j <= k1; ++j) {
        int ki = kj;
        if (j > 1) {
            int i = k1;
            for (int jj = j; jj <= k1; ++jj) {
                const int li = i + lk, lj = li + kj;
                h[i] = sub(h[i], h[i - 1]) / sub(T(lj), T(li));
                --i;
            }
        }
        for (int i = j; i <= k1; ++i) dd[i] = h[i];
        if (j < k1) {
            for (int jj = j + 1; jj <= k1; ++jj) {
                ki = ki - 1;
                int i = k1;
                for (int j2 = jj; j2 <= k1; ++j2) {
                    const int li = i + lk, lj = li + ki;
                    dd[i] = add(mul(sub(x, T(li)), dd[i]), mul(sub(T(lj), x), dd[i - 1])) / sub(T(lj), T(li));
                    --i;
                }
            }
        }
        d[j - 1] = mul(dd[k1], fac);
        const double ak = (double)(k1 - j);
        fac = mul(fac, ak);
        kj = kj - 1;
    }
}

// splint + fpintb, core:17919-17939, 8058-8166: bint(nk1) and the integral out[0]
__global__ void splint_kernel(const double *__restrict__ t, int n, const double *__restrict__ c, int k,
                              double x, double y, double *__restrict__ bint, double *__restrict__ out)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    auto T = [&](int i) { return t[i - 1]; };
    const int k1 = k + 1, nk1 = n - k1;
    const double ak = (double)k1;
    for (int i = 0; i < nk1; ++i) bint[i] = 0.0;
    bool done = fp_equal(x, y);
    double a = fmin(x, y), b = fmax(x, y);
    const bool lmin = x > y;
    if (!done) {
        a = fmax(T(k1), a);
        b = fmin(T(nk1 + 1), b);
        if (a > b) done = true;
    }
    if (!done) {
        double aint[8], h[8], h1[8];
        int l = k1, ia = 0;
        double arg = a;
        for (int it = 1; it <= 2; ++it) {
            l = knot_interval(t, arg, l, nk1);
            aint[1] = sub(arg, T(l)) / sub(T(l + 1), T(l));
            for (int i = 2; i <= k1; ++i) aint[i] = 0.0;
            h1[1] = 1.0;
            for (int j = 1; j <= k; ++j) {
                h[1] = 0.0;
                for (int i = 1; i <= j; ++i) {
                    const int li = l + i, lj = li - j;
                    const double f = h1[i] / sub(T(li), T(lj));
                    h[i] = add(h[i], mul(f, sub(T(li), arg)));
                    h[i + 1] = mul(f, sub(arg, T(lj)));
                }
                const int j1 = j + 1;
                for (int i = 1; i <= j1; ++i) {
                    const int li = l + i, lj = li - j1;
                    aint[i] = add(aint[i], mul(h[i], sub(arg, T(lj))) / sub(T(li), T(lj)));
                    h1[i] = h[i];
                }
            }
            if (it < 2) {
                int lk = l - k;
                ia = lk;
                for (int i = 1; i <= k1; ++i) {
                    bint[lk - 1] = -aint[i];
                    ++lk;
                }
                arg = b;
            }
        }
        int lk = l - k;
        const int ib = lk - 1;
        for (int i = 1; i <= k1; ++i) {
            bint[lk - 1] = add(bint[lk - 1], aint[i]);
            ++lk;
        }
        for (int i = ia; i <= ib; ++i) bint[i - 1] = add(bint[i - 1], 1.0);
        const double f = 1.0 / ak;
        for (int i = 1; i <= nk1; ++i) bint[i - 1] = mul(mul(bint[i - 1], sub(T(i + k1), T(i))), f);
        if (lmin)
            for (int i = 0; i < nk1; ++i) bint[i] = -bint[i];
    }
    double s = 0.0;
    for (int i = 0; i < nk1; ++i) s = add(s, mul(c[i], bint[i]));
    out[0] = s;
}

struct DevSel {
    int prev = -1;
    bool ok = true;
    explicit DevSel(int dev)
    {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != dev) ok = cuda_ok(cudaSetDevice(dev), "cudaSetDevice");
    }
    ~DevSel()
    {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

// ---- fourco: Fourier integrals of a cubic spline (core:1474-1520, fpbfou 1947-2123, fpcsin 4636-4686) ----
// One thread per alfa(i): the integrals of the n-4 cubic B-splines against sin / cos(alfa x) -- divided differences at
// the boundaries and for |alfa * support| > 1, a series otherwise -- summed with the coefficients in index order.
// sin / cos are CUDA's double-precision functions (<= 2 ulp; the reference's are libm's): results agree with the
// reference to rounding level, not bit for bit -- the one routine here with a tolerance instead of identity.
__device__ void fpcsin_dev(double a, double b, double par, double sia, double coa, double sib, double cob, double &ress,
                           double &resc)
{
    const double eps = 1.0e-10;
    const double ab = sub(b, a), ab4 = mul(mul(ab, ab), mul(ab, ab)), alfa = mul(ab, par);
    if (fabs(alfa) <= 1.0) {
        double fac = 0.25, f1 = fac, f2 = 0.0;
        int i = 4;
        for (int j = 1; j <= 5; ++j) {
            ++i;
            fac = mul(fac, alfa) / (double)i;
            f2 = add(f2, fac);
            if (fabs(fac) <= eps) break;
            ++i;
            fac = mul(-fac, alfa) / (double)i;
            f1 = add(f1, fac);
            if (fabs(fac) <= eps) break;
        }
        ress = mul(ab4, add(mul(coa, f2), mul(sia, f1)));
        resc = mul(ab4, sub(mul(coa, f1), mul(sia, f2)));
    } else {
        const double beta = 1.0 / alfa, b2 = mul(beta, beta), b4 = mul(6.0, mul(b2, b2));
        const double f1 = mul(mul(3.0, b2), sub(1.0, mul(2.0, b2))), f2 = mul(beta, sub(1.0, mul(6.0, b2)));
        ress = mul(ab4, add(add(mul(coa, f2), mul(sia, f1)), mul(sib, b4)));
        resc = mul(ab4, add(sub(mul(coa, f1), mul(sia, f2)), mul(cob, b4)));
    }
}

__global__ void fourco_kernel(const double *__restrict__ t, int n, const double *__restrict__ c,
                              const double *__restrict__ alfa, int m, double *__restrict__ ress,
                              double *__restrict__ resc, double *__restrict__ wrk1, double *__restrict__ wrk2)
{
    const int ia = blockIdx.x * blockDim.x + threadIdx.x;
    if (ia >= m) return;
    const bool keep = (ia == m - 1);   // the reference leaves the integrals of the LAST alfa in wrk1 / wrk2
    auto T = [&](int i) { return t[i - 1]; };
    const double par = alfa[ia];
    const double eps = 1.0e-08, con1 = 0.5e-01, con2 = 0.12e+03;
    double co[6], si[6], hs[6], hc[6], rs[4], rc[4];   // 1-based
    const int nm3 = n - 3, nm7 = n - 7, n4 = n - 4;
    const double term = fp_is_zero(par) ? 0.0 : 6.0 / par;
    double sum_s = 0.0, sum_c = 0.0;   // dot_product(c(1:n4), .) in index order
    auto emit = [&](int j, double vs, double vc) {
        sum_s = add(sum_s, mul(c[j - 1], vs));
        sum_c = add(sum_c, mul(c[j - 1], vc));
        if (keep) {
            wrk1[j - 1] = vs;
            wrk2[j - 1] = vc;
        }
    };
    double beta = mul(par, T(4));
    co[1] = cos(beta);
    si[1] = sin(beta);
    for (int j = 1; j <= 3; ++j) {   // left boundary
        const int jp1 = j + 1, jp4 = j + 4;
        beta = mul(par, T(jp4));
        co[jp1] = cos(beta);
        si[jp1] = sin(beta);
        fpcsin_dev(T(4), T(jp4), par, si[1], co[1], si[jp1], co[jp1], rs[j], rc[j]);
        int i = 5 - j;
        hs[i] = 0.0;
        hc[i] = 0.0;
        for (int jj = 1; jj <= j; ++jj) {
            hs[i + jj] = rs[jj];
            hc[i + jj] = rc[jj];
        }
        for (int jj = 1; jj <= 3; ++jj) {
            if (i < jj) i = jj;
            int k = 5, li = jp4;
            for (int ll = i; ll <= 4; ++ll) {
                const double fac = sub(T(li), T(li - jj));
                hs[k] = sub(hs[k], hs[k - 1]) / fac;
                hc[k] = sub(hc[k], hc[k - 1]) / fac;
                --k;
                --li;
            }
        }
        emit(j, sub(hs[5], hs[4]), sub(hc[5], hc[4]));
    }
    for (int j = 4; j <= nm7; ++j) {   // interior
        const int jp4 = j + 4;
        beta = mul(par, T(jp4));
        co[5] = cos(beta);
        si[5] = sin(beta);
        const double delta = sub(T(jp4), T(j));
        beta = mul(delta, par);
        double s2, c2;
        if (fabs(beta) > 1.0) {
            for (int i = 1; i <= 5; ++i) {
                hs[i] = si[i];
                hc[i] = co[i];
            }
            for (int jj = 1; jj <= 3; ++jj) {
                int k = 5, li = jp4;
                for (int ll = jj; ll <= 4; ++ll) {
                    const double fac = mul(par, sub(T(li), T(li - jj)));
                    hs[k] = sub(hs[k], hs[k - 1]) / fac;
                    hc[k] = sub(hc[k], hc[k - 1]) / fac;
                    --k;
                    --li;
                }
            }
            s2 = mul(sub(hs[5], hs[4]), term);
            c2 = mul(sub(hc[5], hc[4]), term);
        } else {
            for (int i = 1; i <= 4; ++i) {
                hs[i] = mul(par, sub(T(j + i), T(j)));
                hc[i] = hs[i];
            }
            double f3 = mul(con1, add(add(add(hs[1], hs[2]), hs[3]), hs[4]));
            double c1 = 0.25, s1 = f3;
            if (fabs(f3) > eps) {
                double sign = 1.0, fac = con2;
                int k = 5, is = 0;
                for (int ic = 1; ic <= 20; ++ic) {
                    ++k;
                    fac = mul(fac, (double)k);
                    double f1 = 0.0;
                    f3 = 0.0;
                    for (int i = 1; i <= 4; ++i) {
                        f1 = add(f1, hc[i]);
                        const double f2 = mul(f1, hs[i]);
                        hc[i] = f2;
                        f3 = add(f3, f2);
                    }
                    f3 = mul(f3, 6.0) / fac;
                    if (is == 0) {
                        sign = -sign;
                        is = 1;
                        c1 = add(c1, mul(f3, sign));
                    } else {
                        is = 0;
                        s1 = add(s1, mul(f3, sign));
                    }
                    if (fabs(f3) <= eps) break;
                }
            }
            s2 = mul(delta, add(mul(co[1], s1), mul(si[1], c1)));
            c2 = mul(delta, sub(mul(co[1], c1), mul(si[1], s1)));
        }
        emit(j, s2, c2);
        for (int i = 1; i <= 4; ++i) {
            co[i] = co[i + 1];
            si[i] = si[i + 1];
        }
    }
    double es[4], ec[4];   // right boundary: computed for j = n-4, n-5, n-6, summed in index order
    for (int j = 1; j <= 3; ++j) {
        const int nmj = nm3 - j;
        int i = 5 - j;
        fpcsin_dev(T(nm3), T(nmj), par, si[4], co[4], si[i - 1], co[i - 1], rs[j], rc[j]);
        hc[i] = 0.0;
        hs[i] = 0.0;
        for (int jj = 1; jj <= j; ++jj) {
            hc[i + jj] = rc[jj];
            hs[i + jj] = rs[jj];
        }
        for (int jj = 1; jj <= 3; ++jj) {
            if (i < jj) i = jj;
            int k = 5, li = nmj;
            for (int ll = i; ll <= 4; ++ll) {
                const double fac = sub(T(li + jj), T(li));
                hs[k] = sub(hs[k - 1], hs[k]) / fac;
                hc[k] = sub(hc[k - 1], hc[k]) / fac;
                --k;
                ++li;
            }
        }
        es[j] = sub(hs[4], hs[5]);
        ec[j] = sub(hc[4], hc[5]);
    }
    for (int j = 3; j >= 1; --j) emit(nm3 - j, es[j], ec[j]);
    (void)n4;
    ress[ia] = sum_s;
    resc[ia] = sum_c;
}

// insert_inplace -> insert -> fpinst (core:15138-15159, 15064-15120, 7961-8026), in place on the device copies of t and c:
// the knot x goes in after t(l), the k coefficients it touches become convex combinations, the rest shift by one; a
// periodic spline (iopt /= 0) gets its wrap-around knots and coefficients repaired.  One thread: O(n) moves.
__global__ void insert_kernel(int iopt, double *__restrict__ t, int *__restrict__ n_io, double *__restrict__ c, int k,
                              double x, int nest, int *__restrict__ ier)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    auto T = [&](int i) -> double & { return t[i - 1]; };
    auto Cc = [&](int i) -> double & { return c[i - 1]; };
    const int n = *n_io, k1 = k + 1, nk = n - k;
    *ier = IER_INPUT;
    if (nest <= n) return;
    if (x < T(k1) || x > T(nk)) return;
    const int l = knot_interval(t, x, k1, nk - 1);
    if (T(l) >= T(l + 1)) return;
    if (iopt != 0) {
        const int kk = 2 * k;
        if (l <= kk && l >= (n - kk)) return;
    }
    *ier = IER_OK;
    const int nk1 = n - k1, ll = l + 1;
    for (int i = n; i >= ll; --i) T(i + 1) = T(i);
    T(ll) = x;
    for (int i = nk1; i >= l; --i) Cc(i + 1) = Cc(i);
    int i = l;
    for (int j = 1; j <= k; ++j) {
        const int m = i + k1;
        const double fac = sub(x, T(i)) / sub(T(m), T(i));
        Cc(i) = add(mul(fac, Cc(i)), mul(sub(1.0, fac), Cc(i - 1)));
        --i;
    }
    const int nn = n + 1;
    if (iopt != 0) {
        const int nkk = nn - k, nl = nkk - k1;
        const double per = sub(T(nkk), T(k1));
        int ii = k1, jj = nkk;
        if (ll > nl) {
            for (int m = 1; m <= k; ++m) {
                Cc(m) = Cc(m + nl);
                --ii;
                --jj;
                T(ii) = sub(T(jj), per);
            }
        } else if (ll <= (k1 + k)) {
            for (int m = 1; m <= k; ++m) {
                Cc(m + nl) = Cc(m);
                ++ii;
                ++jj;
                T(jj) = add(T(ii), per);
            }
        }
    }
    *n_io = nn;
}

// ---- sproot: the zeros of a cubic spline (core:17962-18109; fpcuro core:5108-5198) ----
// One thread walks the knot intervals as the reference does (the polynomial of an interval reuses its neighbour's end
// values), solves each cubic in closed form + one Newton correction, sorts and removes duplicates.  pow / atan2 / cos
// are CUDA's: zeros agree with the reference to rounding level (tolerance in the tests), not bit for bit.
__device__ void fpcuro_dev(double a, double b, double c, double d, double *x, int &n)
{
    const double ovfl = 1.0e4, tent = 0.1;
    const double pi3 = atan(1.0) / 0.75, e3 = tent / 0.3;
    double a1 = fabs(a), b1 = fabs(b), c1 = fabs(c), d1 = fabs(d);
    if (fmax(fmax(b1, c1), d1) < mul(a1, ovfl)) {
        b1 = mul(b / a, e3);
        c1 = c / a;
        d1 = d / a;
        const double q = sub(mul(c1, e3), mul(b1, b1));
        const double r = add(mul(mul(b1, b1), b1), mul(sub(d1, mul(b1, c1)), 0.5));
        const double disc = add(mul(mul(q, q), q), mul(r, r));
        if (disc > 0.0) {
            const double u = sqrt(disc), u1 = add(-r, u), u2 = sub(-r, u);
            n = 1;
            x[0] = sub(add(copysign(pow(fabs(u1), e3), u1), copysign(pow(fabs(u2), e3), u2)), b1);
        } else {
            const double u = copysign(sqrt(fabs(q)), r);
            const double p3 = mul(atan2(sqrt(-disc), fabs(r)), e3);
            const double u2 = add(u, u);
            n = 3;
            x[0] = sub(mul(-u2, cos(p3)), b1);
            x[1] = sub(mul(u2, cos(sub(pi3, p3))), b1);
            x[2] = sub(mul(u2, cos(add(pi3, p3))), b1);
        }
    } else if (fmax(c1, d1) < mul(b1, ovfl)) {
        const double disc = sub(mul(c, c), mul(mul(4.0, b), d));
        if (disc < 0.0) {
            n = 0;
            return;
        }
        n = 2;
        const double u = sqrt(disc), bb = add(b, b);
        x[0] = add(-c, u) / bb;
        x[1] = sub(-c, u) / bb;
    } else if (d1 < mul(c1, ovfl)) {
        n = 1;
        x[0] = -d / c;
    } else {
        n = 0;
        return;
    }
    for (int i = 0; i < n; ++i) {
        const double y = x[i];
        const double f = add(mul(add(mul(add(mul(a, y), b), y), c), y), d);
        const double df = add(mul(add(mul(mul(3.0, a), y), mul(2.0, b)), y), c);
        const double step = (fabs(f) < mul(fabs(df), tent)) ? f / df : 0.0;
        x[i] = sub(y, step);
    }
}

__global__ void sproot_kernel(const double *__restrict__ t, int n, const double *__restrict__ c,
                              double *__restrict__ zeros, int mest, int *__restrict__ m_out, int *__restrict__ ier)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    auto T = [&](int i) { return t[i - 1]; };
    auto Cc = [&](int i) { return c[i - 1]; };
    const int n4 = n - 4;
    int m = 0;
    double h1 = sub(T(4), T(3)), h2 = sub(T(5), T(4)), t1 = sub(T(4), T(2)), t2 = sub(T(5), T(3)), t3 = sub(T(6), T(4)),
           t4 = sub(T(5), T(2)), t5 = sub(T(6), T(3));
    double c1 = Cc(1), c2 = Cc(2), c3 = Cc(3);
    double c4 = sub(c2, c1) / t4, c5 = sub(c3, c2) / t5;
    double d4 = add(mul(h2, c1), mul(t1, c2)) / t4, d5 = add(mul(t3, c2), mul(h1, c3)) / t5;
    double a0 = add(mul(h2, d4), mul(h1, d5)) / t2;
    double ah = mul(3.0, add(mul(h2, c4), mul(h1, c5))) / t2;
    bool z1 = !(ah < 0.0), nz1 = !z1;
    *ier = IER_OK;
    for (int l = 4; l <= n4; ++l) {
        h1 = h2;
        h2 = sub(T(l + 2), T(l + 1));
        t1 = t2;
        t2 = t3;
        t3 = sub(T(l + 3), T(l + 1));
        t4 = t5;
        t5 = sub(T(l + 3), T(l));
        c1 = c2;
        c2 = c3;
        c3 = Cc(l);
        c4 = c5;
        c5 = sub(c3, c2) / t5;
        d4 = add(mul(h2, c1), mul(t1, c2)) / t4;
        d5 = add(mul(h1, c3), mul(t3, c2)) / t5;
        const double b0 = add(mul(h2, d4), mul(h1, d5)) / t2;
        const double bh = mul(3.0, add(mul(h2, c4), mul(h1, c5))) / t2;
        const double a1 = mul(ah, h1), b1 = mul(bh, h1);
        const double a2 = sub(sub(mul(3.0, sub(b0, a0)), b1), mul(2.0, a1));
        const double a3 = add(add(mul(2.0, sub(a0, b0)), b1), a1);
        const bool z0 = !(a0 < 0.0), nz0 = !z0, z2 = !(a2 < 0.0), nz2 = !z2, z3 = !(b1 < 0.0), nz3 = !z3;
        const bool z4 = !(add(mul(3.0, a3), a2) < 0.0), nz4 = !z4;
        if (mul(a0, b0) <= 0.0 || ((z0 && ((nz1 && (z3 || (z2 && nz4))) || (nz2 && z3 && z4))) ||
                                   (nz0 && ((z1 && (nz3 || (nz2 && z4))) || (z2 && nz3 && nz4))))) {
            double y[3];
            int j = 0;
            fpcuro_dev(a3, a2, a1, a0, y, j);
            for (int i = 0; i < j; ++i) {
                if (y[i] < 0.0 || y[i] > 1.0) continue;
                if (m >= mest) {
                    *ier = IER_STORAGE;
                    *m_out = m;
                    return;
                }
                zeros[m++] = add(T(l), mul(h1, y[i]));
            }
        }
        a0 = b0;
        ah = bh;
        z1 = z3;
        nz1 = nz3;
    }
    if (m >= 2) {
        for (int i = 1; i < m; ++i) {   // ascending (the result of fitpack_quicksort is order-unique)
            const double v = zeros[i];
            int j = i - 1;
            for (; j >= 0 && zeros[j] > v; --j) zeros[j + 1] = zeros[j];
            zeros[j + 1] = v;
        }
        const int j = m;
        m = 1;
        for (int i = 2; i <= j; ++i) {
            if (fp_equal(zeros[i - 1], zeros[m - 1])) continue;
            ++m;
            zeros[m - 1] = zeros[i - 1];
        }
    }
    *m_out = m;
}

// upload knots and coefficients of one spline into the engine's calculus buffers
bool upload_spline(fpb_engine *eng, const double *t, const double *c, int n, double **dt, double **dc,
                   double **dw)
{
    DevBuf *B = eng->calc;
    if (!B[0].reserve((size_t)n * 8) || !B[1].reserve((size_t)n * 8) || !B[2].reserve((size_t)n * 8 + 64))
        return false;
    *dt = B[0].as<double>();
    *dc = B[1].as<double>();
    *dw = B[2].as<double>();
    return cuda_ok(cudaMemcpyAsync(*dt, t, (size_t)n * 8, cudaMemcpyHostToDevice, eng->stream), "H2D t") &&
           cuda_ok(cudaMemcpyAsync(*dc, c, (size_t)n * 8, cudaMemcpyHostToDevice, eng->stream), "H2D c");
}

}  // namespace

}  // namespace fpb

using namespace fpb;

extern "C" {

// reference: src/fitpack_core_c.f90:157-165 -> splder core:17699-17802
void fp_splder_c(const FP_REAL *t, FP_SIZE n, const FP_REAL *c, FP_SIZE k, FP_SIZE nu, const FP_REAL *x,
                 FP_REAL *y, FP_SIZE m, FP_FLAG e, FP_REAL *wrk, FP_FLAG *ier)
{
    *ier = FPB_INPUT_ERROR;
    if (nu < 0 || nu > k) return;
    if (m < 1) return;
    if (k < 1 || k > 5 || n < 2 * (k + 1)) return;  // not a valid spline: nothing safe to evaluate
    HostApiLock api_lock(-1);  // per-device serialisation (fpb_engine.h)
    fpb_engine *eng = default_engine(-1);
    if (!eng) {
        *ier = FPB_ERR_CUDA;
        return;
    }
    DevSel g(eng->device);
    const int nk1 = n - k - 1;
    double *dt, *dc, *dw;
    if (!upload_spline(eng, t, c, n, &dt, &dc, &dw)) {
        *ier = FPB_ERR_ALLOC;
        return;
    }
    deriv_coef_kernel<<<1, 32, 0, eng->stream>>>(dt, n, dc, k, nu, dw);
    eng->launches++;
    if (!cuda_ok(cudaMemcpyAsync(wrk, dw, (size_t)nk1 * 8, cudaMemcpyDeviceToHost, eng->stream), "D2H wrk") ||
        !cuda_ok(cudaStreamSynchronize(eng->stream), "sync")) {
        *ier = FPB_ERR_CUDA;
        return;
    }
    const int kk = k - nu;
    if (kk >= 1) {
        // degree k-nu spline on t(nu+1 : n-nu): evaluated by the splev kernel.  splder has no
        // "nearest boundary" policy: anything but 1 / 2 extrapolates (core:17764-17775)
        const FP_FLAG es = (e == 1) ? 1 : (e == 2) ? 2 : 0;
        FP_FLAG ie = FPB_INPUT_ERROR;
        fp_splev_c(t + nu, n - 2 * nu, wrk, kk, x, y, m, es, &ie);
        *ier = (ie == FPB_INVALID_RANGE) ? FPB_INSUFFICIENT_STORAGE : ie;  // splder reports code 1
        return;
    }
    DevBuf *B = eng->calc;
    if (!B[3].reserve((size_t)m * 8) || !B[4].reserve((size_t)m * 8)) {
        *ier = FPB_ERR_ALLOC;
        return;
    }
    bool ok = cuda_ok(cudaMemcpyAsync(B[3].ptr, x, (size_t)m * 8, cudaMemcpyHostToDevice, eng->stream), "H2D x");
    if (ok) {
        flat_deriv_kernel<<<(m + 127) / 128, 128, 0, eng->stream>>>(dt, n, dw, k, B[3].as<double>(),
                                                                     B[4].as<double>(), m, e);
        eng->launches++;
    }
    long long upto = m;
    int result = FPB_OK;
    if (e == 2) {  // y is filled up to the first point outside the support (core:17771-17773)
        const double tb = t[k], te = t[n - k - 1];
        for (long long i = 0; i < m; ++i)
            if (x[i] < tb || x[i] > te) {
                upto = i;
                result = FPB_INSUFFICIENT_STORAGE;
                break;
            }
    }
    if (ok && upto > 0)
        ok = cuda_ok(cudaMemcpyAsync(y, B[4].ptr, (size_t)upto * 8, cudaMemcpyDeviceToHost, eng->stream), "D2H y");
    ok = ok && cuda_ok(cudaStreamSynchronize(eng->stream), "sync");
    *ier = ok ? result : FPB_ERR_CUDA;
}

// reference: src/fitpack_core_c.f90:168-175 -> spalde core:16865-16895
void fp_spalde_c(const FP_REAL *t, FP_SIZE n, const FP_REAL *c, FP_SIZE k1, FP_REAL x, FP_REAL *d,
                 FP_FLAG *ier)
{
    *ier = FPB_INPUT_ERROR;
    if (k1 < 2 || k1 > 6 || n < 2 * k1) return;
    HostApiLock api_lock(-1);  // per-device serialisation (fpb_engine.h)
    fpb_engine *eng = default_engine(-1);
    if (!eng) {
        *ier = FPB_ERR_CUDA;
        return;
    }
    DevSel g(eng->device);
    double *dt, *dc, *dw;
    if (!upload_spline(eng, t, c, n, &dt, &dc, &dw)) {
        *ier = FPB_ERR_ALLOC;
        return;
    }
    int *dier = reinterpret_cast<int *>(dw + 6);
    spalde_kernel<<<1, 32, 0, eng->stream>>>(dt, n, dc, k1, x, dw, dier);
    eng->launches++;
    double hd[8];
    if (!cuda_ok(cudaMemcpyAsync(hd, dw, 8 * 8, cudaMemcpyDeviceToHost, eng->stream), "D2H d") ||
        !cuda_ok(cudaStreamSynchronize(eng->stream), "sync")) {
        *ier = FPB_ERR_CUDA;
        return;
    }
    int hier;
    std::memcpy(&hier, &hd[6], sizeof(int));
    *ier = hier;
    if (hier == FPB_OK)
        for (int j = 0; j < k1; ++j) d[j] = hd[j];
}

// reference: src/fitpack_core_c.f90:207-213 -> splint core:17919-17939
FP_REAL fp_splint_c(const FP_REAL *t, FP_SIZE n, const FP_REAL *c, FP_SIZE k, FP_REAL a, FP_REAL b,
                    FP_REAL *wrk)
{
    const double nan = __builtin_nan("");
    if (k < 1 || k > 5 || n < 2 * (k + 1)) return nan;
    HostApiLock api_lock(-1);  // per-device serialisation (fpb_engine.h)
    fpb_engine *eng = default_engine(-1);
    if (!eng) return nan;
    DevSel g(eng->device);
    const int nk1 = n - k - 1;
    double *dt, *dc, *dw;
    if (!upload_spline(eng, t, c, n, &dt, &dc, &dw)) return nan;
    splint_kernel<<<1, 32, 0, eng->stream>>>(dt, n, dc, k, a, b, dw, dw + n);
    eng->launches++;
    std::vector<double> h((size_t)n + 1);
    if (!cuda_ok(cudaMemcpyAsync(h.data(), dw, ((size_t)n + 1) * 8, cudaMemcpyDeviceToHost, eng->stream), "D2H") ||
        !cuda_ok(cudaStreamSynchronize(eng->stream), "sync"))
        return nan;
    if (wrk)
        for (int i = 0; i < nk1; ++i) wrk[i] = h[i];
    return h[n];
}

// reference: src/fitpack_core_c.f90:216-223 -> fourco core:1474-1520
void fp_fourco_c(const FP_REAL *t, FP_SIZE n, const FP_REAL *c, const FP_REAL *alfa, FP_SIZE m, FP_REAL *ress,
                 FP_REAL *resc, FP_REAL *wrk1, FP_REAL *wrk2, FP_FLAG *ier)
{
    // data checks, core:1489-1505
    if (n < 10) {
        *ier = FPB_INSUFFICIENT_KNOTS;
        return;
    }
    *ier = FPB_INPUT_ERROR;
    for (int j = 0; j < 3; ++j)
        if (t[j] > t[j + 1]) return;
    for (int j = n - 3; j < n; ++j)
        if (t[j] < t[j - 1]) return;
    for (int j = 3; j < n - 4; ++j)
        if (t[j] >= t[j + 1]) return;
    *ier = FPB_OK;
    if (m < 1) return;
    HostApiLock api_lock(-1);  // per-device serialisation (fpb_engine.h)
    fpb_engine *eng = default_engine(-1);
    if (!eng) {
        *ier = FPB_ERR_CUDA;
        return;
    }
    DevSel g(eng->device);
    double *dt, *dc, *dw;
    DevBuf *B = eng->calc;
    if (!upload_spline(eng, t, c, n, &dt, &dc, &dw) || !B[2].reserve((size_t)2 * n * 8 + 64) || !B[3].reserve((size_t)m * 8) ||
        !B[4].reserve((size_t)2 * m * 8)) {
        *ier = FPB_ERR_ALLOC;
        return;
    }
    dw = B[2].as<double>();
    double *da = B[3].as<double>(), *dr = B[4].as<double>();
    bool ok = cuda_ok(cudaMemcpyAsync(da, alfa, (size_t)m * 8, cudaMemcpyHostToDevice, eng->stream), "H2D alfa");
    if (ok) {
        fourco_kernel<<<(m + 63) / 64, 64, 0, eng->stream>>>(dt, n, dc, da, m, dr, dr + m, dw, dw + n);
        eng->launches++;
    }
    ok = ok && cuda_ok(cudaMemcpyAsync(ress, dr, (size_t)m * 8, cudaMemcpyDeviceToHost, eng->stream), "D2H ress") &&
         cuda_ok(cudaMemcpyAsync(resc, dr + m, (size_t)m * 8, cudaMemcpyDeviceToHost, eng->stream), "D2H resc");
    if (ok && wrk1) ok = cuda_ok(cudaMemcpyAsync(wrk1, dw, (size_t)(n - 4) * 8, cudaMemcpyDeviceToHost, eng->stream), "D2H wrk1");
    if (ok && wrk2) ok = cuda_ok(cudaMemcpyAsync(wrk2, dw + n, (size_t)(n - 4) * 8, cudaMemcpyDeviceToHost, eng->stream), "D2H wrk2");
    ok = ok && cuda_ok(cudaStreamSynchronize(eng->stream), "sync");
    if (!ok) *ier = FPB_ERR_CUDA;
}

// reference: src/fitpack_core_c.f90:187-194 -> cualde core:1062-1101: all derivatives of a parametric curve at u,
// d((j-1)*idim + i) = d^(j-1) s_i / du^(j-1) -- fpader per coordinate curve (the spalde kernel), interleaved
void fp_cualde_c(FP_SIZE idim, const FP_REAL *t, FP_SIZE n, const FP_REAL *c, FP_SIZE nc, FP_SIZE k1, FP_REAL u,
                 FP_REAL *d, FP_SIZE nd, FP_FLAG *ier)
{
    *ier = FPB_INPUT_ERROR;
    if (idim < 1 || k1 < 2 || k1 > 6 || n < 2 * k1) return;
    if (nd < k1 * idim || nc < n * (idim - 1) + n - k1) return;   // core:1079
    HostApiLock api_lock(-1);
    double h[8];
    for (FP_SIZE i = 0; i < idim; ++i) {
        FP_FLAG ie = FPB_INPUT_ERROR;
        fp_spalde_c(t, n, c + (size_t)i * n, k1, u, h, &ie);   // same range / interval checks as cualde (core:1081-1085)
        if (ie != FPB_OK) {
            *ier = ie;
            return;
        }
        for (FP_SIZE kk = 0; kk < k1; ++kk) d[(size_t)kk * idim + i] = h[kk];
    }
    *ier = FPB_OK;
}

// reference: src/fitpack_core_c.f90:197-204 -> insert_inplace core:15138-15159
void fp_insert_c(FP_SIZE iopt, FP_REAL *t, FP_SIZE *n, FP_REAL *c, FP_SIZE k, FP_REAL x, FP_SIZE nest, FP_FLAG *ier)
{
    *ier = FPB_INPUT_ERROR;
    const int n0 = *n;
    if (k < 1 || k > 5 || n0 < 2 * (k + 1) || nest <= n0) return;   // nest <= n: core:15080
    HostApiLock api_lock(-1);
    fpb_engine *eng = default_engine(-1);
    if (!eng) {
        *ier = FPB_ERR_CUDA;
        return;
    }
    DevSel g(eng->device);
    DevBuf *B = eng->calc;
    const size_t len = (size_t)n0 + 1;
    if (!B[0].reserve(len * 8) || !B[1].reserve(len * 8) || !B[2].reserve(64)) {
        *ier = FPB_ERR_ALLOC;
        return;
    }
    double *dt = B[0].as<double>(), *dc = B[1].as<double>();
    int *dn = B[2].as<int>(), *dier = dn + 1;
    int hni[2] = {n0, FPB_INPUT_ERROR};
    bool ok = cuda_ok(cudaMemcpyAsync(dt, t, (size_t)n0 * 8, cudaMemcpyHostToDevice, eng->stream), "H2D t") &&
              cuda_ok(cudaMemcpyAsync(dc, c, (size_t)n0 * 8, cudaMemcpyHostToDevice, eng->stream), "H2D c") &&
              cuda_ok(cudaMemcpyAsync(dn, hni, 8, cudaMemcpyHostToDevice, eng->stream), "H2D n");
    if (ok) {
        insert_kernel<<<1, 32, 0, eng->stream>>>(iopt, dt, dn, dc, k, x, nest, dier);
        eng->launches++;
    }
    std::vector<double> ht(len), hc(len);
    ok = ok && cuda_ok(cudaMemcpyAsync(hni, dn, 8, cudaMemcpyDeviceToHost, eng->stream), "D2H n") &&
         cuda_ok(cudaMemcpyAsync(ht.data(), dt, len * 8, cudaMemcpyDeviceToHost, eng->stream), "D2H t") &&
         cuda_ok(cudaMemcpyAsync(hc.data(), dc, len * 8, cudaMemcpyDeviceToHost, eng->stream), "D2H c") &&
         cuda_ok(cudaStreamSynchronize(eng->stream), "sync");
    if (!ok) {
        *ier = FPB_ERR_CUDA;
        return;
    }
    *ier = hni[1];
    if (hni[1] != FPB_OK) return;   // t, c, n stay the caller's
    *n = hni[0];
    std::memcpy(t, ht.data(), (size_t)hni[0] * 8);
    std::memcpy(c, hc.data(), (size_t)(hni[0] - k - 1) * 8);
}

// reference: src/fitpack_core_c.f90:226-233 -> sproot core:17962-18109
void fp_sproot_c(const FP_REAL *t, FP_SIZE n, const FP_REAL *c, FP_REAL *zeros, FP_SIZE mest, FP_SIZE *m, FP_FLAG *ier)
{
    *m = 0;
    // data checks, core:17990-18006
    if (n < 8) {
        *ier = FPB_INSUFFICIENT_KNOTS;
        return;
    }
    *ier = FPB_INPUT_ERROR;
    for (int i = 0, j = n - 1; i < 3; ++i, --j) {
        if (t[i] > t[i + 1]) return;
        if (t[j] < t[j - 1]) return;
    }
    for (int i = 3; i < n - 4; ++i)
        if (t[i] >= t[i + 1]) return;
    HostApiLock api_lock(-1);
    fpb_engine *eng = default_engine(-1);
    if (!eng) {
        *ier = FPB_ERR_CUDA;
        return;
    }
    DevSel g(eng->device);
    double *dt, *dc, *dw;
    DevBuf *B = eng->calc;
    const size_t mz = (size_t)std::max(mest, 1);
    if (!upload_spline(eng, t, c, n, &dt, &dc, &dw) || !B[3].reserve(mz * 8) || !B[4].reserve(64)) {
        *ier = FPB_ERR_ALLOC;
        return;
    }
    double *dz = B[3].as<double>();
    int *dm = B[4].as<int>();
    sproot_kernel<<<1, 32, 0, eng->stream>>>(dt, n, dc, dz, mest, dm, dm + 1);
    eng->launches++;
    int hm[2] = {0, FPB_INPUT_ERROR};
    if (!cuda_ok(cudaMemcpyAsync(hm, dm, 8, cudaMemcpyDeviceToHost, eng->stream), "D2H m") ||
        !cuda_ok(cudaStreamSynchronize(eng->stream), "sync")) {
        *ier = FPB_ERR_CUDA;
        return;
    }
    *m = hm[0];
    *ier = hm[1];
    if (hm[0] > 0 && !cuda_ok(cudaMemcpy(zeros, dz, (size_t)std::min(hm[0], mest) * 8, cudaMemcpyDeviceToHost), "D2H zeros"))
        *ier = FPB_ERR_CUDA;
}

}  // extern "C"
