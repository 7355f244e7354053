// fpb_percur.cu -- batched periodic smoothing-spline fit (percur/fpperi), SURVEY 8f rank 4.
//
//   percur  src/fitpack_core.F90:15784-15863  (argument checks, periodic knot extension, fpchep)
//   fpperi  src/fitpack_core.F90:9668-10192   (knot loop, two-block Givens QR, p-iteration)
//   fpperi_reset_interp 10220-10257, fp_rotate_row_2mat 5968-6009, fpbacp 1869-1921,
//   fpchep 2689-2770; shared helpers as in fpb_curfit.cu (fpbspl, fpgivs, fprota, fpdisc, fpknot,
//   root_finding_iterate/fprati).  C binding: fp_percur_c, src/fitpack_core_c.f90:77-87.
//
// First correct path for this row: one thread per curve runs the whole of fpperi (no host
// round-trips), persistent warps pull 32-curve tiles from an atomic queue, inputs are read
// point-major (coalesced across the lanes of a warp), and all per-knot state -- t, c, z, fpint,
// the band block a1, the k wrap-around columns a2, b, g1, g2, the stored basis values q -- lives
// in a per-warp workspace interleaved by lane.  The arithmetic follows the reference statement by
// statement; this translation unit is compiled with -fmad=false (no contraction) and IEEE
// division / square root, so n, knots, ier AND c, fp are bit-identical to the reference order of
// operations (tests compare them exactly against the oracle).
// Not yet cut into passes like the curfit pipeline: lanes whose curve needs fewer knot sets idle
// until the slowest curve of the warp finishes.
#include <algorithm>
#include <cstring>
#include <vector>

#include "../../include/fitpack_b200.h"
#include "fpb_common.cuh"
#include "fpb_engine.h"

namespace fpb {

namespace {

constexpr int kPerThreads = 128;
constexpr int kPerBlocksPerSm = 3;

struct PerLayout {
    int nestp;
    int off_t, off_c, off_z, off_fpint, off_a1, off_a2, off_b, off_g1, off_g2, off_q, per_lane;
};

PerLayout per_layout(int m, int k, int nest)
{
    const int k1 = k + 1, k2 = k + 2;
    PerLayout L;
    int nestp = std::min(nest, m + 2 * k);
    nestp = std::max(nestp, 2 * k1);
    L.nestp = nestp;
    int off = 0;
    L.off_t = off; off += nestp;
    L.off_c = off; off += nestp;
    L.off_z = off; off += nestp;
    L.off_fpint = off; off += nestp;
    L.off_a1 = off; off += nestp * k1;
    L.off_a2 = off; off += nestp * k;
    L.off_b = off; off += nestp * k2;
    L.off_g1 = off; off += nestp * k2;
    L.off_g2 = off; off += nestp * k1;
    L.off_q = off; off += m * k1;
    L.per_lane = off;
    return L;
}

struct PerParams {
    int nbatch, iopt, m, k, nest;
    const double *x, *y, *w;
    long long x_sb, x_si, y_sb, y_si, w_sb, w_si;
    const double *s;
    int s_stride;
    int *n;
    double *t, *c, *fp;
    int *ier;
    double *fpint;
    int *nrdata;
    double *ws;
    int *wsi;
    unsigned *counter;
    PerLayout L;
};

#define W_(off, i) wsd[(size_t)((off) + (i)) * 32]
#define T_(i) W_(P.L.off_t, (i) - 1)
#define C_(i) W_(P.L.off_c, (i) - 1)
#define Z_(i) W_(P.L.off_z, (i) - 1)
#define FPI_(i) W_(P.L.off_fpint, (i) - 1)
#define A1_(i, j) W_(P.L.off_a1, ((i) - 1) * K1 + ((j) - 1))
#define A2_(i, j) W_(P.L.off_a2, ((i) - 1) * K + ((j) - 1))
#define B_(i, j) W_(P.L.off_b, ((i) - 1) * K2 + ((j) - 1))
#define G1_(i, j) W_(P.L.off_g1, ((i) - 1) * K2 + ((j) - 1))
#define G2_(i, j) W_(P.L.off_g2, ((i) - 1) * K1 + ((j) - 1))
#define Q_(it, j) W_(P.L.off_q, ((it) - 1) * K1 + ((j) - 1))
#define NRD_(i) wsi[(size_t)((i) - 1) * 32]

// band-matrix views handed to the two-block rotation / back-substitution: element (i,j) of the
// block at workspace offset `off` with `ld` columns per row
struct Blk {
    double *wsd;
    int off, ld;
    __device__ __forceinline__ double &operator()(int i, int j) const
    {
        return wsd[(size_t)(off + (i - 1) * ld + (j - 1)) * 32];
    }
};

// fp_rotate_row_2mat, core:5968-6009
__device__ void rotate_row_2mat(double *h1, int band1, double *h2, int band2, const Blk &a1,
                                const Blk &a2, double &yi, double *wsd, int offz, int j1_start,
                                int j1_end)
{
    for (int j = j1_start; j <= j1_end; ++j) {
        const double piv = h1[0];
        if (fp_is_zero(piv)) {
            for (int q = 0; q < band1 - 1; ++q) h1[q] = h1[q + 1];
            h1[band1 - 1] = 0.0;
            continue;
        }
        double cs, sn, d = a1(j, 1), zj = W_(offz, j - 1);
        givens(piv, d, cs, sn);
        a1(j, 1) = d;
        rota(cs, sn, yi, zj);
        W_(offz, j - 1) = zj;
        for (int q = 1; q <= band2; ++q) {
            double v = a2(j, q);
            rota(cs, sn, h2[q - 1], v);
            a2(j, q) = v;
        }
        if (j < j1_end) {
            const int i2 = min(j1_end - j, band1 - 1) + 1;
            for (int q = 2; q <= i2; ++q) {
                double v = a1(j, q);
                rota(cs, sn, h1[q - 1], v);
                a1(j, q) = v;
            }
            for (int q = 0; q < i2 - 1; ++q) h1[q] = h1[q + 1];
            h1[i2 - 1] = 0.0;
        }
    }
    for (int j = 1; j <= band2; ++j) {
        const int ij = j1_end + j;
        if (ij <= 0) continue;
        const double piv = h2[j - 1];
        if (fp_is_zero(piv)) continue;
        double cs, sn, d = a2(ij, j), zj = W_(offz, ij - 1);
        givens(piv, d, cs, sn);
        a2(ij, j) = d;
        rota(cs, sn, yi, zj);
        W_(offz, ij - 1) = zj;
        for (int q = j + 1; q <= band2; ++q) {
            double v = a2(ij, q);
            rota(cs, sn, h2[q - 1], v);
            a2(ij, q) = v;
        }
    }
}

// fpbacp, core:1869-1921: a (band, width k1) and b (n x k); right-hand side at offz, result at offc
__device__ void backsub_periodic(const Blk &a, const Blk &b, double *wsd, int offz, int offc, int n, int k)
{
    const int n2 = n - k;
    int l = n;
    for (int i = 1; i <= k; ++i) {
        double store = W_(offz, l - 1);
        const int j = k + 2 - i;
        if (i != 1) {
            int l0 = l;
            for (int l1 = j; l1 <= k; ++l1) {
                ++l0;
                store = sub(store, mul(W_(offc, l0 - 1), b(l, l1)));
            }
        }
        W_(offc, l - 1) = store / b(l, j - 1);
        --l;
        if (l == 0) return;
    }
    for (int i = 1; i <= n2; ++i) {
        double store = W_(offz, i - 1);
        l = n2;
        for (int j = 1; j <= k; ++j) {
            ++l;
            store = sub(store, mul(W_(offc, l - 1), b(i, j)));
        }
        W_(offc, i - 1) = store;
    }
    int i = n2;
    W_(offc, i - 1) = W_(offc, i - 1) / a(i, 1);
    if (i == 1) return;
    for (int j = 2; j <= n2; ++j) {
        --i;
        double store = W_(offc, i - 1);
        const int i1 = (j <= k) ? j - 1 : k;
        l = i;
        for (int l0 = 1; l0 <= i1; ++l0) {
            ++l;
            store = sub(store, mul(W_(offc, l - 1), a(i, l0 + 1)));
        }
        W_(offc, i - 1) = store / a(i, 1);
    }
}

template <int K, bool HAS_W>
__device__ void percur_one(const PerParams &P, int b, double *__restrict__ wsd, int *__restrict__ wsi)
{
    constexpr int K1 = K + 1, K2 = K + 2;
    const int m = P.m, nest = P.nest, iopt = P.iopt;
    const double *xp = P.x + (size_t)b * P.x_sb;
    const double *yp = P.y + (size_t)b * P.y_sb;
    const double *wp = HAS_W ? P.w + (size_t)b * P.w_sb : nullptr;
    auto X = [&](int i) { return xp[(size_t)(i - 1) * P.x_si]; };
    auto Y = [&](int i) { return yp[(size_t)(i - 1) * P.y_si]; };
    auto Wt = [&](int i) { return HAS_W ? wp[(size_t)(i - 1) * P.w_si] : 1.0; };
    const double s = P.s[(size_t)b * P.s_stride];
    const int nmin = 2 * K1, m1 = m - 1, nmax = m + 2 * K;
    int n = P.n[b], ier = IER_INPUT;  // n is in/out: the early returns leave the caller's value
    double fp = 0.0;
    bool write_fit = false, write_fp = false, t_valid = true;

    // ---- percur argument checks that need the data, core:15817-15846 ----
    {
        bool bad = false;
        double prev = X(1);
        for (int i = 1; i <= m1; ++i) {
            const double cur = X(i + 1);
            if (HAS_W && Wt(i) <= 0.0) bad = true;
            if (prev >= cur) bad = true;
            prev = cur;
        }
        if (iopt >= 0) {
            if (s < 0.0 || (fp_is_zero(s) && nest < (m + 2 * K))) bad = true;
        } else {
            n = P.n[b];
            if (n <= nmin || n > nest) bad = true;
        }
        if (bad) {
            P.ier[b] = IER_INPUT;
            return;
        }
    }
    const double per = sub(X(m), X(1));
    if (iopt != 0) {
        const bool have = (iopt < 0) || (n >= nmin && n <= nest && n <= nmax);
        if (iopt > 0 && !have) n = nmin;
        if (have) {
            for (int i = 1; i <= n; ++i) T_(i) = P.t[(size_t)b * nest + (i - 1)];
            if (iopt > 0) {
                for (int i = 1; i <= n; ++i) {
                    FPI_(i) = P.fpint[(size_t)b * nest + (i - 1)];
                    NRD_(i) = P.nrdata[(size_t)b * nest + (i - 1)];
                }
            }
        }
        if (iopt < 0) {  // periodic extension of the user's knots + fpchep, core:15828-15845
            int j1 = K1, i1 = n - K, j2 = K1, i2 = n - K;
            T_(j1) = X(1);
            T_(i1) = X(m);
            for (int i = 1; i <= K; ++i) {
                ++i1; --i2; ++j1; --j2;
                T_(j2) = sub(T_(i2), per);
                T_(i1) = add(T_(j1), per);
            }
            for (int i = 1; i <= n; ++i) P.t[(size_t)b * nest + (i - 1)] = T_(i);
            // fpchep, core:2689-2770
            const int nk1 = n - K1, nk2 = nk1 + 1;
            int chk = IER_OK;
            if (nk1 < K1 || n > m + 2 * K) chk = IER_INPUT;
            if (chk == IER_OK) {
                int j = n;
                for (int i = 1; i <= K; ++i) {
                    if (T_(i) > T_(i + 1) || T_(j) < T_(j - 1)) chk = IER_INPUT;
                    --j;
                }
                for (int i = K2; i <= nk2; ++i)
                    if (T_(i) <= T_(i - 1)) chk = IER_INPUT;
                if (X(1) < T_(K1) || X(m) > T_(nk2)) chk = IER_INPUT;
            }
            if (chk == IER_OK) {
                int l1 = K1, l2 = 1, l = 1;
                bool brk = false;
                for (l = 1; l <= m && !brk; ++l) {
                    const double xi = X(l);
                    while (!(xi < T_(l1 + 1) || l == nk1)) {
                        ++l1;
                        ++l2;
                        if (l2 > K1) { brk = true; break; }
                    }
                    if (brk) break;
                }
                if (!brk) l = m;
                const double perk = sub(T_(nk2), T_(K1));
                chk = IER_INPUT;
                for (int i1 = 2; i1 <= l && chk != IER_OK; ++i1) {
                    int i = i1 - 1;
                    const int mm = i + m1;
                    bool ok = true;
                    for (int j = K1; j <= nk1 && ok; ++j) {
                        const double tj = T_(j), tl = T_(j + K1);
                        double xi = -DBL_MAX;
                        while (xi <= tj) {
                            ++i;
                            if (i > mm) { ok = false; break; }
                            const int i2 = i - m1;
                            xi = (i2 <= 0) ? X(i) : add(X(i2), perk);
                        }
                        if (ok && xi >= tl) ok = false;
                    }
                    if (ok) chk = IER_OK;
                }
            }
            if (chk != IER_OK) {
                P.ier[b] = chk;
                return;
            }
        }
    }

    // ---- fpperi, core:9668-10192 ----
    const double acc = mul(kTol, s);
    double fpold = 0.0, fp0 = 0.0, fpms = DBL_MAX;
    int nplus = 0, kk = K, kk1 = K1;
    ier = IER_OK;
    write_fit = true;
    const Blk a1{wsd, P.L.off_a1, K1}, a2{wsd, P.L.off_a2, K}, g1{wsd, P.L.off_g1, K2}, g2{wsd, P.L.off_g2, K1};
    int n7 = 0, n10 = 0, nk1 = 0;
    bool done = false;

    // fpperi_reset_interp, core:10220-10257
    auto reset_interp = [&]() -> bool {
        if (K % 2 != 0) {
            for (int i = 2; i <= m1; ++i) T_(K + i) = X(i);
            if (s <= 0.0) {
                kk = K - 1;
                kk1 = K;
                if (kk <= 0) {
                    T_(1) = sub(T_(m), per);
                    T_(2) = X(1);
                    T_(m + 1) = X(m);
                    T_(m + 2) = add(T_(3), per);
                    for (int i = 1; i <= m1; ++i) C_(i) = Y(i);
                    C_(m) = Y(1);
                    fp = 0.0;
                    FPI_(n - 1) = 0.0;
                    FPI_(n) = fp0;
                    NRD_(n) = 0;
                    return true;
                }
            }
        } else {
            for (int i = 2; i <= m1; ++i) T_(K + i) = mul(0.5, add(X(i), X(i - 1)));
        }
        return false;
    };

    if (iopt >= 0) {
        if (s <= 0.0 && nmax != nmin) {
            n = nmax;
            if (n > nest) {
                ier = IER_STORAGE;
                done = true;
                t_valid = false;  // knots untouched (core:9731-9734)
            } else if (reset_interp()) {
                ier = IER_INTERP;
                write_fp = true;
                done = true;
            }
        } else {
            if (iopt != 0 && n != nmin) {
                fp0 = FPI_(n);
                fpold = FPI_(n - 1);
                nplus = NRD_(n);
            }
            if (iopt == 0 || n == nmin || (iopt != 0 && n != nmin && s >= fp0)) {
                fp0 = 0.0;
                double d1 = 0.0, c1 = 0.0;
                for (int it = 1; it <= m1; ++it) {
                    const double wi = Wt(it);
                    double yi = mul(Y(it), wi), cs, sn;
                    givens(wi, d1, cs, sn);
                    rota(cs, sn, yi, c1);
                    fp0 = add(fp0, mul(yi, yi));
                }
                c1 = c1 / d1;
                fpms = sub(fp0, s);
                if (fpms < acc || nmax == nmin) {
                    ier = IER_LSQ;
                    for (int i = 1; i <= K1; ++i) {
                        T_(i) = sub(X(1), mul((double)(K1 - i), per));
                        C_(i) = c1;
                        T_(i + K1) = add(X(m), mul((double)(i - 1), per));
                    }
                    n = nmin;
                    fp = fp0;
                    FPI_(n - 1) = 0.0;
                    FPI_(n) = fp0;
                    NRD_(n) = 0;
                    write_fp = true;
                    done = true;
                } else {
                    fpold = fp0;
                    if (nmin >= nest) {
                        ier = IER_STORAGE;
                        done = true;
                        t_valid = false;
                    } else {
                        nplus = 1;
                        n = nmin + 1;
                        const int mm = (m + 1) / 2;
                        T_(K2) = X(mm);
                        NRD_(1) = mm - 2;
                        NRD_(2) = m1 - mm;
                    }
                }
            }
        }
    }

    bool part2 = false;
    int iter = 0;
    while (!done && iter < m) {  // update_knots, core:9818-10060
        ++iter;
        int nrint = n - nmin + 1;
        T_(K1) = X(1);
        nk1 = n - K1;
        const int nk2 = nk1 + 1;
        T_(nk2) = X(m);
        for (int j = 1; j <= K; ++j) {
            T_(nk2 + j) = add(T_(K1 + j), per);
            T_(K1 - j) = sub(T_(nk2 - j), per);
        }
        for (int i = 1; i <= nk1; ++i) {
            Z_(i) = 0.0;
            for (int j = 1; j <= kk1; ++j) A1_(i, j) = 0.0;
        }
        n7 = nk1 - K;
        n10 = n7 - kk;
        int jper = 0, l = K1;
        fp = 0.0;
        for (int it = 1; it <= m1; ++it) {  // get_coefs, core:9860-9937
            const double xi = X(it), wi = Wt(it);
            double yi = HAS_W ? mul(Y(it), wi) : Y(it);
            while (l < nk1 && xi >= T_(l + 1)) ++l;  // fp_knot_interval == linear walk
            double tw[2 * K], h[K1];
#pragma unroll
            for (int j = 0; j < 2 * K; ++j) tw[j] = T_(l - K + 1 + j);
            bspl<K>(tw, xi, h);
#pragma unroll
            for (int j = 0; j < K1; ++j) {
                Q_(it, j + 1) = h[j];
                if (HAS_W) h[j] = mul(h[j], wi);
            }
            const int l5 = l - K1;
            if (l5 >= n10) {
                if (jper == 0) {  // initialise a2 from the last columns of a1, core:9884-9897
                    for (int i = 1; i <= n7; ++i)
                        for (int j = 1; j <= kk; ++j) A2_(i, j) = 0.0;
                    int jk = n10 + 1;
                    for (int i = 1; i <= kk; ++i) {
                        int ik = jk;
                        for (int j = 1; j <= kk1; ++j) {
                            if (ik <= 0) break;
                            A2_(ik, i) = A1_(ik, j);
                            --ik;
                        }
                        ++jk;
                    }
                    jper = 1;
                }
                double h1[K2], h2[K1];
#pragma unroll
                for (int i = 0; i < K2; ++i) h1[i] = 0.0;
#pragma unroll
                for (int i = 0; i < K1; ++i) h2[i] = 0.0;
                int j = l5 - n10;
                for (int i = 1; i <= kk1; ++i) {
                    ++j;
                    int l0 = j, l1 = l0 - kk;
                    while (l1 > max(0, n10)) {
                        l0 = l1 - n10;
                        l1 = l0 - kk;
                    }
                    if (l1 > 0) h1[l1 - 1] = h[i - 1];
                    else h2[l0 - 1] = add(h2[l0 - 1], h[i - 1]);
                }
                rotate_row_2mat(h1, kk1, h2, kk, a1, a2, yi, wsd, P.L.off_z, 1, n10);
            } else {
                // fp_rotate_row, core:5691-5717, band kk1
                int j = l5;
                for (int i = 1; i <= kk1; ++i) {
                    ++j;
                    const double piv = h[i - 1];
                    if (fp_is_zero(piv)) continue;
                    double cs, sn, d = A1_(j, 1), zj = Z_(j);
                    givens(piv, d, cs, sn);
                    A1_(j, 1) = d;
                    rota(cs, sn, yi, zj);
                    Z_(j) = zj;
                    for (int q = 1; q <= kk1 - i; ++q) {
                        double v = A1_(j, 1 + q);
                        rota(cs, sn, h[i - 1 + q], v);
                        A1_(j, 1 + q) = v;
                    }
                }
            }
            fp = add(fp, mul(yi, yi));
        }
        FPI_(n - 1) = fpold;
        FPI_(n) = fp0;
        NRD_(n) = nplus;
        {
            const Blk a1k{wsd, P.L.off_a1, K1}, a2k{wsd, P.L.off_a2, K};
            backsub_periodic(a1k, a2k, wsd, P.L.off_z, P.L.off_c, n7, kk);
        }
        for (int i = 1; i <= K; ++i) C_(n7 + i) = C_(i);
        write_fp = true;
        fpms = sub(fp, s);
        if (iopt < 0 || fabs(fpms) < acc) { done = true; break; }
        if (fpms < 0.0) { part2 = true; break; }
        if (n == nmax) { ier = IER_INTERP; done = true; break; }
        if (n == nest) { ier = IER_STORAGE; done = true; break; }
        {
            int npl1 = nplus * 2;
            if (sub(fpold, fp) > acc) npl1 = x86_int(mul((double)nplus, fpms) / sub(fpold, fp));
            nplus = min(nplus * 2, max(max(npl1, nplus / 2), 1));
        }
        fpold = fp;
        {  // residual per knot interval, core:9979-10007
            double fpart = 0.0;
            int i = 1, newk = 0;
            l = K1;
            for (int it = 1; it <= m1; ++it) {
                if (X(it) >= T_(l)) { newk = 1; ++l; }
                double term = 0.0;
                int l0 = l - K2;
#pragma unroll
                for (int j = 1; j <= K1; ++j) {
                    ++l0;
                    term = add(term, mul(C_(l0), Q_(it, j)));
                }
                term = sub(term, Y(it));
                if (HAS_W) term = mul(Wt(it), term);
                term = mul(term, term);
                fpart = add(fpart, term);
                if (newk) {
                    if (l > K2) {
                        const double store = mul(term, 0.5);
                        FPI_(i) = sub(fpart, store);
                        ++i;
                        fpart = store;
                    } else {
                        FPI_(nrint) = term;
                    }
                    newk = 0;
                }
            }
            FPI_(nrint) = add(FPI_(nrint), fpart);
        }
        for (int lp = 1; lp <= nplus; ++lp) {  // add_new_knots, core:10009-10032 (fpknot 8199-8275)
            double fpmax = 0.0;
            int number = 0, maxpt = 0, maxbeg = 0, jbegin = 1;
            for (int j = 1; j <= nrint; ++j) {
                const int jpoint = NRD_(j);
                const double fj = FPI_(j);
                if (jpoint != 0 && (number == 0 || fj > fpmax)) {
                    fpmax = fj;
                    number = j;
                    maxpt = jpoint;
                    maxbeg = jbegin;
                }
                jbegin = jbegin + jpoint + 1;
            }
            if (number != 0) {
                const int ihalf = maxpt / 2 + 1, nrx = maxbeg + ihalf, next = number + 1;
                for (int jj = nrint; jj >= next; --jj) {
                    FPI_(jj + 1) = FPI_(jj);
                    NRD_(jj + 1) = NRD_(jj);
                    T_(jj + K + 1) = T_(jj + K);
                }
                NRD_(number) = ihalf - 1;
                NRD_(next) = maxpt - ihalf;
                const double am = (double)maxpt;
                FPI_(number) = mul(fpmax, (double)(ihalf - 1)) / am;
                FPI_(next) = mul(fpmax, (double)(maxpt - ihalf)) / am;
                T_(next + K) = X(nrx);
                ++n;
                ++nrint;
            }
            if (n == nmax) {
                if (reset_interp()) {
                    ier = IER_INTERP;
                    done = true;
                } else {
                    iter = 0;
                }
                break;
            }
            if (n == nest) break;
        }
    }
    if (!done && !part2) part2 = true;  // trial bound reached: the reference falls through to part 2

    if (part2 && !done) {
        // ---- part 2, core:10060-10190 ----
        // fpdisc, core:5453-5495
        {
            const int nrint = nk1 - K;
            const double fac = (double)nrint / sub(T_(nk1 + 1), T_(K1));
            for (int l = K2; l <= nk1; ++l) {
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
        double p = 0.0, p1 = 0.0, f1 = sub(fp0, s), p3 = -1.0, f3 = fpms, p2 = 0.0, f2 = 0.0;
        const int n11 = n10 - 1, n8 = n7 - 1;
        {
            int l = n7;
            for (int i = 1; i <= K; ++i) {
                p = add(p, A2_(l, K + 1 - i));
                --l;
                if (l == 0) break;
            }
            if (l != 0) {  // p + sum(a1(1:n10,1)), core:10081
                double sm = 0.0;
                for (int i = 1; i <= n10; ++i) sm = add(sm, A1_(i, 1));
                p = add(p, sm);
            }
            p = (double)n7 / p;
        }
        bool check1 = false, check3 = false;
        for (int it2 = 1; it2 <= kMaxIt; ++it2) {
            const double pinv = 1.0 / p;
            for (int i = 1; i <= n7; ++i) {
                C_(i) = Z_(i);
#pragma unroll
                for (int j = 1; j <= K1; ++j) G1_(i, j) = A1_(i, j);
                G1_(i, K2) = 0.0;
                G2_(i, 1) = 0.0;
#pragma unroll
                for (int j = 1; j <= K; ++j) G2_(i, j + 1) = A2_(i, j);
            }
            {
                int l = n10;
                for (int j = 1; j <= K1; ++j) {
                    if (l <= 0) break;
                    G2_(l, 1) = A1_(l, j);
                    --l;
                }
            }
            for (int it = 1; it <= n8; ++it) {  // n8_rows, core:10117-10162
                double yi = 0.0, h1[K2], h2[K1];
#pragma unroll
                for (int i = 0; i < K2; ++i) h1[i] = 0.0;
#pragma unroll
                for (int i = 0; i < K1; ++i) h2[i] = 0.0;
                int l;
                if (it <= n11) {
                    l = it;
                    int l0 = it;
                    for (int j = 1; j <= K2; ++j) {
                        if (l0 == n10) {
                            l0 = 1;
                            for (int l1 = j; l1 <= K2; ++l1) {
                                h2[l0 - 1] = mul(B_(it, l1), pinv);
                                ++l0;
                            }
                            break;
                        }
                        h1[j - 1] = mul(B_(it, j), pinv);
                        ++l0;
                    }
                } else {
                    l = 1;
                    int i = it - n10;
                    for (int j = 1; j <= K2; ++j) {
                        ++i;
                        int l0 = i, l1 = l0 - K1;
                        while (l1 > max(0, n11)) {
                            l0 = l1 - n11;
                            l1 = l0 - K1;
                        }
                        if (l1 > 0) h1[l1 - 1] = mul(B_(it, j), pinv);
                        else h2[l0 - 1] = add(h2[l0 - 1], mul(B_(it, j), pinv));
                    }
                }
                rotate_row_2mat(h1, K2, h2, K1, g1, g2, yi, wsd, P.L.off_c, l, n11);
            }
            backsub_periodic(g1, g2, wsd, P.L.off_c, P.L.off_c, n7, K1);
            for (int i = 1; i <= K; ++i) C_(n7 + i) = C_(i);
            fp = 0.0;
            {
                int l = K1;
                for (int it = 1; it <= m1; ++it) {
                    if (X(it) >= T_(l)) ++l;
                    const int l0 = l - K2;
                    double term = 0.0;
#pragma unroll
                    for (int j = 1; j <= K1; ++j) term = add(term, mul(C_(l0 + j), Q_(it, j)));
                    term = sub(term, Y(it));
                    if (HAS_W) term = mul(Wt(it), term);
                    fp = add(fp, mul(term, term));
                }
            }
            fpms = sub(fp, s);
            if (fabs(fpms) < acc) break;
            if (it2 == kMaxIt) { ier = IER_MAXIT; break; }
            // root_finding_iterate core:11641-11691 + fprati core:11996-12025
            {
                const double con1 = 0.1, con9 = 0.9, con4 = 0.04;
                bool ret = false, ok = true;
                p2 = p;
                f2 = fpms;
                if (!check3) {
                    if (sub(f2, f3) > acc) {
                        check3 = f2 < 0.0;
                    } else {
                        p3 = p2;
                        f3 = f2;
                        p = mul(p, con4);
                        if (p <= p1) p = add(mul(p1, con9), mul(p2, con1));
                        ret = true;
                    }
                }
                if (!ret && !check1) {
                    if (sub(f1, f2) > acc) {
                        check1 = f2 > 0.0;
                    } else {
                        p1 = p2;
                        f1 = f2;
                        p = p / con4;
                        if (p3 >= 0.0 && p >= p3) p = add(mul(p2, con1), mul(p3, con9));
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
                            p = -num / den;
                        } else {
                            const double num = sub(mul(mul(p1, sub(f1, f3)), f2),
                                                   mul(mul(p2, sub(f2, f3)), f1));
                            p = num / mul(sub(f1, f2), f3);
                        }
                        if (f2 >= 0.0) { p1 = p2; f1 = f2; } else { p3 = p2; f3 = f2; }
                    }
                }
                if (!ok) { ier = IER_S_SMALL; break; }
            }
        }
    }

    // ---- write back ----
    P.ier[b] = ier;
    if (write_fit) {
        P.n[b] = n;
        if (t_valid)
            for (int i = 1; i <= n; ++i) P.t[(size_t)b * nest + (i - 1)] = T_(i);
        if (write_fp) {
            P.fp[b] = fp;
            for (int i = 1; i <= n - K1; ++i) P.c[(size_t)b * nest + (i - 1)] = C_(i);
        }
        if (P.fpint && P.nrdata && iopt >= 0 && t_valid) {
            for (int i = 1; i <= n; ++i) {
                P.fpint[(size_t)b * nest + (i - 1)] = FPI_(i);
                P.nrdata[(size_t)b * nest + (i - 1)] = NRD_(i);
            }
        }
    }
}

template <int K, bool HAS_W>
__global__ void __launch_bounds__(kPerThreads, kPerBlocksPerSm) percur_kernel(PerParams P)
{
    const int lane = threadIdx.x & 31;
    const int warp_slot = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    double *wsd = P.ws + (size_t)warp_slot * P.L.per_lane * 32 + lane;
    int *wsi = P.wsi + (size_t)warp_slot * P.L.nestp * 32 + lane;
    for (;;) {
        unsigned tile = 0;
        if (lane == 0) tile = atomicAdd(P.counter, 1u);
        tile = __shfl_sync(0xffffffffu, tile, 0);
        const long long base = (long long)tile * 32;
        if (base >= P.nbatch) break;
        if (base + lane < P.nbatch) percur_one<K, HAS_W>(P, (int)(base + lane), wsd, wsi);
        __syncwarp();
    }
}

template <int K>
cudaError_t launch_percur_k(const PerParams &P, int grid, cudaStream_t st)
{
    if (P.w) percur_kernel<K, true><<<grid, kPerThreads, 0, st>>>(P);
    else percur_kernel<K, false><<<grid, kPerThreads, 0, st>>>(P);
    return cudaGetLastError();
}

struct Guard {
    int prev = -1;
    bool ok = false;
    explicit Guard(int dev) { ok = cudaGetDevice(&prev) == cudaSuccess && cudaSetDevice(dev) == cudaSuccess; }
    ~Guard()
    {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

}  // namespace
}  // namespace fpb

using namespace fpb;

extern "C" {

int32_t fp_percur_batch_dev_c(fpb_engine *eng, FP_SIZE nbatch, FP_SIZE iopt, FP_SIZE m, const FP_REAL *x,
                              int64_t x_sb, int64_t x_si, const FP_REAL *y, int64_t y_sb, int64_t y_si,
                              const FP_REAL *w, int64_t w_sb, int64_t w_si, FP_SIZE k, const FP_REAL *s,
                              FP_SIZE s_stride, FP_SIZE nest, FP_SIZE *n, FP_REAL *t, FP_REAL *c, FP_REAL *fp,
                              FP_FLAG *ier, FP_REAL *fpint, FP_SIZE *nrdata)
{
    if (!eng) {
        set_last_error("null engine");
        return FPB_ERR_ARG;
    }
    if (nbatch <= 0) return FPB_STATUS_OK;
    if (!x || !y || !s || !n || !t || !c || !fp || !ier) {
        set_last_error("fp_percur_batch_dev_c: null required pointer");
        return FPB_ERR_ARG;
    }
    if (iopt == 1 && (!fpint || !nrdata)) {
        set_last_error("fp_percur_batch_dev_c: iopt=1 needs the fpint/nrdata state arrays");
        return FPB_ERR_ARG;
    }
    Guard g(eng->device);
    if (!g.ok) return FPB_ERR_CUDA;
    // call-level part of percur's data check (core:15813-15815)
    if (k <= 0 || k > 5 || iopt < -1 || iopt > 1 || m < 2 || nest < 2 * (k + 1)) {
        if (!cuda_ok(launch_fill_i32(ier, FPB_INPUT_ERROR, nbatch, eng->stream), "fill ier")) return FPB_ERR_CUDA;
        eng->launches++;
        return FPB_STATUS_OK;
    }
    PerParams P;
    std::memset(&P, 0, sizeof P);
    P.L = per_layout(m, k, nest);
    const long long blocks_needed = ((long long)nbatch + kPerThreads - 1) / kPerThreads;
    const long long grid = std::min<long long>(blocks_needed, (long long)eng->nsm * kPerBlocksPerSm);
    const size_t warps = (size_t)grid * (kPerThreads / 32);
    if (!eng->ws.reserve(warps * (size_t)P.L.per_lane * 32 * sizeof(double)) ||
        !eng->wsi.reserve(warps * (size_t)P.L.nestp * 32 * sizeof(int)) || !eng->counter.reserve(256))
        return FPB_ERR_ALLOC;
    if (!cuda_ok(cudaMemsetAsync(eng->counter.ptr, 0, 64, eng->stream), "memset")) return FPB_ERR_CUDA;
    P.nbatch = nbatch; P.iopt = iopt; P.m = m; P.k = k; P.nest = nest;
    P.x = x; P.y = y; P.w = w;
    P.x_sb = x_sb; P.x_si = x_si; P.y_sb = y_sb; P.y_si = y_si; P.w_sb = w_sb; P.w_si = w_si;
    P.s = s; P.s_stride = s_stride;
    P.n = n; P.t = t; P.c = c; P.fp = fp; P.ier = ier; P.fpint = fpint; P.nrdata = nrdata;
    P.ws = eng->ws.as<double>(); P.wsi = eng->wsi.as<int>();
    P.counter = eng->counter.as<unsigned>();
    cudaError_t e = cudaErrorInvalidValue;
    if (eng->timed) cudaEventRecord(eng->ev0, eng->stream);
    switch (k) {
    case 1: e = launch_percur_k<1>(P, (int)grid, eng->stream); break;
    case 2: e = launch_percur_k<2>(P, (int)grid, eng->stream); break;
    case 3: e = launch_percur_k<3>(P, (int)grid, eng->stream); break;
    case 4: e = launch_percur_k<4>(P, (int)grid, eng->stream); break;
    case 5: e = launch_percur_k<5>(P, (int)grid, eng->stream); break;
    }
    if (eng->timed) cudaEventRecord(eng->ev1, eng->stream);
    if (!cuda_ok(e, "percur kernel")) return FPB_ERR_CUDA;
    eng->launches++;
    return FPB_STATUS_OK;
}

// Host-pointer batch, curve-major: x, y [nbatch][m], w [nbatch][m] or NULL, t/c [nbatch][nest],
// fpint/nrdata [nbatch][nest] continuation state (NULL allowed unless iopt = 1).
int32_t fp_percur_batch_c(FP_SIZE nbatch, FP_SIZE iopt, FP_SIZE m, const FP_REAL *x, const FP_REAL *y,
                          const FP_REAL *w, FP_SIZE k, FP_REAL s, FP_SIZE nest, FP_SIZE *n, FP_REAL *t,
                          FP_REAL *c, FP_REAL *fp, FP_FLAG *ier, FP_REAL *fpint, FP_SIZE *nrdata)
{
    if (nbatch <= 0) return FPB_STATUS_OK;
    if (!x || !y || !n || !t || !c || !fp || !ier) {
        set_last_error("fp_percur_batch_c: null required pointer");
        return FPB_ERR_ARG;
    }
    if (k <= 0 || k > 5 || iopt < -1 || iopt > 1 || m < 2 || nest < 2 * (k + 1)) {
        for (FP_SIZE b = 0; b < nbatch; ++b) ier[b] = FPB_INPUT_ERROR;
        return FPB_STATUS_OK;
    }
    if (iopt == 1 && (!fpint || !nrdata)) {
        set_last_error("fp_percur_batch_c: iopt=1 needs the fpint/nrdata state arrays");
        return FPB_ERR_ARG;
    }
    fpb_engine *eng = default_engine(-1);
    if (!eng) return FPB_ERR_CUDA;  // no device: there is no CPU path
    Guard g(eng->device);
    if (!g.ok) return FPB_ERR_CUDA;
    cudaStream_t st = eng->stream;
    const size_t per_curve = (size_t)m * 48 + (size_t)nest * 32 + 64;
    const long long chunk = std::max<long long>(128, (long long)(((size_t)6 << 30) / per_curve) / 128 * 128);
    enum { XC, XP, YC, YP, WC, WP, SS, NN, TT, CC, FP_IER };
    DevBuf *B = eng->scratch;
    const bool keep = fpint && nrdata;
    for (long long b0 = 0; b0 < nbatch; b0 += chunk) {
        const long long nb = std::min<long long>(chunk, nbatch - b0);
        const long long ld = (nb + 31) / 32 * 32;
        const size_t cm = (size_t)nb * m * 8, pm = (size_t)ld * m * 8;
        if (!B[XC].reserve(cm) || !B[XP].reserve(pm) || !B[YC].reserve(cm) || !B[YP].reserve(pm) ||
            !B[WC].reserve(cm) || !B[WP].reserve(pm) || !B[SS].reserve(64) || !B[NN].reserve((size_t)nb * 8) ||
            !B[TT].reserve((size_t)nb * nest * 20) || !B[CC].reserve((size_t)nb * nest * 8) ||
            !B[FP_IER].reserve((size_t)nb * 16))
            return FPB_ERR_ALLOC;
        double *d_t = B[TT].as<double>(), *d_fpint = d_t + (size_t)nb * nest, *d_c = B[CC].as<double>(),
               *d_fp = B[FP_IER].as<double>(), *d_s = B[SS].as<double>();
        int *d_n = B[NN].as<int>(), *d_ier = reinterpret_cast<int *>(d_fp + nb),
            *d_nrd = reinterpret_cast<int *>(d_fpint + (size_t)nb * nest);
        bool ok = cuda_ok(cudaMemcpyAsync(B[XC].ptr, x + (size_t)b0 * m, cm, cudaMemcpyHostToDevice, st), "H2D x");
        ok = ok && cuda_ok(cudaMemcpyAsync(B[YC].ptr, y + (size_t)b0 * m, cm, cudaMemcpyHostToDevice, st), "H2D y");
        if (w) ok = ok && cuda_ok(cudaMemcpyAsync(B[WC].ptr, w + (size_t)b0 * m, cm, cudaMemcpyHostToDevice, st), "H2D w");
        ok = ok && cuda_ok(cudaMemcpyAsync(d_s, &s, 8, cudaMemcpyHostToDevice, st), "H2D s");
        ok = ok && cuda_ok(cudaMemcpyAsync(d_n, n + b0, (size_t)nb * 4, cudaMemcpyHostToDevice, st), "H2D n");
        {
            // t and c are in/out in the reference: early returns leave the caller's values
            ok = ok && cuda_ok(cudaMemcpyAsync(d_c, c + (size_t)b0 * nest, (size_t)nb * nest * 8, cudaMemcpyHostToDevice, st), "H2D c");
            ok = ok && cuda_ok(cudaMemcpyAsync(d_t, t + (size_t)b0 * nest, (size_t)nb * nest * 8, cudaMemcpyHostToDevice, st), "H2D t");
        }
        if (iopt == 1) {
            ok = ok && cuda_ok(cudaMemcpyAsync(d_fpint, fpint + (size_t)b0 * nest, (size_t)nb * nest * 8, cudaMemcpyHostToDevice, st), "H2D fpint");
            ok = ok && cuda_ok(cudaMemcpyAsync(d_nrd, nrdata + (size_t)b0 * nest, (size_t)nb * nest * 4, cudaMemcpyHostToDevice, st), "H2D nrdata");
        }
        // the reference leaves fp untouched where fpperi does not set it (inout): seed with the caller's
        ok = ok && cuda_ok(cudaMemcpyAsync(d_fp, fp + b0, (size_t)nb * 8, cudaMemcpyHostToDevice, st), "H2D fp");
        if (!ok) return FPB_ERR_CUDA;
        ok = cuda_ok(launch_transpose(B[XC].as<double>(), m, B[XP].as<double>(), ld, nb, m, st), "transpose x");
        ok = ok && cuda_ok(launch_transpose(B[YC].as<double>(), m, B[YP].as<double>(), ld, nb, m, st), "transpose y");
        if (w) ok = ok && cuda_ok(launch_transpose(B[WC].as<double>(), m, B[WP].as<double>(), ld, nb, m, st), "transpose w");
        if (!ok) return FPB_ERR_CUDA;
        eng->launches += 2 + (w ? 1 : 0);
        const int32_t rc = fp_percur_batch_dev_c(eng, (FP_SIZE)nb, iopt, m, B[XP].as<double>(), 1, ld,
                                                 B[YP].as<double>(), 1, ld, w ? B[WP].as<double>() : nullptr, 1,
                                                 ld, k, d_s, 0, nest, d_n, d_t, d_c, d_fp, d_ier,
                                                 (keep || iopt == 1) ? d_fpint : nullptr,
                                                 (keep || iopt == 1) ? d_nrd : nullptr);
        if (rc != FPB_STATUS_OK) return rc;
        std::vector<int> hier((size_t)nb), hn((size_t)nb);
        ok = cuda_ok(cudaMemcpyAsync(hier.data(), d_ier, (size_t)nb * 4, cudaMemcpyDeviceToHost, st), "D2H ier");
        ok = ok && cuda_ok(cudaMemcpyAsync(hn.data(), d_n, (size_t)nb * 4, cudaMemcpyDeviceToHost, st), "D2H n");
        ok = ok && cuda_ok(cudaMemcpyAsync(fp + b0, d_fp, (size_t)nb * 8, cudaMemcpyDeviceToHost, st), "D2H fp");
        ok = ok && cuda_ok(cudaMemcpyAsync(t + (size_t)b0 * nest, d_t, (size_t)nb * nest * 8, cudaMemcpyDeviceToHost, st), "D2H t");
        ok = ok && cuda_ok(cudaMemcpyAsync(c + (size_t)b0 * nest, d_c, (size_t)nb * nest * 8, cudaMemcpyDeviceToHost, st), "D2H c");
        if (keep && iopt >= 0) {
            ok = ok && cuda_ok(cudaMemcpyAsync(fpint + (size_t)b0 * nest, d_fpint, (size_t)nb * nest * 8, cudaMemcpyDeviceToHost, st), "D2H fpint");
            ok = ok && cuda_ok(cudaMemcpyAsync(nrdata + (size_t)b0 * nest, d_nrd, (size_t)nb * nest * 4, cudaMemcpyDeviceToHost, st), "D2H nrdata");
        }
        ok = ok && cuda_ok(cudaStreamSynchronize(st), "sync");
        if (!ok) return FPB_ERR_CUDA;
        for (long long b = 0; b < nb; ++b) {
            ier[b0 + b] = hier[(size_t)b];
            // n is an output only where fpperi ran (the data checks leave it untouched)
            if (hier[(size_t)b] != FPB_INPUT_ERROR) n[b0 + b] = hn[(size_t)b];
        }
    }
    return FPB_STATUS_OK;
}

// the reference's own flat entry point: a batch of one curve
void fp_percur_c(FP_SIZE iopt, FP_SIZE m, const FP_REAL *x, const FP_REAL *y, const FP_REAL *w, FP_SIZE k,
                 FP_REAL s, FP_SIZE nest, FP_SIZE *n, FP_REAL *t, FP_REAL *c, FP_REAL *fp, FP_REAL *wrk,
                 FP_SIZE lwrk, FP_SIZE *iwrk, FP_FLAG *ier)
{
    *ier = FPB_INPUT_ERROR;  // data check that needs no data, core:15806-15816
    if (k <= 0 || k > 5) return;
    if (iopt < -1 || iopt > 1) return;
    if (m < 2 || nest < 2 * (k + 1)) return;
    if (lwrk < m * (k + 1) + nest * (8 + 5 * k)) return;
    // wrk(1:nest) is fpint and iwrk(1:nest) is nrdata (core:15851-15860): the iopt=1 state
    FP_FLAG ier1 = FPB_INPUT_ERROR;
    const int32_t rc = fp_percur_batch_c(1, iopt, m, x, y, w, k, s, nest, n, t, c, fp, &ier1, wrk, iwrk);
    *ier = (rc != FPB_STATUS_OK) ? rc : ier1;
}

}  // extern "C"
