// fpb_curfit_solo.cu -- ONE CURVE PER WARP: the fit behind the flat (batch-of-one) entry point fp_curfit_c and the
// handle API, where latency counts and there is no batch to fill the GPU with.
//
// Reference: curfit core:1240-1299 -> fpcurf core:4737-5087 with fpbspl 2387-2413, fp_knot_interval 2432-2473,
// fpgivs 5638-5658, fprota 12385-12401, fp_rotate_row 5691-5717, fp_rotate_shifted 5800-5820, fpback 1808-1838,
// fpdisc 5453-5495, fpknot 8199-8275, root_finding_iterate 11641-11691, fprati 11996-12025.
//
// The batch kernels (fpb_curfit.cu) give a curve to ONE thread: right for 10^6 curves, but a single fit is then one
// dependent chain of ~300 FP64 instructions per data point at ~8 cycles each (4 ms per call; the CPU takes 0.2 ms).
// Here a warp shares one curve, everything lives in shared memory, and the work is split three ways:
//   * per-point loops (knot interval, B-spline values, residuals) run one point per lane;
//   * the two band-QR sweeps -- data rows into (R, z), smoothing rows into (G, c) -- run as a WAVEFRONT: G lane groups
//     each own one row at a time, lane c of a group holds the row's c-th entry (one more lane its right-hand side),
//     and every tick each group does one Givens step against band row j -- all lanes of the group compute the same
//     (cos, sin) and each applies it to its own column.  Row r waits until row r-1 has left band row j, so every
//     rotation sees exactly the operands the reference's sequential loops give it: the results are the same bits.
//     A least-squares pass takes m + n ticks instead of m (k+1) rotations.  tests/solo_sim replays this schedule on the
//     CPU against the oracle's sequential sweeps;
//   * the short serial pieces (fpknot, fpback, the ordered sums) run on lane 0.
// Sums that the reference forms in point order (fp, fpint) are formed in point order.
//
// The same kernel, with ND right-hand sides per band row instead of one, is fppara (core:8822-9177, fp_rotate_row_vec
// 5742-5771, fp_rotate_shifted_vec 5845-5862) for fp_parcur_c: ND coordinate curves share one knot vector and one
// triangular factor; parcur's own argument checks and chord-length parameters (core:15238-15262) run in the kernel.
#include <algorithm>

#include "fpb_common.cuh"
#include "fpb_kernels.h"

namespace fpb {

namespace {

constexpr unsigned FULL = 0xffffffffu;

template <int K, int ND> struct SoloCfg {
    static constexpr int K1 = K + 1, K2 = K + 2;
    // data rows: K1 entries + ND right-hand sides per group.  Row r meets band row j at tick r * ST + j: with K1 groups
    // (ST = 1) a group is free again exactly when its next row starts; where fewer fit into the warp the rows start
    // ST ticks apart.
    static constexpr int LG_D = K1 + ND;
    static constexpr int G_D = (32 / LG_D) < K1 ? (32 / LG_D) : K1;
    static constexpr int ST_D = (K1 + G_D - 1) / G_D;
    static_assert(G_D >= 1 && G_D * ST_D >= K1, "a group must be free when its next data row starts");
    // smoothing rows: K2 entries + ND right-hand sides per group, as many groups as fit; start ticks from a table
    static constexpr int LG_S = K2 + ND;
    static constexpr int G_S = 32 / LG_S;
    static_assert(G_S >= 1, "one smoothing row must fit into the warp");
};

// fpback core:1808-1838, lane d solving coordinate d.  mat: rows of BAND band entries followed by the ND right-hand
// sides (row length BAND + ND); the solution of coordinate d goes to cv + d * ldc.
template <int BAND, int ND>
__device__ __forceinline__ void back_subst_solo(const double *mat, double *cv, int ldc, int nk1, int lane)
{
    constexpr int LD = BAND + ND;
    if (lane < ND) {
        double cw[BAND - 1];
#pragma unroll
        for (int q = 0; q < BAND - 1; ++q) cw[q] = 0.0;
        for (int i = nk1; i >= 1; --i) {
            double store = mat[(i - 1) * LD + BAND + lane];
            const int cnt = min(nk1 - i, BAND - 1);
#pragma unroll
            for (int q = 1; q <= BAND - 1; ++q)
                if (q <= cnt) store = sub(store, mul(cw[q - 1], mat[(i - 1) * LD + q]));
            const double ci = store / mat[(i - 1) * LD];
            cv[lane * ldc + i - 1] = ci;
#pragma unroll
            for (int q = BAND - 2; q >= 1; --q) cw[q] = cw[q - 1];
            cw[0] = ci;
        }
    }
    __syncwarp();
}

template <int K, int ND, bool PAR>
__global__ void __launch_bounds__(32) curfit_solo_kernel(SoloParams p)
{
    using Cfg = SoloCfg<K, ND>;
    constexpr int K1 = Cfg::K1, K2 = Cfg::K2, LDA = K1 + ND, LDG = K2 + ND;
    extern __shared__ double sm[];
    const int lane = threadIdx.x, b = blockIdx.x;
    const int m = p.m, nest = p.nest, iopt = p.iopt;
    const int ncap = min(nest, m + K1);   // the knot count never exceeds min(nest, m + k + 1): the shared arrays are that long
    // xs: abscissae (curfit) / parameters u (parcur); ys: the data, ND values per point
    double *xs = sm, *ys = xs + m, *ws = ys + (size_t)m * ND, *res = ws + m, *qs = res + (size_t)m * ND;
    double *ts = qs + (size_t)m * K1, *cc = ts + ncap, *fpi = cc + (size_t)ncap * ND;
    // band matrices with their right-hand sides as more columns: (R | z) rows of K1 + ND, (G | c) rows of K2 + ND
    double *as = fpi + ncap, *gs = as + (size_t)ncap * LDA, *bs = gs + (size_t)ncap * LDG;
    double *hw = bs + (size_t)ncap * K2;   // [m][K1 + ND]: the weighted data rows w * (basis values | data)
    int *lidx = reinterpret_cast<int *>(hw + (size_t)m * LDA), *llag = lidx + m, *nrd = llag + m, *sst = nrd + ncap;
#define T_(i) ts[(i) - 1]
#define CD_(d, i) cc[(d) * ncap + (i) - 1]
#define FPI_(i) fpi[(i) - 1]
#define NRD_(i) nrd[(i) - 1]
#define A_(i, j) as[((i) - 1) * LDA + ((j) - 1)]
#define B_(i, j) bs[((i) - 1) * K2 + ((j) - 1)]

    if (p.ier[b] != IER_OK) return;   // rejected by the host's data checks (core:1265-1285): nothing is written
    {
        const double *gy = p.y + (size_t)b * m * ND, *gw = p.w + (size_t)b * m;
        for (int i = lane; i < m; i += 32) ws[i] = gw[i];
        for (int i = lane; i < m * ND; i += 32) ys[i] = gy[i];
        if (!PAR || !p.chord) {
            const double *gx = p.x + (size_t)b * m;
            for (int i = lane; i < m; i += 32) xs[i] = gx[i];
        }
    }
    int n = 0;
    if (iopt != 0) {
        n = p.n[b];
        // -1: rejected below (nk1 > m fails fpchec, core:2519); garbage on iopt = 1: a fresh start
        if (PAR && (n < 2 * K1 || n > nest || n > m + K1)) n = (iopt < 0) ? -1 : 2 * K1;
        for (int i = lane; i < n; i += 32) ts[i] = p.t[(size_t)b * nest + i];
        if (iopt > 0)
            for (int i = lane; i < n; i += 32) {
                fpi[i] = p.fpint[(size_t)b * nest + i];
                nrd[i] = p.nrdata[(size_t)b * nest + i];
            }
    }
    __syncwarp();

    double xb_ = 0.0, xe_ = 0.0;
    if constexpr (!PAR) {
        xb_ = p.ends_from_x ? xs[0] : p.xb;
        xe_ = p.ends_from_x ? xs[m - 1] : p.xe;
    } else {
        // ---- parcur's own part, core:15238-15262: parameters, then the checks on the data ----
        int bad = 0;
        if (p.chord) {
            // u(i) = u(i-1) + norm2(x(:,i) - x(:,i-1)) (gfortran's scaled norm2), then u / u(m)
            for (int i = 1 + lane; i < m; i += 32) {
                double result = 0.0, scale = 1.0;
#pragma unroll
                for (int d = 0; d < ND; ++d) {
                    const double df = sub(ys[i * ND + d], ys[(i - 1) * ND + d]);
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
                res[i] = mul(scale, sqrt(result));
            }
            __syncwarp();
            double tot = 0.0;
            if (lane == 0) {
                xs[0] = 0.0;
                for (int i = 1; i < m; ++i) {
                    tot = add(tot, res[i]);
                    xs[i] = tot;
                }
            }
            tot = __shfl_sync(FULL, tot, 0);
            __syncwarp();
            if (tot <= 0.0) {
                bad = 1;
            } else {
                for (int i = 1 + lane; i < m; i += 32) xs[i] = xs[i] / tot;
                __syncwarp();
                if (lane == 0) xs[m - 1] = 1.0;
                xb_ = 0.0;
                xe_ = 1.0;
                __syncwarp();
            }
            for (int i = lane; i < m; i += 32) p.u_out[(size_t)b * m + i] = xs[i];
            if (!bad && lane == 0) {
                p.ube[2 * (size_t)b] = 0.0;
                p.ube[2 * (size_t)b + 1] = 1.0;
            }
        } else {
            xb_ = p.ube[2 * (size_t)b];
            xe_ = p.ube[2 * (size_t)b + 1];
        }
        if (!bad) {
            int mybad = (xb_ > xs[0] || xe_ < xs[m - 1] || ws[0] <= 0.0) ? 1 : 0;
            for (int i = lane; i + 1 < m; i += 32)
                if (xs[i] >= xs[i + 1] || ws[i + 1] <= 0.0) mybad = 1;
            bad = __any_sync(FULL, mybad) ? 1 : 0;
        }
        int chk = IER_OK;
        if (!bad) {
            if (iopt < 0) {
                if (n < 0) {
                    bad = 1;
                } else {
                    if (lane < K1) {   // visible to the caller even if fpchec fails, core:15264-15270
                        ts[lane] = xb_;
                        ts[n - 1 - lane] = xe_;
                        p.t[(size_t)b * nest + lane] = xb_;
                        p.t[(size_t)b * nest + n - 1 - lane] = xe_;
                    }
                    __syncwarp();
                    if (lane == 0) {   // fpchec, core:2507-2570
                        const int nk1c = n - K1, nk2 = nk1c + 1;
                        auto X = [&](int i) { return xs[i - 1]; };
                        bool f = (nk1c < K1 || nk1c > m);
                        if (!f) {
                            int j = n;
                            for (int i = 1; i <= K; ++i) {
                                if (T_(i) > T_(i + 1) || T_(j) < T_(j - 1)) f = true;
                                --j;
                            }
                            for (int i = K2; i <= nk2; ++i)
                                if (T_(i) <= T_(i - 1)) f = true;
                            if (X(1) < T_(K1) || X(m) > T_(nk2)) f = true;
                            if (X(1) >= T_(K2) || X(m) <= T_(nk1c)) f = true;
                        }
                        if (!f) {
                            int i = 1, l = K2;
                            const int nk3 = nk1c - 1;
                            for (int j = 2; j <= nk3 && !f; ++j) {
                                const double tj = T_(j);
                                ++l;
                                const double tl = T_(l);
                                while (i < m && X(i) <= tj) {
                                    ++i;
                                    if (i >= m) {
                                        f = true;
                                        break;
                                    }
                                }
                                if (!f && X(i) >= tl) f = true;
                            }
                        }
                        chk = f ? IER_INPUT : IER_OK;
                    }
                    chk = __shfl_sync(FULL, chk, 0);
                }
            } else if (p.s < 0.0 || (fp_is_zero(p.s) && nest < m + K1)) {
                bad = 1;
            }
        }
        if (bad || chk != IER_OK) {
            if (lane == 0) p.ier[b] = IER_INPUT;
            return;
        }
    }

    // scalars: every lane carries the same values (pure scalar arithmetic is simply done 32 times over)
    const double s = p.s, xb = xb_, xe = xe_;
    const int nmax = m + K1, nmin = 2 * K1;
    const double acc = mul(kTol, s);
    double fpold = 0.0, fp0 = 0.0, fp = p.fp[b], fpms = DBL_MAX;
    int nplus = 0, ier = IER_OK, iter = 0;
    bool done = false;

    // interpolation knots, core:4806-4816 / 4975-4985
    auto interp_knots = [&]() {
        const int mk1 = m - K1, k3 = K / 2;
        for (int l = lane + 1; l <= mk1; l += 32) {
            const int i = K2 + l - 1, j = k3 + 2 + l - 1;
            T_(i) = (k3 * 2 != K) ? xs[j - 1] : mul(add(xs[j - 1], xs[j - 2]), 0.5);
        }
    };

    // ---- fpcurf prologue, core:4793-4843 ----
    if (iopt >= 0) {
        if (s <= 0.0) {
            n = nmax;
            if (nmax > nest) {
                ier = IER_STORAGE;
                done = true;
            } else {
                interp_knots();
            }
        } else if (iopt != 0 && n != nmin) {
            fp0 = FPI_(n);
            fpold = FPI_(n - 1);
            nplus = NRD_(n);
            __syncwarp();
            if (fp0 <= s) {
                n = nmin;
                fpold = 0.0;
                nplus = 0;
                if (lane == 0) NRD_(1) = m - 2;
            }
        } else {
            n = nmin;
            fpold = 0.0;
            nplus = 0;
            if (lane == 0) NRD_(1) = m - 2;
        }
    }
    __syncwarp();
    const bool untouched = done;   // ier = 1 before anything was computed: t, c, fp stay the caller's

    // ---- main loop, core:4847-4998 ----
    int nk1 = n - K1;
    while (!done) {
        if (!(iter <= m)) break;
        ++iter;
        if (n == nmin) ier = IER_LSQ;
        int nrint = n - nmin + 1;
        nk1 = n - K1;
        if (lane < K1) {
            ts[lane] = xb;
            ts[nk1 + lane] = xe;
        }
        for (int i = lane; i < nk1 * LDA; i += 32) as[i] = 0.0;
        __syncwarp();
        // knot interval and B-spline values of every point, one point per lane (core:4876-4885)
        for (int it = lane; it < m; it += 32) {
            const double xi = xs[it];
            int lo = K1, hi = nk1;
            while (hi > lo) {
                const int mid = (lo + hi) >> 1;
                if (T_(mid + 1) <= xi) lo = mid + 1; else hi = mid;
            }
            lidx[it] = lo;
            double tw[2 * K], h[K1];
#pragma unroll
            for (int j = 0; j < 2 * K; ++j) tw[j] = ts[lo - K + j];
            bspl<K>(tw, xi, h);
            const double wi = ws[it];
#pragma unroll
            for (int j = 0; j < K1; ++j) {
                qs[it * K1 + j] = h[j];
                hw[it * LDA + j] = mul(wi, h[j]);
            }
#pragma unroll
            for (int d = 0; d < ND; ++d) hw[it * LDA + K1 + d] = mul(ys[it * ND + d], wi);
        }
        __syncwarp();
        // the residual loops' own interval counter (core:4947, 5065) advances at most one knot per point:
        // l(it) = min(l(it-1) + 1, lidx(it) + 1), l(0) = k2  ==>  l(it) = it + min(k2, min_{j<=it}(lidx(j) + 1 - j))
        {
            int carry = K2;
            for (int base = 0; base < m; base += 32) {
                const int it = base + lane;
                int v = it < m ? lidx[it] + 1 - (it + 1) : INT_MAX;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const int o = __shfl_up_sync(FULL, v, d);
                    if (lane >= d) v = min(v, o);
                }
                v = min(v, carry);
                if (it < m) llag[it] = it + 1 + v;
                carry = __shfl_sync(FULL, v, 31);
            }
        }
        // ---- data rows into (R, z) as a wavefront (core:4873-4896) ----
        // The tick is straight-line code: every load is unconditional (idle lanes read band row 1), the rotation is
        // computed by every lane (idle lanes rotate (1, 1)), only the stores are predicated, and the row after the
        // group's next one is fetched at the top of the tick, so no lane waits on shared memory when rows change.
        {
            constexpr int LG = Cfg::LG_D, G = Cfg::G_D, ST = Cfg::ST_D, NEVER = INT_MAX / 2;
            const int grp = lane / LG, sub_ = lane - grp * LG;
            const bool member = grp < G, isrhs = sub_ >= K1;
            const int base = member ? grp * LG : 0;
            int row = member ? grp : NEVER, first = 0, start = NEVER, first_n = 0;
            double hv = 0.0, hv_n = 0.0;
            if (row < m) {
                first = lidx[row] - K;
                start = row * ST + first;
                hv = hw[row * LG + sub_];
            }
            if (member && row + G < m) {
                first_n = lidx[row + G] - K;
                hv_n = hw[(row + G) * LG + sub_];
            }
            const int tau_end = (m - 1) * ST + lidx[m - 1];
            for (int tau = lidx[0] - K; tau <= tau_end; ++tau) {
                // the row two ahead (the next one is in first_n, hv_n)
                const int r2 = min(row, m) + 2 * G < m ? row + 2 * G : m - 1;
                const int first_2 = lidx[r2] - K;
                const double hv_2 = hw[r2 * LG + sub_];
                const int i0 = tau - start;                    // the row's entry that is the pivot now
                const bool act = (unsigned)i0 <= (unsigned)K;
                const int i0c = act ? i0 : 0;
                double *rowp = as + (act ? first + i0 - 1 : 0) * LG;   // band row it meets: (R(j, :) | z(j))
                const int q = sub_ - i0c;                      // 0: the pivot's lane, > 0: a column to its right
                const int col = isrhs ? sub_ : max(q, 0);
                const double wwl = rowp[0], s2 = rowp[col];
                const double piv = __shfl_sync(FULL, hv, base + i0c);
                const bool rot = act && !fp_is_zero(piv);
                const bool upd = rot && (isrhs || q >= 0);
                double ww = rot ? wwl : 1.0, cs, sn;
                givens_fast(rot ? piv : 1.0, ww, cs, sn);
                const double nb = add(mul(cs, s2), mul(sn, hv));
                const double na = sub(mul(cs, hv), mul(sn, s2));
                if (upd) rowp[col] = col == 0 ? ww : nb;
                hv = (upd && col != 0) ? na : hv;
                const bool fin = act && i0 == K;               // the row is through: its residual, then the next row
                if (fin && isrhs) res[row * ND + (sub_ - K1)] = mul(hv, hv);
                if (fin) {
                    row += G;
                    first = first_n;
                    hv = hv_n;
                    start = row < m ? row * ST + first : NEVER;
                    first_n = first_2;
                    hv_n = hv_2;
                }
                __syncwarp();
            }
            fp = 0.0;                                          // summed in row order (core:4894 / 8972: fp + sum(xi**2))
            for (int it = 0; it < m; ++it) {
                double sq = res[it * ND];
#pragma unroll
                for (int d = 1; d < ND; ++d) sq = add(sq, res[it * ND + d]);
                fp = add(fp, sq);
            }
            __syncwarp();
        }
        if (ier == IER_LSQ) fp0 = fp;
        if (lane == 0) {
            FPI_(n - 1) = fpold;
            FPI_(n) = fp0;
            NRD_(n) = nplus;
        }
        back_subst_solo<K1, ND>(as, cc, ncap, nk1, lane);
        if (iopt < 0) {
            done = true;
            break;
        }
        fpms = sub(fp, s);
        if (fabs(fpms) < acc) {
            done = true;
            break;
        }
        if (fpms < 0.0) break;
        if (n == nmax) {
            ier = IER_INTERP;
            done = true;
            break;
        }
        if (n == nest) {
            ier = IER_STORAGE;
            done = true;
            break;
        }
        // number of knots to add, core:4927-4935
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
        // squared residual of every point (core:4948-4953), then the per-interval split in point order (4954-4966)
        for (int it = lane; it < m; it += 32) {
            const int l0 = llag[it] - K2;
            if constexpr (!PAR) {
                double term = 0.0;
#pragma unroll
                for (int j = 1; j <= K1; ++j) term = add(term, mul(CD_(0, l0 + j), qs[it * K1 + j - 1]));
                term = mul(ws[it], sub(term, ys[it]));
                res[it] = mul(term, term);
            } else {   // core:9030-9041
                double term = 0.0;
#pragma unroll
                for (int d = 0; d < ND; ++d) {
                    double fac = 0.0;
#pragma unroll
                    for (int j = 1; j <= K1; ++j) fac = add(fac, mul(CD_(d, l0 + j), qs[it * K1 + j - 1]));
                    const double store = mul(ws[it], sub(fac, ys[it * ND + d]));
                    term = add(term, mul(store, store));
                }
                res[it] = term;
            }
        }
        __syncwarp();
        int flag = 0;
        if (lane == 0) {
            double fpart = 0.0;
            int i = 1, prev = K2;
            for (int it = 0; it < m; ++it) {
                const double term = res[it];
                const int l = llag[it];
                fpart = add(fpart, term);
                if (l != prev) {
                    const double store = mul(term, 0.5);
                    FPI_(i) = sub(fpart, store);
                    ++i;
                    fpart = store;
                }
                prev = l;
            }
            FPI_(nrint) = fpart;
            // add_new_knots core:4968-4996, fpknot core:8199-8275
            for (int lp = 1; lp <= nplus; ++lp) {
                double fpmax = 0.0;
                int number = 0, maxpt = 0, maxbeg = 0, jbegin = 1;
                for (int jj = 1; jj <= nrint; ++jj) {
                    const int jpoint = NRD_(jj);
                    if (jpoint != 0 && (number == 0 || FPI_(jj) > fpmax)) {
                        fpmax = FPI_(jj);
                        number = jj;
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
                    double an = (double)NRD_(number);
                    FPI_(number) = mul(fpmax, an) / am;
                    an = (double)NRD_(next);
                    FPI_(next) = mul(fpmax, an) / am;
                    T_(next + K) = xs[nrx - 1];
                    ++n;
                    ++nrint;
                }
                if (n == nmax) {
                    flag = 1;
                    break;
                }
                if (n == nest) break;
            }
        }
        n = __shfl_sync(FULL, n, 0);
        flag = __shfl_sync(FULL, flag, 0);
        __syncwarp();
        if (flag) {   // as many knots as an interpolating spline has: take its knots and start over (core:4975-4988)
            interp_knots();
            iter = 0;
            __syncwarp();
        }
    }

    // ---- part 2: the smoothing spline sp(x) with fp(p) = s, core:5002-5084 ----
    if (!done && ier != IER_LSQ) {
        nk1 = n - K1;
        // fpdisc core:5453-5495, one band row per lane
        {
            const double fac = (double)(nk1 - K) / sub(T_(nk1 + 1), T_(K1));
            for (int l = K2 + lane; l <= nk1; l += 32) {
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
        double p1 = 0.0, f1 = sub(fp0, s), p3 = -1.0, f3 = fpms, p2 = 0.0, f2 = 0.0, pp = 0.0;
        bool check1 = false, check3 = false;
        for (int i = 1; i <= nk1; ++i) pp = add(pp, A_(i, 1));
        pp = PAR ? pp / (double)nk1 : (double)nk1 / pp;   // core:5033 / 9119
        const int n8 = n - nmin;
        // start ticks of the smoothing rows: two ticks behind the row before, and not before the group is free
        // (row r meets band rows r+1 .. nk1, one per tick)
        if (lane == 0) {
            constexpr int G = Cfg::G_S;
            for (int r = 0; r < n8; ++r) {
                int st = r == 0 ? 0 : sst[r - 1] + 2;
                if (r >= G) st = max(st, sst[r - G] + nk1 - (r - G));
                sst[r] = st;
            }
        }
        __syncwarp();
        int it2 = 0;
        for (;;) {
            if (!(it2 < kMaxIt)) {
                ier = IER_MAXIT;
                break;
            }
            ++it2;
            const double pinv = 1.0 / pp;
            // c := z, g := (a | 0), core:5047-5049
            for (int i = lane; i < nk1 * LDG; i += 32) {
                const int r = i / LDG, q = i - r * LDG;
                gs[i] = q < K1 ? as[r * LDA + q] : (q >= K2 ? as[r * LDA + K1 + (q - K2)] : 0.0);
            }
            __syncwarp();
            // ---- smoothing rows b(it,:)/p into (G, c) as a wavefront (core:5052-5057, fp_rotate_shifted 5800-5820) ----
            {
                constexpr int LG = Cfg::LG_S, G = Cfg::G_S, NEVER = INT_MAX / 2;
                const int grp = lane / LG, sub_ = lane - grp * LG;
                const bool member = grp < G, isrhs = sub_ >= K2;
                const int base = member ? grp * LG : 0;
                const int bcol = isrhs ? 0 : sub_;               // column of b this lane reads
                // lane sub_ < K2 holds the row's entry of band column `col`; the pivot's lane then takes the column
                // that enters the band with value 0 (the reference shifts the row left instead)
                int row = member ? grp : NEVER, start = NEVER, start_n = NEVER, col = 0;
                double hv = 0.0, hv_n = 0.0;
                if (row < n8) {
                    start = sst[row];
                    col = row + 1 + sub_;
                    hv = isrhs ? 0.0 : mul(bs[row * K2 + bcol], pinv);
                }
                if (member && row + G < n8) {
                    start_n = sst[row + G];
                    hv_n = isrhs ? 0.0 : mul(bs[(row + G) * K2 + bcol], pinv);
                }
                const int tau_end = sst[n8 - 1] + nk1 - n8;
                for (int tau = 0; tau <= tau_end; ++tau) {
                    const bool has2 = min(row, n8) + 2 * G < n8;
                    const int r2 = has2 ? row + 2 * G : n8 - 1;
                    const int start_2 = has2 ? sst[r2] : NEVER;
                    const double hb_2 = bs[r2 * K2 + bcol];
                    const int d0 = tau - start;
                    const bool act = d0 >= 0;                    // start = NEVER once the group has no row left
                    const int j = act ? row + 1 + d0 : 1;        // band row the row meets now
                    double *rowp = gs + (j - 1) * LG;            // (G(j, :) | c(j))
                    const int ps = act ? d0 % K2 : 0;            // lane of the entry in column j
                    const int noff = min(nk1 - j, K2 - 1);
                    const int dc = col - j;
                    const bool pivlane = !isrhs && sub_ == ps;
                    const bool mine = isrhs || pivlane || (dc >= 1 && dc <= noff);
                    const int tc = isrhs ? sub_ : (mine && !pivlane ? dc : 0);
                    const double wwl = rowp[0], s2 = rowp[tc];
                    const double piv = __shfl_sync(FULL, hv, base + ps);
                    const bool rot = act && (PAR || !fp_is_zero(piv));   // the vector variant has no zero-pivot skip
                    const bool upd = rot && mine;
                    double ww = rot ? wwl : 1.0, cs, sn;
                    givens_fast(rot ? piv : 1.0, ww, cs, sn);
                    const double nb = add(mul(cs, s2), mul(sn, hv));
                    const double na = sub(mul(cs, hv), mul(sn, s2));
                    if (upd) rowp[tc] = pivlane ? ww : nb;
                    hv = (upd && !pivlane) ? na : hv;
                    if (act && pivlane) {
                        hv = 0.0;
                        col = j + K2;
                    }
                    if (act && j == nk1) {
                        row += G;
                        start = row < n8 ? start_n : NEVER;
                        hv = hv_n;
                        col = row + 1 + sub_;
                        start_n = start_2;
                        hv_n = isrhs ? 0.0 : mul(hb_2, pinv);
                    }
                    __syncwarp();
                }
            }
            back_subst_solo<K2, ND>(gs, cc, ncap, nk1, lane);
            // fp = sum of squared residuals in point order, core:5061-5069
            for (int it = lane; it < m; it += 32) {
                const int l0 = llag[it] - K2;
                if constexpr (!PAR) {
                    double term = 0.0;
#pragma unroll
                    for (int j = 1; j <= K1; ++j) term = add(term, mul(CD_(0, l0 + j), qs[it * K1 + j - 1]));
                    term = mul(ws[it], sub(term, ys[it]));
                    res[it] = mul(term, term);
                } else {   // core:9147-9158: term * w**2
                    double term = 0.0;
#pragma unroll
                    for (int d = 0; d < ND; ++d) {
                        double fac = 0.0;
#pragma unroll
                        for (int j = 1; j <= K1; ++j) fac = add(fac, mul(CD_(d, l0 + j), qs[it * K1 + j - 1]));
                        const double store = sub(fac, ys[it * ND + d]);
                        term = add(term, mul(store, store));
                    }
                    res[it] = mul(term, mul(ws[it], ws[it]));
                }
            }
            __syncwarp();
            fp = 0.0;
            for (int it = 0; it < m; ++it) fp = add(fp, res[it]);
            __syncwarp();
            fpms = sub(fp, s);
            if (fabs(fpms) < acc) break;
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
                        const double num = add(add(mul(mul(p1, p2), h3), mul(mul(p2, p3), h1)), mul(mul(p3, p1), h2));
                        const double den = add(add(mul(p1, h1), mul(p2, h2)), mul(p3, h3));
                        pp = -num / den;
                    } else {
                        const double num = sub(mul(mul(p1, sub(f1, f3)), f2), mul(mul(p2, sub(f2, f3)), f1));
                        pp = num / mul(sub(f1, f2), f3);
                    }
                    if (f2 >= 0.0) { p1 = p2; f1 = f2; } else { p3 = p2; f3 = f2; }
                }
            }
            if (!ok) {
                ier = IER_S_SMALL;
                break;
            }
        }
    }

    // ---- results ----
    __syncwarp();
    if (lane == 0) {
        p.n[b] = n;
        p.ier[b] = ier;
        if (!untouched) p.fp[b] = fp;
    }
    if (!untouched) {
        nk1 = n - K1;
        for (int i = lane; i < n; i += 32) {
            p.t[(size_t)b * nest + i] = ts[i];
            p.fpint[(size_t)b * nest + i] = fpi[i];
            p.nrdata[(size_t)b * nest + i] = nrd[i];
        }
        // curfit: c(1:nk1); parcur: coordinate d at c(d*n + 1 : d*n + nk1) (core:8981-8985)
        for (int d = 0; d < ND; ++d)
            for (int i = lane; i < nk1; i += 32) p.c[(size_t)b * p.c_stride + (size_t)d * n + i] = cc[d * ncap + i];
    }
#undef T_
#undef CD_
#undef FPI_
#undef NRD_
#undef A_
#undef B_
}

}  // namespace

size_t solo_smem_bytes(int m, int k, int nest, int idim)
{
    const size_t k1 = k + 1, k2 = k + 2, nd = idim;
    const size_t ncap = (size_t)std::min(nest, m + k + 1);
    const size_t doubles = (size_t)m * (2 + 2 * nd + k1 + k1 + nd) + ncap * (2 + nd + (k1 + nd) + (k2 + nd) + k2);
    return doubles * 8 + ((size_t)2 * m + 2 * ncap) * 4;
}

bool solo_supports(int k, int idim, int par)
{
    if (k < 1 || k > 5) return false;
    return par ? (idim == 2 || idim == 3) : idim == 1;
}

namespace {
template <int K, int ND, bool PAR>
cudaError_t launch_solo_one(const SoloParams &p, size_t smem, int dev, cudaStream_t st)
{
    static size_t opted[64] = {0};   // the attribute is per device
    if (smem > opted[dev & 63]) {
        const cudaError_t err = cudaFuncSetAttribute(curfit_solo_kernel<K, ND, PAR>,
                                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (err != cudaSuccess) return err;
        opted[dev & 63] = smem;
    }
    curfit_solo_kernel<K, ND, PAR><<<p.nbatch, 32, smem, st>>>(p);
    return cudaGetLastError();
}

template <int K>
cudaError_t launch_solo_k(const SoloParams &p, size_t smem, int dev, cudaStream_t st)
{
    if (!p.par) return launch_solo_one<K, 1, false>(p, smem, dev, st);
    if (p.idim == 2) return launch_solo_one<K, 2, true>(p, smem, dev, st);
    if (p.idim == 3) return launch_solo_one<K, 3, true>(p, smem, dev, st);
    return cudaErrorInvalidValue;
}
}  // namespace

cudaError_t launch_curfit_solo(const SoloParams &p, cudaStream_t st)
{
    if (!solo_supports(p.k, p.idim, p.par)) return cudaErrorInvalidValue;
    const size_t smem = solo_smem_bytes(p.m, p.k, p.nest, p.idim);
    int dev = 0;
    const cudaError_t err = cudaGetDevice(&dev);
    if (err != cudaSuccess) return err;
    switch (p.k) {
    case 1: return launch_solo_k<1>(p, smem, dev, st);
    case 2: return launch_solo_k<2>(p, smem, dev, st);
    case 3: return launch_solo_k<3>(p, smem, dev, st);
    case 4: return launch_solo_k<4>(p, smem, dev, st);
    case 5: return launch_solo_k<5>(p, smem, dev, st);
    default: return cudaErrorInvalidValue;
    }
}

}  // namespace fpb
