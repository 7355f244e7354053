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
#include "fpb_common.cuh"
#include "fpb_kernels.h"

namespace fpb {

namespace {

constexpr unsigned FULL = 0xffffffffu;

template <int K> struct SoloCfg {
    static constexpr int K1 = K + 1, K2 = K + 2;
    // data rows: K1 entries + right-hand side per group.  Row r meets band row j at tick r * ST + j: with K1 groups
    // (ST = 1) a group is free again exactly when its next row starts; where only fewer fit into the warp (k = 5)
    // the rows start two ticks apart.
    static constexpr int LG_D = K1 + 1;
    static constexpr int G_D = (32 / LG_D) < K1 ? (32 / LG_D) : K1;
    static constexpr int ST_D = G_D >= K1 ? 1 : 2;
    static_assert(G_D * ST_D >= K1, "a group must be free when its next data row starts");
    // smoothing rows: K2 entries + right-hand side per group, as many groups as fit; start ticks from a table
    static constexpr int LG_S = K2 + 1;
    static constexpr int G_S = 32 / LG_S;
};

// fpback core:1808-1838 on lane 0.  mat: rows of BAND band entries followed by the right-hand side (row length BAND + 1)
template <int BAND>
__device__ __forceinline__ void back_subst_solo(const double *mat, double *cv, int nk1, int lane)
{
    constexpr int LD = BAND + 1;
    if (lane == 0) {
        double cw[BAND - 1];
#pragma unroll
        for (int q = 0; q < BAND - 1; ++q) cw[q] = 0.0;
        for (int i = nk1; i >= 1; --i) {
            double store = mat[(i - 1) * LD + BAND];
            const int cnt = min(nk1 - i, BAND - 1);
#pragma unroll
            for (int q = 1; q <= BAND - 1; ++q)
                if (q <= cnt) store = sub(store, mul(cw[q - 1], mat[(i - 1) * LD + q]));
            const double ci = store / mat[(i - 1) * LD];
            cv[i - 1] = ci;
#pragma unroll
            for (int q = BAND - 2; q >= 1; --q) cw[q] = cw[q - 1];
            cw[0] = ci;
        }
    }
    __syncwarp();
}

template <int K>
__global__ void __launch_bounds__(32) curfit_solo_kernel(SoloParams p)
{
    using Cfg = SoloCfg<K>;
    constexpr int K1 = Cfg::K1, K2 = Cfg::K2;
    extern __shared__ double sm[];
    const int lane = threadIdx.x, b = blockIdx.x;
    const int m = p.m, nest = p.nest, iopt = p.iopt;
    double *xs = sm, *ys = xs + m, *ws = ys + m, *res = ws + m, *qs = res + m;
    double *ts = qs + (size_t)m * K1, *cc = ts + nest, *fpi = cc + nest;
    // band matrices with their right-hand sides as one more column: (R | z) rows of K1 + 1, (G | c) rows of K2 + 1
    double *as = fpi + nest, *gs = as + (size_t)nest * (K1 + 1), *bs = gs + (size_t)nest * (K2 + 1);
    double *hw = bs + (size_t)nest * K2;   // [m][K1 + 1]: the weighted data rows w * (basis values | y)
    int *lidx = reinterpret_cast<int *>(hw + (size_t)m * (K1 + 1)), *llag = lidx + m, *nrd = llag + m, *sst = nrd + nest;
#define T_(i) ts[(i) - 1]
#define C_(i) cc[(i) - 1]
#define FPI_(i) fpi[(i) - 1]
#define NRD_(i) nrd[(i) - 1]
#define A_(i, j) as[((i) - 1) * (K1 + 1) + ((j) - 1)]
#define B_(i, j) bs[((i) - 1) * K2 + ((j) - 1)]

    {
        const double *gx = p.x + (size_t)b * m, *gy = p.y + (size_t)b * m, *gw = p.w + (size_t)b * m;
        for (int i = lane; i < m; i += 32) {
            xs[i] = gx[i];
            ys[i] = gy[i];
            ws[i] = gw[i];
        }
    }
    if (p.ier[b] != IER_OK) return;   // rejected by the host's data checks (core:1265-1285): nothing is written
    int n = 0;
    if (iopt != 0) {
        n = p.n[b];
        for (int i = lane; i < n; i += 32) ts[i] = p.t[(size_t)b * nest + i];
        if (iopt > 0)
            for (int i = lane; i < n; i += 32) {
                fpi[i] = p.fpint[(size_t)b * nest + i];
                nrd[i] = p.nrdata[(size_t)b * nest + i];
            }
    }
    __syncwarp();

    // scalars: every lane carries the same values (pure scalar arithmetic is simply done 32 times over)
    const double s = p.s, xb = p.ends_from_x ? xs[0] : p.xb, xe = p.ends_from_x ? xs[m - 1] : p.xe;
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
        for (int i = lane; i < nk1 * (K1 + 1); i += 32) as[i] = 0.0;
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
                hw[it * (K1 + 1) + j] = mul(wi, h[j]);
            }
            hw[it * (K1 + 1) + K1] = mul(ys[it], wi);
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
            const bool member = grp < G, isrhs = sub_ == K1;
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
                const int col = isrhs ? K1 : max(q, 0);
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
                if (fin && isrhs) res[row] = mul(hv, hv);
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
            fp = 0.0;                                          // summed in row order (core:4894)
            for (int it = 0; it < m; ++it) fp = add(fp, res[it]);
            __syncwarp();
        }
        if (ier == IER_LSQ) fp0 = fp;
        if (lane == 0) {
            FPI_(n - 1) = fpold;
            FPI_(n) = fp0;
            NRD_(n) = nplus;
        }
        back_subst_solo<K1>(as, cc, nk1, lane);
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
            double term = 0.0;
#pragma unroll
            for (int j = 1; j <= K1; ++j) term = add(term, mul(C_(l0 + j), qs[it * K1 + j - 1]));
            term = mul(ws[it], sub(term, ys[it]));
            res[it] = mul(term, term);
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
        pp = (double)nk1 / pp;
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
            for (int i = lane; i < nk1 * (K2 + 1); i += 32) {
                const int r = i / (K2 + 1), q = i - r * (K2 + 1);
                gs[i] = q < K1 ? as[r * (K1 + 1) + q] : (q == K2 ? as[r * (K1 + 1) + K1] : 0.0);
            }
            __syncwarp();
            // ---- smoothing rows b(it,:)/p into (G, c) as a wavefront (core:5052-5057, fp_rotate_shifted 5800-5820) ----
            {
                constexpr int LG = Cfg::LG_S, G = Cfg::G_S, NEVER = INT_MAX / 2;
                const int grp = lane / LG, sub_ = lane - grp * LG;
                const bool member = grp < G, isrhs = sub_ == K2;
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
                    const int tc = isrhs ? K2 : (mine && !pivlane ? dc : 0);
                    const double wwl = rowp[0], s2 = rowp[tc];
                    const double piv = __shfl_sync(FULL, hv, base + ps);
                    const bool rot = act && !fp_is_zero(piv);
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
            back_subst_solo<K2>(gs, cc, nk1, lane);
            // fp = sum of squared residuals in point order, core:5061-5069
            for (int it = lane; it < m; it += 32) {
                const int l0 = llag[it] - K2;
                double term = 0.0;
#pragma unroll
                for (int j = 1; j <= K1; ++j) term = add(term, mul(C_(l0 + j), qs[it * K1 + j - 1]));
                term = mul(ws[it], sub(term, ys[it]));
                res[it] = mul(term, term);
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
        for (int i = lane; i < nk1; i += 32) p.c[(size_t)b * nest + i] = cc[i];
    }
#undef T_
#undef C_
#undef FPI_
#undef NRD_
#undef A_
#undef B_
}

}  // namespace

size_t solo_smem_bytes(int m, int k, int nest)
{
    const size_t k1 = k + 1, k2 = k + 2;
    const size_t doubles = (size_t)m * (4 + k1 + k1 + 1) + (size_t)nest * (3 + (k1 + 1) + (k2 + 1) + k2);
    return doubles * 8 + ((size_t)2 * m + 2 * (size_t)nest) * 4;
}

cudaError_t launch_curfit_solo(const SoloParams &p, cudaStream_t st)
{
    const size_t smem = solo_smem_bytes(p.m, p.k, p.nest);
    cudaError_t err = cudaSuccess;
    int dev = 0;
    if ((err = cudaGetDevice(&dev)) != cudaSuccess) return err;
#define FPB_SOLO_CASE(KK)                                                                                          \
    case KK: {                                                                                                     \
        static size_t opted[64] = {0};   /* the attribute is per device */                                         \
        if (smem > opted[dev & 63]) {                                                                              \
            err = cudaFuncSetAttribute(curfit_solo_kernel<KK>, cudaFuncAttributeMaxDynamicSharedMemorySize,        \
                                       (int)smem);                                                                 \
            if (err != cudaSuccess) return err;                                                                    \
            opted[dev & 63] = smem;                                                                                \
        }                                                                                                          \
        curfit_solo_kernel<KK><<<p.nbatch, 32, smem, st>>>(p);                                                     \
        break;                                                                                                     \
    }
    switch (p.k) {
        FPB_SOLO_CASE(1)
        FPB_SOLO_CASE(2)
        FPB_SOLO_CASE(3)
        FPB_SOLO_CASE(4)
        FPB_SOLO_CASE(5)
    default: return cudaErrorInvalidValue;
    }
#undef FPB_SOLO_CASE
    return cudaGetLastError();
}

}  // namespace fpb
