/* Lockstep simulation of the warp-cooperative schedules of fpb_curfit_solo.cu (test infrastructure).
 *
 * The single-curve kernel runs the two band-QR sweeps of fpcurf (data rows: core:4873-4896 with fp_rotate_row
 * core:5691-5717; smoothing rows: core:5052-5057 with fp_rotate_shifted core:5800-5820) as a wavefront: G lane
 * groups each own one row at a time and all groups take one Givens step per tick, a row waiting until the row
 * before it has left the band row it needs.  This file replays exactly that schedule on the CPU -- every tick reads
 * the state the previous tick left, groups of one tick must touch distinct band rows (asserted) -- with the
 * oracle's own fpgivs / fprota as the arithmetic, so that the result can be compared bit for bit with the oracle's
 * sequential sweeps.  It proves the ORDER; the CUDA kernel is tested against the oracle on the GPU.
 */
#include "../../oracle/fitpack_oracle.c"
#include <assert.h>

#define SIM_MAXG 8

typedef struct {
    int row;      /* current row (>= nrows: exhausted) */
    int j;        /* next band row (1-based) */
    int last;     /* last band row of the current row */
    int first;    /* first band row of the current row */
    double h[8];  /* the row's entries; LSQ: h[c] <-> band column first+c; smoothing: by absolute column, see below */
    int col[8];   /* smoothing: absolute column each slot holds */
    double yi;
} sim_group;

/* may group g take its step this tick?  (state of the tick's beginning) */
static int sim_can_go(const sim_group *gr, int G, int g, int nrows)
{
    int r = gr[g].row, pg = (g + G - 1) % G, pr;
    if (r >= nrows) return 0;
    if (r == 0) return 1;
    pr = gr[pg].row;
    if (pr > r - 1) return 1;
    if (pr < r - 1) return 0;
    return gr[pg].j > gr[g].j;
}

/* data rows of one least-squares pass; lidx[r] = knot interval of point r (1-based l), q = basis values [m][k1] */
int32_t sim_lsq_sweep(int32_t m, int32_t k, int32_t nk1, int32_t nest, const int32_t *lidx, const double *q,
                      const double *y, const double *w, int32_t G, double *a, double *z, double *fp_out,
                      int32_t *ticks_out)
{
    int k1 = k + 1, g, c, i, ticks = 0;
    sim_group gr[SIM_MAXG];
    double fp = 0.0;
    for (i = 0; i < nk1; ++i) z[i] = 0.0;
    for (c = 0; c < k1; ++c)
        for (i = 0; i < nk1; ++i) a[c * nest + i] = 0.0;
    for (g = 0; g < G; ++g) {
        gr[g].row = g;
        if (g < m) {
            for (c = 0; c < k1; ++c) gr[g].h[c] = w[g] * q[g * k1 + c];
            gr[g].yi = y[g] * w[g];
            gr[g].first = lidx[g] - k;
            gr[g].last = lidx[g];
            gr[g].j = gr[g].first;
        }
    }
    for (;;) {
        int go[SIM_MAXG], any = 0, used[SIM_MAXG], nu = 0, done_rows = 0;
        sim_group snap[SIM_MAXG];
        for (g = 0; g < G; ++g) snap[g] = gr[g];
        for (g = 0; g < G; ++g) {
            go[g] = sim_can_go(snap, G, g, m);
            if (snap[g].row < m) any = 1;
        }
        if (!any) break;
        ++ticks;
        for (g = 0; g < G; ++g) {
            int j, i0, qq;
            double piv, cs, sn;
            if (!go[g]) continue;
            j = gr[g].j;
            for (i = 0; i < nu; ++i) assert(used[i] != j);
            used[nu++] = j;
            i0 = j - gr[g].first;
            piv = gr[g].h[i0];
            if (!orc_equal(piv, 0.0)) {
                fpgivs(piv, &a[j - 1], &cs, &sn);
                fprota(cs, sn, &gr[g].yi, &z[j - 1]);
                for (qq = 1; qq <= k - i0; ++qq) fprota(cs, sn, &gr[g].h[i0 + qq], &a[qq * nest + (j - 1)]);
            }
            if (j == gr[g].last) {
                int r2 = gr[g].row + G;
                assert(done_rows == 0); /* one row at most completes per tick: fp is summed in row order */
                ++done_rows;
                fp = fp + gr[g].yi * gr[g].yi;
                gr[g].row = r2;
                if (r2 < m) {
                    for (c = 0; c < k1; ++c) gr[g].h[c] = w[r2] * q[r2 * k1 + c];
                    gr[g].yi = y[r2] * w[r2];
                    gr[g].first = lidx[r2] - k;
                    gr[g].last = lidx[r2];
                    gr[g].j = gr[g].first;
                }
            } else {
                gr[g].j = j + 1;
            }
        }
    }
    *fp_out = fp;
    *ticks_out = ticks;
    return 0;
}

/* smoothing rows of one p-iteration: g (band k2) and c start as copies of a and z; rows it = 1..n8 of b/p */
int32_t sim_smooth_sweep(int32_t n, int32_t k, int32_t nest, const double *b, double pinv, int32_t G, double *gm,
                         double *cz, int32_t *ticks_out)
{
    int k1 = k + 1, k2 = k + 2, nk1 = n - k1, n8 = n - 2 * k1, g, c, i, ticks = 0;
    sim_group gr[SIM_MAXG];
    for (g = 0; g < G; ++g) {
        gr[g].row = g;
        if (g < n8) {
            for (c = 0; c < k2; ++c) {
                gr[g].h[c] = b[c * nest + g] * pinv;
                gr[g].col[c] = g + 1 + c;
            }
            gr[g].yi = 0.0;
            gr[g].first = g + 1;
            gr[g].last = nk1;
            gr[g].j = g + 1;
        }
    }
    for (;;) {
        int go[SIM_MAXG], any = 0, used[SIM_MAXG], nu = 0;
        sim_group snap[SIM_MAXG];
        for (g = 0; g < G; ++g) snap[g] = gr[g];
        for (g = 0; g < G; ++g) {
            go[g] = sim_can_go(snap, G, g, n8);
            if (snap[g].row < n8) any = 1;
        }
        if (!any) break;
        ++ticks;
        for (g = 0; g < G; ++g) {
            int j, ps, noff;
            double piv, cs, sn;
            if (!go[g]) continue;
            j = gr[g].j;
            for (i = 0; i < nu; ++i) assert(used[i] != j);
            used[nu++] = j;
            ps = (j - gr[g].first) % k2;          /* slot that holds column j */
            assert(gr[g].col[ps] == j);
            piv = gr[g].h[ps];
            noff = nk1 - j;
            if (noff > k2 - 1) noff = k2 - 1;
            if (!orc_equal(piv, 0.0)) {
                fpgivs(piv, &gm[j - 1], &cs, &sn);
                fprota(cs, sn, &gr[g].yi, &cz[j - 1]);
                for (c = 0; c < k2; ++c) {
                    int d = gr[g].col[c] - j;
                    if (d >= 1 && d <= noff) fprota(cs, sn, &gr[g].h[c], &gm[d * nest + (j - 1)]);
                }
            }
            /* the shift: the pivot's slot takes the column that enters the band, value 0 */
            gr[g].h[ps] = 0.0;
            gr[g].col[ps] = j + k2;
            if (j == gr[g].last) {
                int r2 = gr[g].row + G;
                gr[g].row = r2;
                if (r2 < n8) {
                    for (c = 0; c < k2; ++c) {
                        gr[g].h[c] = b[c * nest + r2] * pinv;
                        gr[g].col[c] = r2 + 1 + c;
                    }
                    gr[g].yi = 0.0;
                    gr[g].first = r2 + 1;
                    gr[g].last = nk1;
                    gr[g].j = r2 + 1;
                }
            } else {
                gr[g].j = j + 1;
            }
        }
    }
    *ticks_out = ticks;
    return 0;
}


/* ---- the static schedules the kernel uses (fpb_curfit_solo.cu, SoloCfg) ----
 * data rows: row r meets band row j at tick r * ST + j; smoothing rows: start ticks from the table sst[]. */
int32_t sim_lsq_sweep_static(int32_t m, int32_t k, int32_t nk1, int32_t nest, const int32_t *lidx, const double *q,
                             const double *y, const double *w, int32_t G, int32_t ST, double *a, double *z,
                             double *fp_out, int32_t *ticks_out)
{
    int k1 = k + 1, g, c, i, tau, ticks = 0;
    int row[SIM_MAXG], first[SIM_MAXG], start[SIM_MAXG];
    double h[SIM_MAXG][8], yi[SIM_MAXG], *res = (double *)malloc(sizeof(double) * m), fp = 0.0;
    for (i = 0; i < nk1; ++i) z[i] = 0.0;
    for (c = 0; c < k1; ++c)
        for (i = 0; i < nk1; ++i) a[c * nest + i] = 0.0;
    for (g = 0; g < G; ++g) {
        row[g] = g;
        if (g < m) {
            first[g] = lidx[g] - k;
            start[g] = g * ST + first[g];
            for (c = 0; c < k1; ++c) h[g][c] = w[g] * q[g * k1 + c];
            yi[g] = y[g] * w[g];
        }
    }
    for (tau = lidx[0] - k; tau <= (m - 1) * ST + lidx[m - 1]; ++tau) {
        int used[SIM_MAXG], nu = 0;
        ++ticks;
        for (g = 0; g < G; ++g) {
            int i0, j, qq;
            double piv, cs, sn;
            if (row[g] >= m) continue;
            i0 = tau - start[g];
            if (i0 < 0) continue;
            assert(i0 <= k);
            j = first[g] + i0;
            for (i = 0; i < nu; ++i) assert(used[i] != j);
            used[nu++] = j;
            piv = h[g][i0];
            if (!orc_equal(piv, 0.0)) {
                fpgivs(piv, &a[j - 1], &cs, &sn);
                fprota(cs, sn, &yi[g], &z[j - 1]);
                for (qq = 1; qq <= k - i0; ++qq) fprota(cs, sn, &h[g][i0 + qq], &a[qq * nest + (j - 1)]);
            }
            if (i0 == k) {
                int r2 = row[g] + G;
                res[row[g]] = yi[g] * yi[g];
                row[g] = r2;
                if (r2 < m) {
                    first[g] = lidx[r2] - k;
                    start[g] = r2 * ST + first[g];
                    assert(start[g] > tau);   /* the group is free before its next row starts */
                    for (c = 0; c < k1; ++c) h[g][c] = w[r2] * q[r2 * k1 + c];
                    yi[g] = y[r2] * w[r2];
                }
            }
        }
    }
    for (g = 0; g < G; ++g) assert(row[g] >= m);
    for (i = 0; i < m; ++i) fp = fp + res[i];
    free(res);
    *fp_out = fp;
    *ticks_out = ticks;
    return 0;
}

int32_t sim_smooth_sweep_static(int32_t n, int32_t k, int32_t nest, const double *b, double pinv, int32_t G,
                                double *gm, double *cz, int32_t *ticks_out)
{
    int k1 = k + 1, k2 = k + 2, nk1 = n - k1, n8 = n - 2 * k1, g, c, i, r, tau, ticks = 0;
    int row[SIM_MAXG], start[SIM_MAXG], col[SIM_MAXG][8];
    double h[SIM_MAXG][8], yi[SIM_MAXG];
    int *sst = (int *)malloc(sizeof(int) * (n8 > 0 ? n8 : 1));
    for (r = 0; r < n8; ++r) {
        int st = r == 0 ? 0 : sst[r - 1] + 2;
        if (r >= G && sst[r - G] + nk1 - (r - G) > st) st = sst[r - G] + nk1 - (r - G);
        sst[r] = st;
    }
    for (g = 0; g < G; ++g) {
        row[g] = g;
        if (g < n8) {
            start[g] = sst[g];
            for (c = 0; c < k2; ++c) { h[g][c] = b[c * nest + g] * pinv; col[g][c] = g + 1 + c; }
            yi[g] = 0.0;
        }
    }
    for (tau = 0; n8 > 0 && tau <= sst[n8 - 1] + nk1 - n8; ++tau) {
        int used[SIM_MAXG], nu = 0;
        ++ticks;
        for (g = 0; g < G; ++g) {
            int d0, j, ps, noff;
            double piv, cs, sn;
            if (row[g] >= n8) continue;
            d0 = tau - start[g];
            if (d0 < 0) continue;
            j = row[g] + 1 + d0;
            assert(j <= nk1);
            for (i = 0; i < nu; ++i) assert(used[i] != j);
            used[nu++] = j;
            ps = d0 % k2;
            assert(col[g][ps] == j);
            piv = h[g][ps];
            noff = nk1 - j;
            if (noff > k2 - 1) noff = k2 - 1;
            if (!orc_equal(piv, 0.0)) {
                fpgivs(piv, &gm[j - 1], &cs, &sn);
                fprota(cs, sn, &yi[g], &cz[j - 1]);
                for (c = 0; c < k2; ++c) {
                    int d = col[g][c] - j;
                    if (d >= 1 && d <= noff) fprota(cs, sn, &h[g][c], &gm[d * nest + (j - 1)]);
                }
            }
            h[g][ps] = 0.0;
            col[g][ps] = j + k2;
            if (j == nk1) {
                int r2 = row[g] + G;
                row[g] = r2;
                if (r2 < n8) {
                    start[g] = sst[r2];
                    assert(start[g] > tau);
                    for (c = 0; c < k2; ++c) { h[g][c] = b[c * nest + r2] * pinv; col[g][c] = r2 + 1 + c; }
                    yi[g] = 0.0;
                }
            }
        }
    }
    for (g = 0; g < G; ++g) assert(row[g] >= n8);
    free(sst);
    *ticks_out = ticks;
    return 0;
}

/* the oracle's sequential sweeps on the same inputs */
void sim_ref_lsq_sweep(int32_t m, int32_t k, int32_t nk1, int32_t nest, const int32_t *lidx, const double *q,
                       const double *y, const double *w, double *a, double *z, double *fp_out)
{
    int k1 = k + 1, c, i, it;
    double fp = 0.0, h[8];
    for (i = 0; i < nk1; ++i) z[i] = 0.0;
    for (c = 0; c < k1; ++c)
        for (i = 0; i < nk1; ++i) a[c * nest + i] = 0.0;
    for (it = 0; it < m; ++it) {
        double yi = y[it] * w[it];
        for (c = 0; c < k1; ++c) h[c] = w[it] * q[it * k1 + c];
        rotate_row(h, k1, a, nest, &yi, z, lidx[it] - k1);
        fp = fp + yi * yi;
    }
    *fp_out = fp;
}

void sim_ref_smooth_sweep(int32_t n, int32_t k, int32_t nest, const double *b, double pinv, double *gm, double *cz)
{
    int k1 = k + 1, k2 = k + 2, nk1 = n - k1, n8 = n - 2 * k1, it, j;
    double h[8];
    for (it = 1; it <= n8; ++it) {
        double yi = 0.0;
        for (j = 0; j < k2; ++j) h[j] = b[j * nest + (it - 1)] * pinv;
        rotate_shifted(h, k2, gm, nest, &yi, cz, it, nk1);
    }
}

void sim_fpdisc(const double *t, int32_t n, int32_t k2, double *b, int32_t nest) { fpdisc(t, n, k2, b, nest); }
