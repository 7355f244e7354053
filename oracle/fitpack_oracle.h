/*
 * fitpack_oracle.h -- CPU oracle for the curfit/fpcurf + splev/fpbspl hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (fitpack_b200/, include/)
 * may include, link or call this.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py use it, as the checker / CPU
 * baseline.
 *
 * It is a plain-C restatement (not a copy) of the algorithms in the reference's
 * src/fitpack_core.F90; every function cites the reference lines it follows.
 *
 * PARITY PIN STATUS: the reference is 100 % Fortran and this image has no
 * Fortran compiler, so the reference itself cannot be run here ("oracle/_ref"
 * is unbuildable).  The oracle is pinned instead against (i) every assertion
 * the reference's own tests make for this path (mncurf, mnspev,
 * test_fpknot_crash, analytic sine fit), and (ii) golden vectors generated
 * from scipy.interpolate._fitpack (C translation of the same Dierckx
 * algorithms) on inputs where the reference does not deliberately deviate from
 * netlib -- see tests/golden/make_golden.py -- and (iii) directed cases for
 * every documented reference-vs-netlib deviation with results derived by hand
 * from the reference source (tests/deviation_cases.py).  Bit-level agreement
 * with the Fortran build itself is therefore NOT pinned ("parity unpinned"
 * w.r.t. the gfortran binary; pinned w.r.t. the second source).  The recipe
 * that closes this on a box with gfortran is oracle/ref_recipe/ (build_ref.sh
 * builds the reference core + its C binding from /root/reference where they
 * lie, make_ref_golden.py writes tests/golden/ref_*.npz, tests/test_ref_golden.py
 * and tests/test_gpu_ref_golden.py consume them).
 */
#ifndef FITPACK_ORACLE_H
#define FITPACK_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* per-fit operation counters (used for the algorithmic flop figure) */
typedef struct orc_stats {
    int32_t lsq_passes;   /* number of knot sets tried (part 1 passes)        */
    int32_t p_iters;      /* number of smoothing-parameter iterations (part 2) */
    int64_t rotations;    /* number of Givens rotations generated (fpgivs)     */
} orc_stats;

/* curfit, core:1240-1299 (same argument list as fp_curfit_c) */
void orc_curfit(int32_t iopt, int32_t m, const double *x, const double *y, const double *w,
                double xb, double xe, int32_t k, double s, int32_t nest,
                int32_t *n, double *t, double *c, double *fp,
                double *wrk, int32_t lwrk, int32_t *iwrk, int32_t *ier);

/* same, also returning the counters of this call */
void orc_curfit_stats(int32_t iopt, int32_t m, const double *x, const double *y, const double *w,
                      double xb, double xe, int32_t k, double s, int32_t nest,
                      int32_t *n, double *t, double *c, double *fp,
                      double *wrk, int32_t lwrk, int32_t *iwrk, int32_t *ier, orc_stats *st);

/* splev, core:17826-17896 (same argument list as fp_splev_c) */
void orc_splev(const double *t, int32_t n, const double *c, int32_t k,
               const double *x, double *y, int32_t m, int32_t e, int32_t *ier);

/* splev that also reports the knot interval index l (1-based) used per point */
void orc_splev_idx(const double *t, int32_t n, const double *c, int32_t k,
                   const double *x, double *y, int32_t *lidx, int32_t m, int32_t e, int32_t *ier);

/* fpchec, core:2507-2570 */
int32_t orc_fpchec(const double *x, int32_t m, const double *t, int32_t n, int32_t k);

/* fpbspl, core:2387-2413 : h[0..k] = the k+1 non-zero B-splines at x, l 1-based */
void orc_fpbspl(const double *t, int32_t n, int32_t k, double x, int32_t l, double *h);

/* fp_knot_interval, core:2432-2473 (1-based l) */
int32_t orc_knot_interval(const double *t, double x, int32_t l_start, int32_t l_max);

/* equal(), core:18644-18647 */
int32_t orc_equal(double a, double b);

/* FITPACK_SUCCESS core:393-396 ; FITPACK_MESSAGE core:315-339 */
int32_t orc_success(int32_t ier);
const char *orc_message(int32_t ier);

/* fitpack_argsort/quicksort, core:18486-18590 (1-based permutation out) */
void orc_argsort(const double *list, int32_t n, int32_t *ilist);

/*
 * Batch drivers (OpenMP over curves; curve-major arrays).  These are the CPU
 * baseline of bench.py: each curve is an independent orc_curfit call with a
 * per-thread workspace, exactly how one would drive the (pure) reference.
 *   x,y,w : [nbatch][m] (w may be NULL => weights 1; shared_x!=0 => x,w are [m])
 *   t,c   : [nbatch][nest]; n,ier : [nbatch]; fp : [nbatch]
 *   s     : scalar smoothing factor for all curves
 *   stats : NULL or [nbatch]
 */
void orc_curfit_batch(int32_t nbatch, int32_t iopt, int32_t m, const double *x, const double *y,
                      const double *w, int32_t shared_x, double xb, double xe, int32_t use_data_range,
                      int32_t k, double s, int32_t nest, int32_t *n, double *t, double *c,
                      double *fp, int32_t *ier, orc_stats *stats, int32_t nthreads);

/* splev over many splines: t,c [nspl][ldt]; n [nspl]; x,y [nspl][m] */
void orc_splev_batch(int32_t nspl, const double *t, const int32_t *n, const double *c, int32_t ldt,
                     int32_t k, const double *x, double *y, int32_t m, int32_t e, int32_t *ier,
                     int32_t nthreads);

/* regrid (dims = 2), core:16732-16839 + fpregr 12079-12303 + fpgrre 7111-7360 (argument list of
   fp_regrid_c, src/fitpack_core_c.f90:253-290).  z, c flat row-major with x slowest.
   State for iopt=1: wrk[0]=fp0, wrk[1]=fpold, wrk[2..3]=reduc, iwrk[0]=lastdi, iwrk[1..2]=nplus. */
void orc_regrid(int32_t iopt, int32_t mx, const double *x, int32_t my, const double *y,
                const double *z, double xb, double xe, double yb, double ye, int32_t kx, int32_t ky,
                double s, int32_t nxest, int32_t nyest, int32_t *nx, double *tx, int32_t *ny,
                double *ty, double *c, double *fp, double *wrk, int32_t lwrk, int32_t *iwrk,
                int32_t kwrk, int32_t *ier);

/* bispev, core:424-458 + fpndsp 2160-2241 at dims = 2 (argument list of fp_bispev_c) */
int32_t orc_bispev(const double *tx, int32_t nx, const double *ty, int32_t ny, const double *c,
                   int32_t kx, int32_t ky, const double *x, int32_t mx, const double *y, int32_t my,
                   double *z, double *wrk, int32_t lwrk, int32_t *iwrk, int32_t kwrk);

/* splder core:17699-17802, spalde core:16865-16895 (+fpader 1546-1603), splint core:17919-17939
   (+fpintb 8058-8166): argument lists of fp_splder_c / fp_spalde_c / fp_splint_c */
void orc_splder(const double *t, int32_t n, const double *c, int32_t k, int32_t nu, const double *x,
                double *y, int32_t m, int32_t e, double *wrk, int32_t *ier);
void orc_spalde(const double *t, int32_t n, const double *c, int32_t k1, double x, double *d,
                int32_t *ier);
double orc_splint(const double *t, int32_t n, const double *c, int32_t k, double a, double b,
                  double *wrk);
/* fourco core:1474-1520 with fpbfou 1947-2123, fpcsin 4636-4686: Fourier integrals of a cubic spline (fp_fourco_c) */
/* cualde core:1062-1101 (fp_cualde_c), insert core:15064-15159 + fpinst 7961-8026 (fp_insert_c, in place) */
void orc_cualde(int32_t idim, const double *t, int32_t n, const double *c, int32_t nc, int32_t k1, double u,
                double *d, int32_t nd, int32_t *ier);
void orc_insert(int32_t iopt, double *t, int32_t *n_io, double *c, int32_t k, double x, int32_t nest,
                int32_t *ier);
/* sproot core:17962-18109 with fpcuro 5108-5198 (fp_sproot_c): the zeros of a cubic spline in [t(4), t(n-3)] */
void orc_sproot(const double *t, int32_t n, const double *c, double *zeros, int32_t mest, int32_t *m_out,
                int32_t *ier);
void orc_fourco(const double *t, int32_t n, const double *c, const double *alfa, int32_t m, double *ress,
                double *resc, double *wrk1, double *wrk2, int32_t *ier);

/* parcur core:15201-15285 + fppara 8822-9177 (argument list of fp_parcur_c), curev core:1126-1182
   (argument list of fp_curev_c).  x(idim,m) coordinate-fastest; c((d-1)*n + i). */
void orc_parcur(int32_t iopt, int32_t ipar, int32_t idim, int32_t m, double *u, int32_t mx,
                const double *x, const double *w, double *ub, double *ue, int32_t k, double s,
                int32_t nest, int32_t *n, double *t, int32_t nc, double *c, double *fp, double *wrk,
                int32_t lwrk, int32_t *iwrk, int32_t *ier);
void orc_curev(int32_t idim, const double *t, int32_t n, const double *c, int32_t nc, int32_t k,
               const double *u, int32_t m, double *x, int32_t mx, int32_t *ier);
void orc_parcur_batch(int32_t nbatch, int32_t iopt, int32_t ipar, int32_t idim, int32_t m, double *u,
                      const double *x, const double *w, double *ub, double *ue, int32_t k, double s,
                      int32_t nest, int32_t *n, double *t, double *c, double *fp, int32_t *ier,
                      int32_t nthreads);

/* percur core:15784-15863 + fpperi 9668-10192 (argument list of fp_percur_c); fpchep core:2689-2770 */
void orc_percur(int32_t iopt, int32_t m, const double *x, const double *y, const double *w, int32_t k,
                double s, int32_t nest, int32_t *n, double *t, double *c, double *fp, double *wrk,
                int32_t lwrk, int32_t *iwrk, int32_t *ier);
int32_t orc_fpchep(const double *x, int32_t m, const double *t, int32_t n, int32_t k);
void orc_percur_batch(int32_t nbatch, int32_t iopt, int32_t m, const double *x, const double *y,
                      const double *w, int32_t k, double s, int32_t nest, int32_t *n, double *t, double *c,
                      double *fp, int32_t *ier, int32_t nthreads);

void orc_curev_batch(int32_t nspl, int32_t idim, const double *t, const int32_t *n, const double *c,
                     int32_t ldt, int64_t ldc, int32_t k, const double *u, double *x, int32_t m,
                     int32_t *ier, int32_t nthreads);

int32_t orc_max_threads(void);

/* tests only: switch the documented reference-vs-netlib deviations off (see .c) */
void orc_set_netlib_compat(int32_t on);

#ifdef __cplusplus
}
#endif
#endif
