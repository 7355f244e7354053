/*
 * fitpack_b200.h -- C-ABI of the B200-native batched spline engine.
 *
 * Drop-in boundary for the curfit/fpcurf + splev/fpbspl hot path of
 * perazz/fitpack.  Three groups of entry points:
 *
 *  (A) the reference's own symbols for this path, same names, same argument
 *      lists, same error behaviour -- a caller linked against the reference's
 *      libfitpack can link against libfitpack_b200 instead:
 *        fp_curfit_c, fp_splev_c, FITPACK_SUCCESS_c, fitpack_message_c
 *          (reference: include/fitpack_core_c.h:69,84,98-101,131-132;
 *           src/fitpack_core_c.f90:44-74,145-154)
 *        fitpack_curve_c_* opaque-handle API
 *          (reference: include/fitpack_curves_c.h:37-63;
 *           src/fitpack_curves_c.f90:56-396)
 *      These run the SAME CUDA kernels as a batch of one curve.
 *
 *  (B) NEW batch entry points (the reference has none): many independent
 *      curves per call, host-pointer and device-pointer variants, optional
 *      sharding over the GPUs of one node.
 *
 *  (C) engine lifetime / introspection.
 *
 * All reals are IEEE binary64, all integers int32 (FP_REAL/FP_SIZE/FP_FLAG of
 * the reference, src/fitpack_core.F90:26-29).  There is NO CPU fallback: every
 * compute entry point needs a CUDA device and reports FPB_ERR_CUDA otherwise.
 */
#ifndef FITPACK_B200_H_INCLUDED
#define FITPACK_B200_H_INCLUDED

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef double  FP_REAL;
typedef int32_t FP_SIZE;
typedef int32_t FP_FLAG;
typedef bool    FP_BOOL;

/* out-of-support policy of splev (reference core:101-105) */
#define FPB_OUTSIDE_EXTRAPOLATE 0
#define FPB_OUTSIDE_ZERO        1
#define FPB_OUTSIDE_NOT_ALLOWED 2
#define FPB_OUTSIDE_NEAREST_BND 3

/* iopt (reference core:107-111) */
#define FPB_IOPT_NEW_LEASTSQUARES (-1)
#define FPB_IOPT_NEW_SMOOTHING      0
#define FPB_IOPT_OLD_FIT            1

/* ier codes (reference core:129-144) */
#define FPB_OK                     0
#define FPB_INTERPOLATING_OK     (-1)
#define FPB_LEASTSQUARES_OK      (-2)
#define FPB_INSUFFICIENT_STORAGE   1
#define FPB_S_TOO_SMALL            2
#define FPB_MAXIT                  3
#define FPB_INVALID_RANGE          6
#define FPB_INPUT_ERROR           10
#define FPB_INSUFFICIENT_KNOTS    13   /* fourco: a cubic spline with n >= 10 knots is required */

/* status of the batch/engine calls themselves (not per-curve ier) */
#define FPB_STATUS_OK      0
#define FPB_ERR_CUDA     (-100)  /* no device / CUDA runtime error (see fpb_last_error) */
#define FPB_ERR_ARG      (-101)  /* invalid call-level argument (k, m, nest, strides...) */
#define FPB_ERR_ALLOC    (-102)  /* device or host allocation failed */

/* splev evaluation mode for the batch entry points */
#define FPB_SPLEV_EXACT 0  /* de Boor-Cox exactly as the reference: bit-identical values    */
#define FPB_SPLEV_FAST  1  /* per-interval local polynomial + Horner/FMA: HBM-bound;
                              identical interval indices, values within 1e-12*max|c|        */

/* ------------------------------------------------------------------ (A) -- */
/* reference: src/fitpack_core_c.f90:44-47 */
FP_BOOL FITPACK_SUCCESS_c(FP_FLAG ierr);
/* reference: src/fitpack_core_c.f90:50-61 (NUL-terminated into caller buffer) */
void fitpack_message_c(FP_FLAG ierr, char *msg);

/* reference: src/fitpack_core_c.f90:64-74 -> curfit core:1240-1299.
   Host pointers.  wrk/iwrk keep the reference's meaning: on return
   wrk[0..nest) holds fpint and iwrk[0..nest) holds nrdata, which is the state an
   iopt=1 continuation reads back (core:4825-4834, 4898-4900). */
void fp_curfit_c(FP_SIZE iopt, FP_SIZE m, const FP_REAL *x, const FP_REAL *y, const FP_REAL *w,
                 FP_REAL xb, FP_REAL xe, FP_SIZE k, FP_REAL s, FP_SIZE nest,
                 FP_SIZE *n, FP_REAL *t, FP_REAL *c, FP_REAL *fp,
                 FP_REAL *wrk, FP_SIZE lwrk, FP_SIZE *iwrk, FP_FLAG *ier);

/* reference: src/fitpack_core_c.f90:145-154 -> splev core:17826-17896 */
void fp_splev_c(const FP_REAL *t, FP_SIZE n, const FP_REAL *c, FP_SIZE k, const FP_REAL *x,
                FP_REAL *y, FP_SIZE m, FP_FLAG e, FP_FLAG *ier);

/* Opaque-handle curve API, reference include/fitpack_curves_c.h:37-63.
   Optional arguments are nullable pointers exactly as in the reference
   (src/fitpack_curves_c.f90:159-163,176-182,195-199). */
typedef struct fitpack_curve_c {
    void *cptr;
} fitpack_curve_c;

void    fitpack_curve_c_allocate     (fitpack_curve_c *self);
void    fitpack_curve_c_destroy      (fitpack_curve_c *self);
void    fitpack_curve_c_copy         (fitpack_curve_c *self, const fitpack_curve_c *other);
void    fitpack_curve_c_move_alloc   (fitpack_curve_c *to, fitpack_curve_c *from);
void    fitpack_curve_c_new_points   (fitpack_curve_c *self, FP_SIZE npts, FP_REAL *x, FP_REAL *y, FP_REAL *w);
FP_FLAG fitpack_curve_c_new_fit      (fitpack_curve_c *self, FP_SIZE npts, FP_REAL *x, FP_REAL *y, FP_REAL *w, FP_REAL *smoothing);
FP_FLAG fitpack_curve_c_fit          (fitpack_curve_c *self, FP_REAL *smoothing, FP_SIZE *order);
FP_FLAG fitpack_curve_c_interpolating(fitpack_curve_c *self, FP_SIZE *order);
FP_REAL fitpack_curve_c_eval_one     (fitpack_curve_c *self, FP_REAL x, FP_FLAG *ierr);
void    fitpack_curve_c_eval_many    (fitpack_curve_c *self, FP_SIZE npts, FP_REAL *x, FP_REAL *y, FP_FLAG *ierr);
FP_REAL fitpack_curve_c_integral     (fitpack_curve_c *self, FP_REAL from, FP_REAL to);
void    fitpack_curve_c_fourier      (fitpack_curve_c *self, FP_SIZE nparm, const FP_REAL *alpha, FP_REAL *A, FP_REAL *B,
                                      FP_FLAG *ierr);
FP_REAL fitpack_curve_c_derivative   (fitpack_curve_c *self, FP_REAL x, FP_SIZE order, FP_SIZE *ierr);
FP_FLAG fitpack_curve_c_all_derivatives(fitpack_curve_c *self, FP_REAL x, FP_REAL *ddx);
FP_REAL fitpack_curve_c_smoothing    (fitpack_curve_c *self);
FP_REAL fitpack_curve_c_mse          (fitpack_curve_c *self);
FP_SIZE fitpack_curve_c_degree       (fitpack_curve_c *self);
void    fitpack_curve_c_set_bc       (fitpack_curve_c *self, FP_FLAG bc);
FP_FLAG fitpack_curve_c_get_bc       (fitpack_curve_c *self);
/* extras the Fortran type exposes as public components (src/fitpack_curves.f90:54-78);
   needed by host code that reads knots/coefficients back */
FP_FLAG fitpack_curve_c_least_squares(fitpack_curve_c *self, FP_REAL *smoothing, FP_BOOL *reset_knots);
FP_SIZE fitpack_curve_c_knots        (fitpack_curve_c *self);
FP_SIZE fitpack_curve_c_get_t        (fitpack_curve_c *self, FP_SIZE nmax, FP_REAL *t);
FP_SIZE fitpack_curve_c_get_c        (fitpack_curve_c *self, FP_SIZE nmax, FP_REAL *c);
FP_SIZE fitpack_curve_c_set_knots    (fitpack_curve_c *self, FP_SIZE n, const FP_REAL *t);
FP_FLAG fitpack_curve_c_iopt         (fitpack_curve_c *self);
/* comm_size / comm_pack / comm_expand of type(fitpack_curve) (src/fitpack_curves.f90:821-895,
   src/fitpack_fitters.f90:105-147, wire helpers core:18654-18956): a fitted curve as a flat array of
   FP_REAL, same wire format as the reference (scalars as doubles; arrays = bounds header packed as
   two int32 per dimension, -99999999 when not allocated, + raw payload), so a buffer produced by
   either side expands on the other -- the shipping format for fitted batches between ranks. */
FP_SIZE fitpack_curve_c_comm_size    (fitpack_curve_c *self);
void    fitpack_curve_c_comm_pack    (fitpack_curve_c *self, FP_REAL *buffer);
void    fitpack_curve_c_comm_expand  (fitpack_curve_c *self, const FP_REAL *buffer);

/* ------------------------------------------------------------------ (C) -- */
typedef struct fpb_engine fpb_engine;   /* one per GPU: stream + workspace   */

/* create an engine on CUDA device `device` (-1: current device). */
int32_t fpb_engine_create(int32_t device, fpb_engine **out);
void    fpb_engine_destroy(fpb_engine *eng);
/* make the engine launch on an externally owned stream (cudaStream_t, e.g. torch's current
   stream).  The handle is taken literally: NULL is CUDA's default stream.
   fpb_engine_reset_stream goes back to the engine's own non-blocking stream. */
int32_t fpb_engine_set_stream(fpb_engine *eng, void *cuda_stream);
int32_t fpb_engine_reset_stream(fpb_engine *eng);
int32_t fpb_engine_synchronize(fpb_engine *eng);
/* arithmetic of the shared-abscissa least-squares path (fp_curfit_shared_dev_c):
   FPB_SHARED_EXACT (default): the reference's rotations, 4 mul + 2 add, bit-identical to
   curfit(iopt=-1) per curve;  FPB_SHARED_FAST: fused multiply-add rotations (2 mul + 2 fma),
   c and fp within rounding (tests: 1e-12 relative to max|c|, max(1,fp)). */
#define FPB_SHARED_EXACT 0
#define FPB_SHARED_FAST  1
int32_t fpb_engine_set_shared_mode(fpb_engine *eng, int32_t mode);
/* arithmetic of fp_curev_batch_dev_c: FPB_CUREV_EXACT (default): the reference's de Boor-Cox recurrence and
   operation order, bit-identical values;  FPB_CUREV_FAST: one local polynomial per knot interval +
   Horner/FMA through the fast splev kernel, one pass per coordinate (identical knot intervals, values within
   1e-12 * max(1, max|c|) of exact; ier as in exact mode). */
/* arithmetic of fp_bispev_dev_c: FPB_BISPEV_EXACT (default): the reference's sums in the reference's order,
   no FMA, bit-identical values;  FPB_BISPEV_FAST: the same sums as fused multiply-adds (values within
   1e-12 * max(1, max|c|) of exact; same interval indices, same ier). */
#define FPB_BISPEV_EXACT 0
#define FPB_BISPEV_FAST  1
int32_t fpb_engine_set_bispev_mode(fpb_engine *eng, int32_t mode);
#define FPB_CUREV_EXACT 0
#define FPB_CUREV_FAST  1
int32_t fpb_engine_set_curev_mode(fpb_engine *eng, int32_t mode);
/* Which kernels a batch of ONE curve from host memory runs on (fp_curfit_c, the fitpack_curve_c_* handle functions,
   fp_curfit_batch_c with nbatch = 1): 1 (default) the warp-per-curve kernel -- one warp shares the curve, the two
   band-QR sweeps of fpcurf run as a wavefront over lane groups (fitpack_b200/csrc/fpb_curfit_solo.cu), ~10x lower
   latency; 0 the batch kernels with one thread per curve.  Same bits either way.  Process-wide; returns the
   previous setting (the environment variable FPB_SOLO=0 sets the initial one). */
int32_t fpb_set_single_curve_kernel(int32_t on);
/* How many ranks / slices this process assumes are copying through the same host (LOCAL_WORLD_SIZE,
   OMPI_COMM_WORLD_LOCAL_SIZE, SLURM_NTASKS_PER_NODE or FPB_LOCAL_RANKS; at least 1), and which form of the
   host-pointer batch fit a fresh fit (iopt = 0, own abscissae) of `nbatch` curves per GPU slice takes with it:
   "warp-per-curve" (small batches), "wavefront" (one pipeline the chunks join as they land) or "chunks" (chunk per
   engine; the choice when more than 2 sharers make the host the bound).  DESIGN.md section 5.  No GPU needed. */
int32_t fpb_host_sharers(void);
const char *fpb_host_form(int64_t nbatch);
/* number of kernels this engine has launched since creation (bench bookkeeping) */
int64_t fpb_engine_launch_count(const fpb_engine *eng);
/* same counter for the DEFAULT engine of `device` -- the one the host-pointer entry points
   (fp_curfit_c, fp_curfit_batch_c, ...) run on; 0 when that device has not been used yet.  Lets a
   caller verify that a batch sharded with ngpus > 1 really ran on several GPUs. */
int64_t fpb_default_engine_launch_count(int32_t device);
/* device time in ms of the last dominant kernel (fit or splev), measured with
   CUDA events on the engine's stream; valid after fpb_engine_synchronize */
double  fpb_engine_last_kernel_ms(fpb_engine *eng);
/* FP64 pipe probe: achieved 1e9 FP64 instructions/s of a register-resident kernel issuing
   DMUL+DADD pairs (fused=0; the only forms the parity-exact kernels may use) or DFMA
   (fused=1); 2 / 3: IEEE divisions / square roots per second.  Roofline denominators of the fit kernel.
   fused = 4 / 5 / 6: DEPENDENT-issue latency in SM cycles per instruction of a DMUL+DADD chain, a DFMA chain,
   a MUFU.RCP64H+DFMA chain (one warp): the fit kernels are chains of dependent FP64 instructions, so their
   issue rate is bounded by (warps per scheduler) / latency. */
double  fpb_probe_fp64_gops(fpb_engine *eng, int32_t fused);
/* Self-test of the hand-written fast-path sequences the fit kernels use in place of the IEEE operators
   (DESIGN.md 4.1): `samples` random operand sets per check, bitwise compared on the device with `/`, sqrt() and
   the plain fpgivs.  mismatches[0..2] = division, square root on [1,2], fpgivs (must all be 0). */
int32_t fpb_selftest_fastmath(fpb_engine *eng, int64_t samples, int64_t seed, int64_t *mismatches);
int32_t fpb_device_count(void);
const char *fpb_last_error(void);
const char *fpb_version(void);

/* ------------------------------------------------------------------ (B) -- */
/*
 * Batched curfit on DEVICE pointers (the hot path).  One thread per curve.
 * Element i of curve b of array A is A[b*A_sb + i*A_si]:
 *   point-major / SoA (fast, coalesced):  A_sb = 1, A_si = ld (>= nbatch)
 *   curve-major:                          A_sb = ld (>= m), A_si = 1
 *   shared by all curves:                 A_sb = 0
 * w == NULL means unit weights.  xb/xe: per-curve arrays (xbe_stride=1), one
 * shared value (xbe_stride=0) or NULL (use x(1), x(m) of each curve).
 * s: per-curve (s_stride=1) or shared (s_stride=0).
 * Outputs are curve-major: n[b], fp[b], ier[b], t[b*nest+i], c[b*nest+i].
 * iopt=-1: n and t are inputs (interior knots in place, as for curfit).
 * iopt= 1: n, t, fpint, nrdata are in/out state of the previous call
 *          (fpint,nrdata: [nbatch][nest]); they may be NULL for iopt<=0.
 * stats (optional, [nbatch][2]): LSQ passes and p-iterations per curve.
 * Per-curve argument errors give ier[b]=10 exactly as curfit does; the return
 * value only reports call-level problems (FPB_ERR_*).
 */
int32_t fp_curfit_batch_dev_c(fpb_engine *eng, FP_SIZE nbatch, FP_SIZE iopt, FP_SIZE m,
                              const FP_REAL *x, int64_t x_sb, int64_t x_si,
                              const FP_REAL *y, int64_t y_sb, int64_t y_si,
                              const FP_REAL *w, int64_t w_sb, int64_t w_si,
                              const FP_REAL *xb, const FP_REAL *xe, FP_SIZE xbe_stride,
                              FP_SIZE k, const FP_REAL *s, FP_SIZE s_stride, FP_SIZE nest,
                              FP_SIZE *n, FP_REAL *t, FP_REAL *c, FP_REAL *fp, FP_FLAG *ier,
                              FP_REAL *fpint, FP_SIZE *nrdata, FP_SIZE *stats);

/*
 * Shared-abscissa least-squares fast path (iopt=-1, one x/w/knot set for the
 * whole batch): the Givens rotations depend only on (x,w,t) and are generated
 * once; every curve applies them to its own right-hand side.
 * y is strided as above; outputs c[b*ldc + i] (i < n-k-1), fp[b], ier (scalar,
 * from the argument/knot checks of curfit+fpchec).  x,w,t are DEVICE arrays of
 * length m, m, n (w may be NULL).
 */
int32_t fp_curfit_shared_dev_c(fpb_engine *eng, FP_SIZE nbatch, FP_SIZE m,
                               const FP_REAL *x, const FP_REAL *w,
                               const FP_REAL *y, int64_t y_sb, int64_t y_si,
                               FP_REAL xb, FP_REAL xe, FP_SIZE k, FP_SIZE n, FP_REAL *t,
                               FP_REAL *c, int64_t ldc, FP_REAL *fp, FP_FLAG *ier);

/*
 * Batched splev on DEVICE pointers.  Spline j has knots t[j*ldt + 0..n[j]) and
 * coefficients c[j*ldt + ...]; it is evaluated at x[j*m + 0..m) into y[j*m + ..].
 * nshared != 0: a single spline (j=0) evaluated at all nspl*m points.
 * lidx (optional, same shape as y): the 1-based knot interval index used.
 * ier[j] as splev: 0, 10 (m<1), or 6 (e=2 and a point outside the support; y of
 * that spline is then unspecified from the first offender on, as in the reference).
 */
int32_t fp_splev_batch_dev_c(fpb_engine *eng, FP_SIZE nspl, const FP_REAL *t, const FP_SIZE *n,
                             const FP_REAL *c, FP_SIZE ldt, FP_SIZE k, const FP_REAL *x,
                             FP_REAL *y, FP_SIZE m, FP_FLAG e, FP_SIZE mode, FP_SIZE nshared,
                             FP_FLAG *ier, FP_SIZE *lidx);

/*
 * Host-pointer batch calls (curve-major arrays, as a Fortran/C caller holds
 * them): copy in, tile-transpose on the device, run the kernels above, copy
 * out.  ngpus>1 shards the batch in contiguous slices over the first `ngpus`
 * devices (independent curves: no collective).  ngpus<=0: all visible devices.
 *   x,y,w : [nbatch][m]  (w NULL => 1; shared_x!=0 => x,w are [m])
 *   xb,xe : NULL => data range of each curve; else scalars by pointer
 */
int32_t fp_curfit_batch_c(FP_SIZE nbatch, FP_SIZE iopt, FP_SIZE m, const FP_REAL *x,
                          const FP_REAL *y, const FP_REAL *w, FP_SIZE shared_x,
                          const FP_REAL *xb, const FP_REAL *xe, FP_SIZE k, FP_REAL s, FP_SIZE nest,
                          FP_SIZE *n, FP_REAL *t, FP_REAL *c, FP_REAL *fp, FP_FLAG *ier,
                          FP_SIZE ngpus);

int32_t fp_splev_batch_c(FP_SIZE nspl, const FP_REAL *t, const FP_SIZE *n, const FP_REAL *c,
                         FP_SIZE ldt, FP_SIZE k, const FP_REAL *x, FP_REAL *y, FP_SIZE m,
                         FP_FLAG e, FP_SIZE mode, FP_FLAG *ier, FP_SIZE ngpus);

/* device tile transpose helper: dst[j*ld_dst + i] = src[i*ld_src + j], i<rows, j<cols */
int32_t fpb_transpose_dev(fpb_engine *eng, const FP_REAL *src, int64_t ld_src, FP_REAL *dst,
                          int64_t ld_dst, int64_t rows, int64_t cols);

/* ----------------------------------------------------------- spline calculus -- */
/* reference: include/fitpack_core_c.h:134-138,149-150; src/fitpack_core_c.f90:157-175,207-213
   -> splder core:17699-17802, spalde core:16865-16895, splint core:17919-17939.
   Host pointers, same argument lists.  splder: e = 0 extrapolate, 1 zero, 2 => ier=1 at the first
   point outside the support (any other value extrapolates); wrk(n) returns the derivative's
   B-spline coefficients.  spalde: d(j) = s^(j-1)(x), j = 1..k1.  splint: wrk(n) returns the
   integrals of the normalized B-splines. */
void fp_splder_c(const FP_REAL *t, FP_SIZE n, const FP_REAL *c, FP_SIZE k, FP_SIZE nu, const FP_REAL *x,
                 FP_REAL *y, FP_SIZE m, FP_FLAG e, FP_REAL *wrk, FP_FLAG *ier);
void fp_spalde_c(const FP_REAL *t, FP_SIZE n, const FP_REAL *c, FP_SIZE k1, FP_REAL x, FP_REAL *d,
                 FP_FLAG *ier);
FP_REAL fp_splint_c(const FP_REAL *t, FP_SIZE n, const FP_REAL *c, FP_SIZE k, FP_REAL a, FP_REAL b,
                    FP_REAL *wrk);
/* fourco (include/fitpack_core_c.h:152-153, src/fitpack_core_c.f90:216-223 -> core:1474-1520): the integrals
   ress(i) = int s(x) sin(alfa(i) x) dx, resc(i) = int s(x) cos(alfa(i) x) dx over [t(4), t(n-3)] of a CUBIC spline;
   ier = 13 for n < 10, 10 for invalid knots.  wrk1, wrk2 (length n) return the B-spline integrals of the last alfa.
   One thread per alfa; sin / cos are CUDA's (<= 2 ulp), so values agree with the reference to ~1e-13 relative, the
   one routine of this library that is not bit-identical by construction. */
void fp_fourco_c(const FP_REAL *t, FP_SIZE n, const FP_REAL *c, const FP_REAL *alfa, FP_SIZE m, FP_REAL *ress,
                 FP_REAL *resc, FP_REAL *wrk1, FP_REAL *wrk2, FP_FLAG *ier);
/* cualde (include/fitpack_core_c.h, src/fitpack_core_c.f90:187-194 -> core:1062-1101): all derivatives of a parametric
   curve at u, d((j-1)*idim + i) = d^(j-1) s_i / du^(j-1).  insert (src/fitpack_core_c.f90:197-204 -> insert_inplace
   core:15138-15159, fpinst 7961-8026): the knot x inserted into (t, n, c) in place; iopt /= 0: periodic spline; on
   ier = 10 nothing changes.  Both bit-identical to the reference's arithmetic. */
void fp_cualde_c(FP_SIZE idim, const FP_REAL *t, FP_SIZE n, const FP_REAL *c, FP_SIZE nc, FP_SIZE k1, FP_REAL u,
                 FP_REAL *d, FP_SIZE nd, FP_FLAG *ier);
void fp_insert_c(FP_SIZE iopt, FP_REAL *t, FP_SIZE *n, FP_REAL *c, FP_SIZE k, FP_REAL x, FP_SIZE nest, FP_FLAG *ier);
/* sproot (src/fitpack_core_c.f90:226-233 -> core:17962-18109, fpcuro 5108-5198): the zeros of a CUBIC spline in
   [t(4), t(n-3)], ascending, duplicates removed; m = their number; ier = 13 for n < 8, 10 for invalid knots, 1 when more
   than mest zeros exist (the first mest found are returned).  pow / atan2 / cos are CUDA's: rounding-level agreement. */
void fp_sproot_c(const FP_REAL *t, FP_SIZE n, const FP_REAL *c, FP_REAL *zeros, FP_SIZE mest, FP_SIZE *m, FP_FLAG *ier);

/* ---------------------------------------------------------- gridded surfaces -- */
/*
 * regrid / bispev, the dims = 2 case of the reference's gridded engine.
 * Same names and argument lists as the reference's C binding:
 *   fp_regrid_c  include/fitpack_core_c.h:165-171, src/fitpack_core_c.f90:253-290
 *                -> regrid core:16732-16839, fpregr 12079-12303, fpgrre 7111-7360
 *   fp_bispev_c  include/fitpack_core_c.h:209-211, src/fitpack_core_c.f90:404-415
 *                -> bispev core:424-458, fpndsp 2160-2241
 * z and c are flat row-major with x slowest: z[my*i + j], c[(ny-ky-1)*i + j].
 * Workspace: lwrk/kwrk must satisfy the reference's minimum sizes (core:16789-16797), else
 * ier=10; only wrk[0..3] (fp0, fpold, reduc) and iwrk[0..2] (lastdi, nplus) are used -- they
 * are the state an iopt=1 continuation reads back, as in the reference.
 * fp and the smoothing parameter are grid-wide sums: they agree with the reference to rounding
 * (rel 1e-10 stated in the tests), n / knots / ier are the same discrete outputs.
 */
void fp_regrid_c(FP_SIZE iopt, FP_SIZE mx, const FP_REAL *x, FP_SIZE my, const FP_REAL *y,
                 const FP_REAL *z, FP_REAL xb, FP_REAL xe, FP_REAL yb, FP_REAL ye,
                 FP_SIZE kx, FP_SIZE ky, FP_REAL s, FP_SIZE nxest, FP_SIZE nyest,
                 FP_SIZE *nx, FP_REAL *tx, FP_SIZE *ny, FP_REAL *ty, FP_REAL *c, FP_REAL *fp,
                 FP_REAL *wrk, FP_SIZE lwrk, FP_SIZE *iwrk, FP_SIZE kwrk, FP_FLAG *ier);

FP_FLAG fp_bispev_c(const FP_REAL *tx, FP_SIZE nx, const FP_REAL *ty, FP_SIZE ny, const FP_REAL *c,
                    FP_SIZE kx, FP_SIZE ky, const FP_REAL *x, FP_SIZE mx, const FP_REAL *y, FP_SIZE my,
                    FP_REAL *z, FP_REAL *wrk, FP_SIZE lwrk, FP_SIZE *iwrk, FP_SIZE kwrk);

/* NEW: the same with the mx*my grid values resident in HBM (z_dev is a DEVICE pointer; every
   other array is a small host array).  Returns FPB_STATUS_* ; ier as regrid / bispev. */
int32_t fp_regrid_dev_c(fpb_engine *eng, FP_SIZE iopt, FP_SIZE mx, const FP_REAL *x, FP_SIZE my,
                        const FP_REAL *y, const FP_REAL *z_dev, FP_REAL xb, FP_REAL xe, FP_REAL yb,
                        FP_REAL ye, FP_SIZE kx, FP_SIZE ky, FP_REAL s, FP_SIZE nxest, FP_SIZE nyest,
                        FP_SIZE *nx, FP_REAL *tx, FP_SIZE *ny, FP_REAL *ty, FP_REAL *c, FP_REAL *fp,
                        FP_REAL *wrk, FP_SIZE lwrk, FP_SIZE *iwrk, FP_SIZE kwrk, FP_FLAG *ier);
int32_t fp_bispev_dev_c(fpb_engine *eng, const FP_REAL *tx, FP_SIZE nx, const FP_REAL *ty, FP_SIZE ny,
                        const FP_REAL *c, FP_SIZE kx, FP_SIZE ky, const FP_REAL *x, FP_SIZE mx,
                        const FP_REAL *y, FP_SIZE my, FP_REAL *z_dev, FP_FLAG *ier);
/* number of fpgrre solves (knot sets + smoothing-parameter iterations) of the last regrid call */
int64_t fpb_engine_grid_fpgrre_calls(const fpb_engine *eng);

/* -------------------------------------------------------- parametric curves -- */
/*
 * parcur / curev: smoothing fit of idim coordinate curves x_d(u) that share one
 * knot vector, and its evaluation.  Same names and argument lists as the reference's C binding:
 *   fp_parcur_c  include/fitpack_core_c.h:109-112, src/fitpack_core_c.f90:90-100
 *                -> parcur core:15201-15285, fppara core:8822-9177
 *   fp_curev_c   include/fitpack_core_c.h:140-141, src/fitpack_core_c.f90:178-184 -> curev core:1126-1182
 * x is x(idim,m) (coordinate fastest), c holds coordinate d at c[d*n .. d*n + n-k-2], u/ub/ue are
 * in/out (ipar=0: u = normalised chord length, ub=0, ue=1), wrk(1:nest)/iwrk(1:nest) carry the
 * iopt=1 continuation state.  idim <= 10, lwrk >= m*(k+1)+nest*(6+idim+3*k), nc >= nest*idim.
 */
void fp_parcur_c(FP_SIZE iopt, FP_SIZE ipar, FP_SIZE idim, FP_SIZE m, FP_REAL *u, FP_SIZE mx,
                 const FP_REAL *x, FP_REAL *w, FP_REAL *ub, FP_REAL *ue, FP_SIZE k, FP_REAL *s,
                 FP_SIZE nest, FP_SIZE *n, FP_REAL *t, FP_SIZE nc, FP_REAL *c, FP_REAL *fp, FP_REAL *wrk,
                 FP_SIZE lwrk, FP_SIZE *iwrk, FP_FLAG *ier);
void fp_curev_c(FP_SIZE idim, const FP_REAL *t, FP_SIZE n, const FP_REAL *c, FP_SIZE nc, FP_SIZE k,
                const FP_REAL *u, FP_SIZE m, FP_REAL *x, FP_SIZE mx, FP_FLAG *ier);

/* NEW: many independent parametric curves.  Host pointers, curve-major: x [nbatch][m][idim],
   u [nbatch][m] in/out, w [nbatch][m] or NULL (unit weights), ub/ue [nbatch] in/out,
   t [nbatch][nest], c [nbatch][nest*idim], fpint/nrdata [nbatch][nest] continuation state (NULL
   allowed unless iopt=1).  ngpus: contiguous slices of the batch over that many GPUs of the node
   (<= 0: all), no collective.  Per-curve ier; returns FPB_STATUS_*. */
int32_t fp_parcur_batch_c(FP_SIZE nbatch, FP_SIZE iopt, FP_SIZE ipar, FP_SIZE idim, FP_SIZE m, FP_REAL *u,
                          const FP_REAL *x, const FP_REAL *w, FP_REAL *ub, FP_REAL *ue, FP_SIZE k,
                          FP_REAL s, FP_SIZE nest, FP_SIZE *n, FP_REAL *t, FP_REAL *c, FP_REAL *fp,
                          FP_FLAG *ier, FP_REAL *fpint, FP_SIZE *nrdata, FP_SIZE ngpus);
/* Device pointers + strides (elements): u[b*u_sb + i*u_si], x[b*x_sb + i*x_si + d*x_sd],
   w[b*w_sb + i*w_si] (NULL = unit weights); the fast layout is point-major (u_sb = 1, u_si = ld;
   x_sb = 1, x_si = idim*ld, x_sd = ld).  ub/ue NULL = the data range [u(1), u(m)] (always used
   when ipar=0 and iopt<=0, where u is computed on the device).  Outputs as
   fp_curfit_batch_dev_c, c [nbatch][nest*idim]. */
int32_t fp_parcur_batch_dev_c(fpb_engine *eng, FP_SIZE nbatch, FP_SIZE iopt, FP_SIZE ipar, FP_SIZE idim,
                              FP_SIZE m, FP_REAL *u, int64_t u_sb, int64_t u_si, const FP_REAL *x,
                              int64_t x_sb, int64_t x_si, int64_t x_sd, const FP_REAL *w, int64_t w_sb,
                              int64_t w_si, const FP_REAL *ub, const FP_REAL *ue, FP_SIZE ube_stride,
                              FP_SIZE k, const FP_REAL *s, FP_SIZE s_stride, FP_SIZE nest, FP_SIZE *n,
                              FP_REAL *t, FP_REAL *c, FP_REAL *fp, FP_FLAG *ier, FP_REAL *fpint,
                              FP_SIZE *nrdata, FP_SIZE *stats);
/* curev for many curves: t [nspl][ldt], c [nspl][ldc] (coordinate d at d*n[j]), u [nspl][m]
   non-decreasing, x [nspl][m][idim]; per-curve ier (10 = decreasing u or invalid n). */
int32_t fp_curev_batch_dev_c(fpb_engine *eng, FP_SIZE nspl, FP_SIZE idim, const FP_REAL *t,
                             const FP_SIZE *n, const FP_REAL *c, FP_SIZE ldt, int64_t ldc, FP_SIZE k,
                             const FP_REAL *u, FP_REAL *x, FP_SIZE m, FP_FLAG *ier);

/* ---------------------------------------------------------- periodic curves -- */
/*
 * percur: periodic smoothing spline, period x(m)-x(1); y(m) and w(m) are not
 * used (core:15744-15783).  Same name and argument list as the reference's C binding:
 *   fp_percur_c  include/fitpack_core_c.h:104-106, src/fitpack_core_c.f90:77-87
 *                -> percur core:15784-15863, fpperi core:9668-10192
 * lwrk >= m*(k+1)+nest*(8+5*k); wrk(1:nest)/iwrk(1:nest) carry the iopt=1 state.  Evaluate the
 * result with fp_splev_c.
 */
void fp_percur_c(FP_SIZE iopt, FP_SIZE m, const FP_REAL *x, const FP_REAL *y, const FP_REAL *w, FP_SIZE k,
                 FP_REAL s, FP_SIZE nest, FP_SIZE *n, FP_REAL *t, FP_REAL *c, FP_REAL *fp, FP_REAL *wrk,
                 FP_SIZE lwrk, FP_SIZE *iwrk, FP_FLAG *ier);
/* NEW: many independent periodic fits; host pointers curve-major (x, y, w [nbatch][m], w NULL = 1) */
int32_t fp_percur_batch_c(FP_SIZE nbatch, FP_SIZE iopt, FP_SIZE m, const FP_REAL *x, const FP_REAL *y,
                          const FP_REAL *w, FP_SIZE k, FP_REAL s, FP_SIZE nest, FP_SIZE *n, FP_REAL *t,
                          FP_REAL *c, FP_REAL *fp, FP_FLAG *ier, FP_REAL *fpint, FP_SIZE *nrdata, FP_SIZE ngpus);
/* device pointers + strides as fp_curfit_batch_dev_c (point-major is the fast layout) */
int32_t fp_percur_batch_dev_c(fpb_engine *eng, FP_SIZE nbatch, FP_SIZE iopt, FP_SIZE m, const FP_REAL *x,
                              int64_t x_sb, int64_t x_si, const FP_REAL *y, int64_t y_sb, int64_t y_si,
                              const FP_REAL *w, int64_t w_sb, int64_t w_si, FP_SIZE k, const FP_REAL *s,
                              FP_SIZE s_stride, FP_SIZE nest, FP_SIZE *n, FP_REAL *t, FP_REAL *c, FP_REAL *fp,
                              FP_FLAG *ier, FP_REAL *fpint, FP_SIZE *nrdata);

#ifdef __cplusplus
}
#endif
#endif /* FITPACK_B200_H_INCLUDED */
