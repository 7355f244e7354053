// fpb_kernels.h -- host-visible declarations of the CUDA kernels' launchers.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

namespace fpb {

constexpr int kFitThreads = 128;  // threads per CTA of the fit kernel (4 warps = 128 curves)
// CTAs per SM the fit kernels are compiled for (register budget 65536/(128*blocks))
#ifndef FPB_FIT_BLOCKS_K3
#define FPB_FIT_BLOCKS_K3 5
#endif
#ifndef FPB_FIT_BLOCKS_DN
#define FPB_FIT_BLOCKS_DN 5   // idim = 2, 3 instantiations, k <= 3 (scan at idim 2: 3 -> 2.95, 4 -> 3.33, 5 -> 3.41 M fits/s)
#endif
constexpr int kMaxIdim = 10;      // MAX_IDIM of the reference (core:115)
// the fit kernels are compiled for idim = 1, 2, 3 and a run-time variant (class 0) for 4..10
constexpr int idim_class(int idim) { return idim <= 3 ? idim : 0; }
constexpr int fit_blocks_per_sm(int k, int idim_cls = 1)
{
    return idim_cls == 1 ? (k <= 3 ? FPB_FIT_BLOCKS_K3 : (k == 4 ? 4 : 3))
         : idim_cls == 0 ? 2
                         : (k <= 3 ? FPB_FIT_BLOCKS_DN : 3);
}
// the part-2 kernel (smoothing-parameter iteration: one long dependent rotation chain per curve) may
// be compiled for a different number of resident CTAs than the pass kernel
#ifndef FPB_FIT_P2_BLOCKS_K3
#define FPB_FIT_P2_BLOCKS_K3 FPB_FIT_BLOCKS_K3
#endif
constexpr int fit_p2_blocks_per_sm(int k, int idim_cls = 1)
{
    return (idim_cls == 1 && k <= 3) ? FPB_FIT_P2_BLOCKS_K3 : fit_blocks_per_sm(k, idim_cls);
}
// shared-memory band window: (k+1)^2 + (k+1)*idim doubles per thread
constexpr unsigned fit_smem_bytes(int k, int idim = 1)
{
    return (unsigned)(((k + 1) * (k + 1) + (k + 1) * idim) * kFitThreads * sizeof(double));
}

// curfit / parcur kernels: band window + the cp.async ring of the point streams (`depth` slots of
// x, y(1:idim) and, with weights, w per thread)
constexpr int fit_window_doubles(int k, int idim) { return (k + 1) * (k + 1) + (k + 1) * idim; }
#ifndef FPB_FIT_SPAN_SMEM
#define FPB_FIT_SPAN_SMEM 0   // span reciprocals of bspl_tab in registers (1: in shared memory, +9 ms measured)
#endif
constexpr unsigned fit_smem_bytes_stream(int k, int idim, bool has_w, int depth)
{
    // (+ k(k+1)/2 span reciprocals per thread when they live in shared memory)
    return (unsigned)((fit_window_doubles(k, idim) + depth * (1 + idim + (has_w ? 1 : 0)) +
                       (FPB_FIT_SPAN_SMEM ? k * (k + 1) / 2 : 0)) *
                      kFitThreads * sizeof(double));
}
// deepest ring (8, 4 or 2 slots) that still lets `blocks` CTAs share the SM's 227 KB (1 KB reserved per CTA)
#ifndef FPB_FIT_STREAM
#define FPB_FIT_STREAM 0   // measured (profiles/r02_ab.md): the cp.async ring loses to plain loads + L2 prefetch
#endif
#ifndef FPB_FIT_STREAM_MAXDEPTH
#define FPB_FIT_STREAM_MAXDEPTH 8
#endif
constexpr int fit_stream_depth(int k, int idim, bool has_w, int blocks)
{
    return !FPB_FIT_STREAM ? 0 : FPB_FIT_STREAM_MAXDEPTH < 8 ? FPB_FIT_STREAM_MAXDEPTH : (fit_smem_bytes_stream(k, idim, has_w, 8) + 1024u) * (unsigned)blocks <= 227u * 1024u   ? 8
           : (fit_smem_bytes_stream(k, idim, has_w, 4) + 1024u) * (unsigned)blocks <= 227u * 1024u ? 4
                                                                                                   : 2;
}

// per-lane workspace layout of the fit kernel, in elements of 8 bytes
struct FitLayout {
    int nestp;  // knots a curve can actually reach: min(nest, m+k+1)
    int off_t, off_c, off_z, off_fpint, off_a, off_g, off_b, off_q, off_x, off_y, off_w;
    int off_a2, off_g2;  // periodic fits: wrap-around column blocks (nestp x k, nestp x (k+1))
    int ws_per_lane;
};
FitLayout fit_layout(int m, int k, int nest, int idim = 1, int periodic = 0);
int fit_max_resident_warps(int k, int idim, int nsm);
int fit_p2_resident_warps(int k, int idim, int nsm);

struct FitParams {
    int nbatch, iopt, m, k, nest, nestp;
    int sdepth;             // slots of the point-stream ring (fit_stream_depth)
    int staged;             // 1: pass lists may hold "no pass yet" entries (host batches arriving in chunks)
    int idim;               // coordinate curves per fit (1 for curfit)
    int par;                // 1: parcur/fppara semantics (checks, initial p, no zero-pivot skip)
    int periodic;           // 1: percur/fpperi (fpb_percur.cu kernels, two-block factor)
    const double *x, *y, *w;  // x: abscissae / curve parameters u; y: ordinates / coordinates
    long long x_sb, x_si, y_sb, y_si, w_sb, w_si;
    long long y_sd;         // stride between the coordinates of one point in y
    long long c_sb;         // stride between curves in c (nest*idim)
    const double *xb, *xe;
    int xbe_stride;
    const double *s;
    int s_stride;
    int *n;
    double *t, *c, *fp;
    int *ier;
    double *fpint;
    int *nrdata;
    int *stats;
    double *ws;             // [resident warps][ws_per_lane][32]
    int *wsi;               // [resident warps][nestp][32]   (nrdata)
    unsigned int *counter;  // tile queue head, zeroed before every launch
    int off_t, off_c, off_z, off_fpint, off_a, off_g, off_b, off_q, off_x, off_y, off_w, ws_per_lane;
    int off_a2, off_g2;
    // per-curve state carried between passes (engine-owned)
    int *st_nplus, *st_iter;
    double *st_fpold, *st_fp0;
    int *st_nrd;            // [nbatch][nestp]  nrdata(1:nrint)
    int *hist;              // [nestp+1] histogram of n over the part-2 list
    // triangular system (R, z) of the accepted knot set, handed from the last pass to part 2 so that part 2
    // does not have to regenerate it: rows of (k+1)+idim doubles, allocated from a pool through az_cursor;
    // az_off[b] = first element of curve b's rows, or -1 (pool full / no pass ran: part 2 recomputes)
    double *az_pool;
    unsigned long long *az_cursor;
    long long az_cap;       // pool capacity in doubles
    long long *az_off;      // [nbatch]
};
cudaError_t launch_curfit_pass(const FitParams &p, int grid_blocks, cudaStream_t stream,
                               const int *in_list, const int *in_count, int *out_list,
                               int *out_count, int *p2_list, int *p2_count, int first);
cudaError_t launch_curfit_part2(const FitParams &p, int grid_blocks, cudaStream_t stream,
                                const int *list, const int *count);
#define FPB_DECL_FIT_D(D)                                                                           \
    cudaError_t launch_curfit_pass_d##D(const FitParams &p, int grid, cudaStream_t st,               \
                                        const int *in_list, const int *in_count, int *out_list,      \
                                        int *out_count, int *p2_list, int *p2_count, int first);     \
    cudaError_t launch_curfit_part2_d##D(const FitParams &p, int grid, cudaStream_t st,              \
                                         const int *list, const int *count);
FPB_DECL_FIT_D(0) FPB_DECL_FIT_D(1) FPB_DECL_FIT_D(2) FPB_DECL_FIT_D(3)
#undef FPB_DECL_FIT_D
// periodic fits (fpb_percur.cu): same launch signatures as the curfit kernels
cudaError_t launch_percur_pass(const FitParams &p, int grid, cudaStream_t st, const int *in_list,
                               const int *in_count, int *out_list, int *out_count, int *p2_list,
                               int *p2_count, int first);
cudaError_t launch_percur_part2(const FitParams &p, int grid, cudaStream_t st, const int *list,
                                const int *count);
int percur_resident_warps(int k, int nsm);
// a few periodic curves: one CTA per curve, workspace in shared memory, one launch (fpb_percur_solo.cu)
size_t percur_solo_smem_bytes(int m, int k, int ws_per_lane, int nestp);
cudaError_t launch_percur_solo(const FitParams &p, cudaStream_t stream);
// chord-length parameters of parcur (ipar = 0), one thread per curve
cudaError_t launch_parcur_param(int nbatch, int m, int idim, const double *x, long long x_sb,
                                long long x_si, long long x_sd, double *u, long long u_sb,
                                long long u_si, cudaStream_t stream);
// curves c0 .. c0+nc-1 join a running pipeline: appended to `list` (which holds `base` entries), *count = base+nc
cudaError_t launch_fit_admit(int *list, int *count, int base, int c0, int nc, cudaStream_t stream);
cudaError_t launch_p2_sort(const int *list, const int *count, const int *n_arr, int *hist,
                           int nbins, int *sorted, int nsm, cudaStream_t stream);

// ---- shared-abscissa least squares (iopt=-1, one x/w/t for the whole batch) ----
struct SharedPlan {
    // device buffers owned by the engine
    double *cs_sn;   // [m][k1][2]  rotation parameters per (point, band column)
    double *wgt;     // [m]         weights (1 when w==NULL)
    int *lrow;       // [m]         first band row (0-based) the point touches = l-k1
    int *skip;       // [m]         bit i set: pivot i was (sub)normal-zero -> rotation skipped
    double *R;       // [nk1][k1]   final triangular band factor
    int *ier;        // [1]
};
cudaError_t launch_shared_plan(const double *x, const double *w, int m, double xb, double xe, int k,
                               int n, double *t, const SharedPlan &plan, cudaStream_t stream);
cudaError_t launch_shared_apply(int nbatch, int m, int k, int n, const double *y, long long y_sb,
                                long long y_si, const SharedPlan &plan, const double *w, int *meta,
                                double *c, long long ldc, double *fp, int nsm, cudaStream_t stream,
                                int fast);

// ---- evaluation ----
struct SplevParams {
    int nspl, ldt, k, m, e, mode, nshared;
    int tab = 0;            // EXACT: span table of bspl_tab in shared memory (set by launch_splev)
    // FAST mode, used by curev's fast path (one launch per coordinate d): coefficients of spline j start at
    // c[j*ldc + c_mul*n(j)], results go to y[(j*m + i)*y_stride + y_off].  Plain splev: ldc = ldt, 0, 1, 0.
    long long ldc = 0;      // 0: ldt
    int c_mul = 0, y_stride = 1, y_off = 0;
    const double *t, *c, *x;
    const int *n;
    double *y;
    int *ier;
    int *lidx;
};
cudaError_t launch_splev(const SplevParams &p, int nsm, cudaStream_t stream);

// ---- utilities ----
cudaError_t launch_transpose(const double *src, long long ld_src, double *dst, long long ld_dst,
                             long long rows, long long cols, cudaStream_t stream);
cudaError_t launch_fill_i32(int *dst, int value, long long count, cudaStream_t stream);
// dense t(1:n), c(1:n-k-1) of every curve with ier != 10 (D2H staging of the host-pointer batch fit)
cudaError_t launch_pack_tc(int nb, int nest, int k, const int *n, const int *ier, const double *t, const double *c,
                           unsigned long long *cursor, long long *off, double *tp, double *cp, int nsm,
                           cudaStream_t stream);
// ---- a few curves, one warp per curve (fpb_curfit_solo.cu): curfit (idim = 1, par = 0) or parcur (par = 1) ----
struct SoloParams {
    int nbatch, iopt, m, k, nest;
    int idim, par, chord;      // right-hand sides per point; fppara semantics; parcur computes chord-length parameters
    double xb, xe, s;
    int ends_from_x;           // curfit: 1: xb = x(1), xe = x(m) of each curve
    const double *x;           // [nbatch][m]        abscissae (curfit) / parameters u (parcur, unless chord)
    const double *y;           // [nbatch][m][idim]  data
    const double *w;           // [nbatch][m]
    double *u_out;             // parcur with chord: the parameters computed, [nbatch][m]
    double *ube;               // parcur: [nbatch][2] ub, ue (in; out when chord)
    int *n;                    // [nbatch]        in (iopt != 0) / out
    double *t, *c;             // t [nbatch][nest] in (iopt != 0) / out; c [nbatch][c_stride] out
    long long c_stride;
    double *fp;                // [nbatch]        in / out (left alone when nothing was computed)
    int *ier;                  // [nbatch]        in: != 0 skips the curve; out
    double *fpint;             // [nbatch][nest]  in (iopt = 1) / out
    int *nrdata;               // [nbatch][nest]  in (iopt = 1) / out
};
size_t solo_smem_bytes(int m, int k, int nest, int idim);
bool solo_supports(int k, int idim, int par);
constexpr size_t kSoloSmemMax = 224 * 1024;   // of the 227 KB a CTA may have on sm_100
cudaError_t launch_curfit_solo(const SoloParams &p, cudaStream_t stream);

cudaError_t launch_fp64_probe(double *scratch, int blocks, int iters, int fused, cudaStream_t stream);
cudaError_t launch_fastmath_selftest(unsigned long long seed, int blocks, int iters, unsigned long long *mismatch,
                                     cudaStream_t stream);

}  // namespace fpb
