// fpb_regrid.cu -- gridded surfaces on sm_100a: regrid/fpregr/fpgrre (smoothing or least-squares
// tensor-product spline fit of data on a rectangular grid) and bispev/fpndsp (evaluation on a
// grid), the dims = 2 case the reference binds as fp_regrid_c / fp_bispev_c.
//
// Reference: regrid core:16732-16839, fpregr core:12079-12303, fpgrre core:7111-7360,
// fp_rotate_row_block / fp_rotate_row_stride core:6034-6092, new_knot_dimension_nd
// core:12347-12362, bispev core:424-458, fpndsp core:2160-2241; C binding
// src/fitpack_core_c.f90:253-290, 404-415 ("core" = src/fitpack_core.F90).
//
// Structure of one fpgrre call (Kronecker form A_y C A_x^T = Q, core:7052-7060):
//   grid_plan_kernel   one CTA per axis.  All threads: basis values sp / interval numbers nr of
//                      the grid coordinates (fpbspl) and the discontinuity rows b (fpdisc).
//                      Warp 0: the matrix half of the band QR -- the Givens rotations depend
//                      only on (grid, knots, p), never on z -- as a systolic pipeline (lane i
//                      performs rotation i of every row) that also records a rotation TAPE: per
//                      pipeline step the (cos,sin) of every lane and a control word {source row
//                      entering, band rows to move down the lanes}.
//   grid_apply_kernel  the right-hand-side half: 8 lanes per independent right-hand side replay
//                      the tape (pass x: my columns of z, pass y: nk1x rows of the reduced
//                      buffer): 4 mul + 2 add per step, no sqrt, no division ("one rotation, many
//                      right-hand sides": the access pattern of fp_rotate_row_block).
//   grid_backsub_kernel   per-line banded back-substitution (fpback, core:1808-1838): axis y, then x.
//   grid_resid_kernel  squared residual of every grid cell (reference operation order), reduced to
//                      per-row / per-column sums in a FIXED order (tile partials, no atomics).
// The knot-placement and smoothing-parameter control flow of fpregr (scalars and O(nest) arrays:
// fpknot, the nplus rule, the axis arbiter, fprati) runs on the host between launches and reads
// back only the mx + my row/column sums per fpgrre call.
//
// Parity: c comes out of the same IEEE operations in the same order as the reference (no FMA).
// fp and fpint are sums over the whole grid; the reference adds the mx*my terms one by one in
// row-major order, which no parallel machine can reproduce bit for bit, so fp/fpint (and what
// follows from them: p, and through p the smoothed c) agree to rounding (tests: rel 1e-10), while
// n, knots and ier are compared exactly on every test case.
#include "fpb_common.cuh"
#include "fpb_engine.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#include "../../include/fitpack_b200.h"

namespace fpb {

namespace {

constexpr int kCs = 8;   // (cos,sin) slots per rotation-list row (band <= k+2 <= 7)
constexpr int kAs = 8;   // row stride of the band / discontinuity matrices (k+2 <= 7 columns)
constexpr int kSp = 6;   // row stride of the basis tables (k+1 <= 6 values)
constexpr int kPlanChunk = 256;   // grid points staged per step for the sequential plan thread
constexpr int kApplyChunk = 128;  // rotation-list rows staged per step in the apply kernel

struct AxisDev {
    const double *x;  // [m] grid coordinates
    const double *t;  // [n] knots
    double *sp;       // [m][kSp]
    int *nr;          // [m]
    double *a;        // [n][kAs] triangular band factor (row-major)
    double *b;        // [n][kAs] discontinuity rows
    int2 *ctl;        // [steps] rotation tape control: {source row entering at this step or -1,
                      //         band rows to move down the lanes before this step}
    double2 *rcs;     // [steps][kCs] rotation tape: (cos,sin) of lane i at this step, (1,0) = none
    int *info;        // [0] steps on the tape, [1] band width of the factor (k+1 or k+2)
    int m, n, k;
};

// rightmost l in [k1, nk1] with t(l) <= arg (1-based); arg >= t(k1) is guaranteed by the callers.
// Same result as the reference's forward walk (core:7184-7188) and as fp_knot_interval.
__device__ __forceinline__ int grid_interval(const double *t, int k1, int nk1, double arg)
{
    int lo = k1, hi = nk1;  // invariant: t(lo) <= arg
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (t[mid - 1] <= arg) lo = mid;
        else hi = mid - 1;
    }
    return lo;
}

// one row of fpdisc (core:5453-5495): b(lmk, 1:k2) for the interior knot l = lmk + k1
template <int K>
__device__ void disc_row(const double *t, int n, int lmk, double *brow)
{
    constexpr int K1 = K + 1, K2 = K + 2;
    const int nk1 = n - K1, nrint = nk1 - K;
    auto T = [&](int i) { return t[i - 1]; };
    const double fac = (double)nrint / sub(T(nk1 + 1), T(K1));
    const int l = lmk + K1;
    double h[2 * K1];
#pragma unroll
    for (int j = 1; j <= K1; ++j) {
        const int ik = j + K1, lj = l + j, lk = lj - K2;
        h[j - 1] = sub(T(l), T(lk));
        h[ik - 1] = sub(T(l), T(lj));
    }
    int lp = lmk;
#pragma unroll
    for (int j = 1; j <= K2; ++j) {
        int jk = j;
        double prod = h[j - 1];
#pragma unroll
        for (int i = 1; i <= K; ++i) {
            ++jk;
            prod = mul(mul(prod, h[jk - 1]), fac);
        }
        const int lk = lp + K1;
        brow[j - 1] = sub(T(lk), T(lp)) / prod;
        ++lp;
    }
}

template <int K>
__device__ void plan_axis(const AxisDev &A, double p, double (*sp_s)[kSp], int *nr_s, double (*b_s)[kAs])
{
    constexpr int K1 = K + 1, K2 = K + 2;
    const int n = A.n, m = A.m, nk1 = n - K1;
    // ---- basis values and interval numbers (core:7176-7196), all threads ----
    for (int it = threadIdx.x; it < m; it += blockDim.x) {
        const double arg = A.x[it];
        const int l = grid_interval(A.t, K1, nk1, arg);
        double tw[2 * K], h[K1];
#pragma unroll
        for (int j = 0; j < 2 * K; ++j) tw[j] = A.t[l - K + j];
        bspl<K>(tw, arg, h);
#pragma unroll
        for (int j = 0; j < K1; ++j) A.sp[(size_t)it * kSp + j] = h[j];
        A.nr[it] = l - K1;
    }
    // ---- discontinuity rows (core:7199-7206), all threads ----
    if (p > 0.0 && n != 2 * K1) {
        for (int lmk = 1 + threadIdx.x; lmk <= nk1 - K1; lmk += blockDim.x)
            disc_row<K>(A.t, n, lmk, A.b + (size_t)(lmk - 1) * kAs);
    }
    __syncthreads();
    // ---- matrix half of the band QR (core:7222-7300) -- a systolic pipeline in warp 0 ----
    // Rotation i of a row needs (a) the row as left by its rotation i-1 and (b) band row irot+i as
    // left by the PREVIOUS row's rotation i.  Lane i therefore owns band row irot+i and performs
    // rotation i of every row; a row enters at lane 0 and moves one lane per step (its remaining
    // elements travel by __shfl_up, already shifted so that the pivot is always element 0).  k+2
    // rotations of consecutive rows run concurrently and the serial chain is one rotation per row
    // instead of k+1.  When the first band row advances (a new knot interval) the pipeline drains
    // and the band rows move down the lanes.  The arithmetic of every rotation is unchanged
    // (fpgivs / fprota, same operands, same order).  The other warps stage the basis rows /
    // interval numbers / discontinuity rows of the next kPlanChunk grid points in shared memory.
    const double pinv = (p > 0.0) ? 1.0 / p : 1.0;
    const int lane = threadIdx.x & 31;
    const bool w0 = threadIdx.x < 32;
    double W[K2], hcur[K2];   // my band row; the row I am rotating (pivot first)
#pragma unroll
    for (int q = 0; q < K2; ++q) {
        W[q] = 0.0;
        hcur[q] = 0.0;
    }
    int rcur = -1, bcur = 0;  // row in my lane (>= 0: a row, -1: bubble) and its band
    int irot_cur = 0, nrold = 0, iband = K1;
    int s_tape = 0, pend_d = 0;  // tape position; band-row shift to announce with the next step
    // one pipeline step: every lane hands its row to the next lane, lane 0 takes the new row.
    // The step is also one line of the rotation TAPE the apply kernels replay: lane i's (cos,sin)
    // (identity when it has no rotation) and a control word {source row entering, pending shift}.
    auto step = [&](bool inject, const double (&hnew)[K2], int src, int bnew) {
        double hin[K2];
#pragma unroll
        for (int q = 0; q < K2 - 1; ++q) hin[q] = __shfl_up_sync(0xffffffffu, hcur[q + 1], 1);
        hin[K2 - 1] = 0.0;
        int rin = __shfl_up_sync(0xffffffffu, rcur, 1);
        int bin = __shfl_up_sync(0xffffffffu, bcur, 1);
        if (lane == 0) {
            rin = inject ? 0 : -1;
            bin = bnew;
#pragma unroll
            for (int q = 0; q < K2; ++q) hin[q] = hnew[q];
            A.ctl[s_tape] = make_int2(inject ? src : -1, pend_d);
        }
        double cs = 1.0, sn = 0.0;  // identity == "no rotation" / "pivot is zero: skip" (core:6043)
        if (rin >= 0 && lane < bin) {
            const double piv = hin[0];
            if (!fp_is_zero(piv)) {
                givens_fast(piv, W[0], cs, sn);   // same bits as givens(), a shorter dependent chain
#pragma unroll
                for (int q = 1; q < K2; ++q)
                    if (lane + q < bin) rota(cs, sn, hin[q], W[q]);
            }
        }
        if (lane < kCs) A.rcs[(size_t)s_tape * kCs + lane] = make_double2(cs, sn);
        rcur = rin;
        bcur = bin;
#pragma unroll
        for (int q = 0; q < K2; ++q) hcur[q] = hin[q];
        ++s_tape;
        pend_d = 0;
    };
    const double hzero[K2] = {};
    // first band row moves from irot_cur to irot_new: finish the rows in flight, retire the band
    // rows that are final, pass the others down the lanes
    auto advance = [&](int irot_new) {
        for (int q = 0; q < K2 - 1; ++q) step(false, hzero, -1, 0);
        rcur = -1;
        const int d = irot_new - irot_cur;
        if (lane < K2 && lane < d && irot_cur + lane < nk1) {
#pragma unroll
            for (int q = 0; q < K2; ++q) A.a[(size_t)(irot_cur + lane) * kAs + q] = W[q];
        }
#pragma unroll
        for (int q = 0; q < K2; ++q) {
            const double v = __shfl_down_sync(0xffffffffu, W[q], (unsigned)min(d, 31));
            W[q] = (d < K2 && lane + d < K2) ? v : 0.0;
        }
        // band rows jumped over without ever being touched stay zero (as in the reference)
        for (int row = irot_cur + K2 + lane; row < irot_new && row < nk1; row += 32) {
#pragma unroll
            for (int q = 0; q < K2; ++q) A.a[(size_t)row * kAs + q] = 0.0;
        }
        irot_cur = irot_new;
        pend_d = d;
    };
    for (int it0 = 0; it0 < m; it0 += kPlanChunk) {
        const int cnt = min(kPlanChunk, m - it0);
        __syncthreads();  // the previous chunk is consumed
        for (int e = threadIdx.x; e < cnt * K1; e += blockDim.x)
            sp_s[e / K1][e % K1] = A.sp[(size_t)(it0 + e / K1) * kSp + e % K1];
        for (int e = threadIdx.x; e < cnt; e += blockDim.x) nr_s[e] = A.nr[it0 + e];
        // smoothing rows this chunk can reach: b(nr(it0-1)+1 ..), at most one per grid point + 1
        const int b0 = (it0 > 0) ? A.nr[it0 - 1] : 0;
        if (p > 0.0)
            for (int e = threadIdx.x; e < kPlanChunk * K2; e += blockDim.x) {
                const int row = b0 + e / K2;
                if (row < nk1 - K1) b_s[e / K2][e % K2] = A.b[(size_t)row * kAs + e % K2];
            }
        __syncthreads();
        if (!w0) continue;
        for (int ii = 0; ii < cnt; ++ii) {
            const int it = it0 + ii;
            const int number = nr_s[ii];
            for (;;) {  // warp-uniform control flow (core:7236-7262)
                double h[K2];
                int irot, src;
                if (nrold == number) {
                    h[K2 - 1] = 0.0;  // h(iband) = 0 when iband = k2; overwritten below when iband = k1
#pragma unroll
                    for (int j = 0; j < K1; ++j) h[j] = sp_s[ii][j];
                    irot = number;
                    src = it;
                } else if (p <= 0.0) {
                    ++nrold;
                    continue;
                } else {
                    iband = K2;
                    const int bi = nrold - b0;
                    if (bi < kPlanChunk) {
#pragma unroll
                        for (int j = 0; j < K2; ++j) h[j] = mul(b_s[bi][j], pinv);
                    } else {  // more interior knots than grid points in this chunk: read through
#pragma unroll
                        for (int j = 0; j < K2; ++j) h[j] = mul(A.b[(size_t)nrold * kAs + j], pinv);
                    }
                    irot = nrold;
                    src = -1;
                }
                if (irot != irot_cur) advance(irot);
                step(true, h, src, iband);
                if (nrold == number) break;
                ++nrold;
            }
        }
    }
    if (!w0) return;
    advance(irot_cur + K2);  // drain and retire every band row still in the lanes
    if (lane == 0) {
        A.info[0] = s_tape;
        A.info[1] = iband;
    }
}

__global__ void __launch_bounds__(256) grid_plan_kernel(AxisDev X, AxisDev Y, double p)
{
    __shared__ double sp_s[kPlanChunk][kSp];
    __shared__ int nr_s[kPlanChunk];
    __shared__ double b_s[kPlanChunk][kAs];  // discontinuity rows the current chunk can reach
    const AxisDev &A = (blockIdx.x == 0) ? X : Y;
    switch (A.k) {
    case 1: plan_axis<1>(A, p, sp_s, nr_s, b_s); break;
    case 2: plan_axis<2>(A, p, sp_s, nr_s, b_s); break;
    case 3: plan_axis<3>(A, p, sp_s, nr_s, b_s); break;
    case 4: plan_axis<4>(A, p, sp_s, nr_s, b_s); break;
    case 5: plan_axis<5>(A, p, sp_s, nr_s, b_s); break;
    default: break;
    }
}

// Right-hand-side half of one reduction pass.  Right-hand side j: source row `s` contributes
// in[s*in_ld + j]; the reduced column goes to out[j*out_ld + row], row < nout.
// Same systolic scheme as the plan, per right-hand side: a group of 8 lanes owns one column, lane i
// holds the column's entry of band row irot+i and applies rotation i of every row (4 mul + 2 add),
// the running right-hand-side value moves one lane per step by __shfl_up.  The serial chain per
// column is one rotation per row instead of k+1 (k+2 with smoothing rows).  The source values are
// streamed through a per-column ring in shared memory filled by cp.async (kRing rows ahead), the
// rotation list is staged per chunk by the whole CTA; the pipeline drains at chunk ends and when
// the first band row advances.
constexpr int kRing = 16;
constexpr int kApplyThreads = 128, kApplyCols = kApplyThreads / 8;

__device__ __forceinline__ void cp_async8(double *dst_smem, const double *src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"((uint32_t)__cvta_generic_to_shared(dst_smem)),
                 "l"(src)
                 : "memory");
}

template <int K>
__global__ void __launch_bounds__(kApplyThreads) grid_apply_kernel(
    const double *__restrict__ in, long long in_ld, int nrhs, const int2 *__restrict__ ctl,
    const double2 *__restrict__ rcs, const int *__restrict__ info, int nsrc, double *__restrict__ out,
    long long out_ld, int nout)
{
    constexpr int K2 = K + 2;
    __shared__ int2 ctl_s[kApplyChunk + 1];
    __shared__ double2 tape_s[kApplyChunk][kCs];
    __shared__ double ring[kRing][kApplyCols];
    const int l8 = threadIdx.x & 7, col = threadIdx.x >> 3;
    const int j = blockIdx.x * kApplyCols + col;
    const bool live = j < nrhs;
    const int jj = live ? j : nrhs - 1;  // idle groups shadow the last column (no stores)
    const int nsteps = info[0];
    double *o = out + (size_t)jj * out_ld;
    const double *src = in + jj;
    const bool feeder = l8 == 0;
    const uint32_t ring_u32 = (uint32_t)__cvta_generic_to_shared(&ring[0][col]);
    if (feeder) {
#pragma unroll
        for (int u = 0; u < kRing; ++u) {  // one commit group per source row, empty past the end
            if (u < nsrc) cp_async8(&ring[u][col], src + (size_t)u * in_ld);
            asm volatile("cp.async.commit_group;\n" ::: "memory");
        }
    }
    // source value of the row entering at a step (lane 0 of the group); rows come in order
    auto fetch = [&](int srow) -> double {
        double v = 0.0;
        if (srow >= 0 && feeder) {
            asm volatile("cp.async.wait_group %0;\n" ::"n"(kRing - 1) : "memory");
            const int sl = srow & (kRing - 1);
            v = ring[sl][col];
            const int nx = srow + kRing;
            if (nx < nsrc) {
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(ring_u32 + (uint32_t)sl * (kApplyCols * 8)),
                             "l"(src + (size_t)nx * in_ld)
                             : "memory");
            }
            asm volatile("cp.async.commit_group;\n" ::: "memory");
        }
        return v;
    };
    double W = 0.0, right = 0.0;
    int irot_cur = 0;
    for (int s0 = 0; s0 < nsteps; s0 += kApplyChunk) {
        const int cnt = min(kApplyChunk, nsteps - s0);
        __syncthreads();
        for (int e = threadIdx.x; e < cnt; e += blockDim.x) ctl_s[e] = ctl[s0 + e];
        if (threadIdx.x == 0) ctl_s[cnt] = make_int2(-1, 0);
        for (int e = threadIdx.x; e < cnt * kCs; e += blockDim.x) tape_s[e / kCs][e % kCs] = rcs[(size_t)s0 * kCs + e];
        __syncthreads();
        // replay: everything a step needs was requested one step earlier
        int2 cw = ctl_s[0];
        double rnew = fetch(cw.x);
        double2 cs = tape_s[0][l8];
        for (int ss = 0; ss < cnt; ++ss) {
            const int2 cn = ctl_s[ss + 1];
            const double2 csn = tape_s[min(ss + 1, kApplyChunk - 1)][l8];
            const double rnext = fetch(cn.x);
            if (cw.y) {  // rare: the first band row advances by d: retire final rows, move the rest down
                const int d = cw.y;
                if (live && l8 < K2 && l8 < d && irot_cur + l8 < nout) o[irot_cur + l8] = W;
                const double v = __shfl_down_sync(0xffffffffu, W, (unsigned)min(d, 7), 8);
                W = (d < K2 && l8 + d < K2) ? v : 0.0;
                if (live)
                    for (int row = irot_cur + K2 + l8; row < irot_cur + d && row < nout; row += 8) o[row] = 0.0;
                irot_cur += d;
            }
            double rin = __shfl_up_sync(0xffffffffu, right, 1, 8);
            if (feeder) rin = rnew;
            if (cs.y != 0.0) rota(cs.x, cs.y, rin, W);  // (1, 0): no rotation for this lane at this step
            right = rin;
            cw = cn;
            cs = csn;
            rnew = rnext;
        }
    }
    if (feeder) asm volatile("cp.async.wait_group 0;\n" ::: "memory");
    if (live && l8 < K2 && irot_cur + l8 < nout) o[irot_cur + l8] = W;  // rows still in the lanes
}

void launch_apply(int k, const double *in, long long in_ld, int nrhs, const AxisDev &A, int nsrc,
                  double *out, long long out_ld, int nout, cudaStream_t st)
{
    const int blocks = (nrhs + kApplyCols - 1) / kApplyCols;
    switch (k) {
    case 1: grid_apply_kernel<1><<<blocks, kApplyThreads, 0, st>>>(in, in_ld, nrhs, A.ctl, A.rcs, A.info, nsrc, out, out_ld, nout); break;
    case 2: grid_apply_kernel<2><<<blocks, kApplyThreads, 0, st>>>(in, in_ld, nrhs, A.ctl, A.rcs, A.info, nsrc, out, out_ld, nout); break;
    case 3: grid_apply_kernel<3><<<blocks, kApplyThreads, 0, st>>>(in, in_ld, nrhs, A.ctl, A.rcs, A.info, nsrc, out, out_ld, nout); break;
    case 4: grid_apply_kernel<4><<<blocks, kApplyThreads, 0, st>>>(in, in_ld, nrhs, A.ctl, A.rcs, A.info, nsrc, out, out_ld, nout); break;
    case 5: grid_apply_kernel<5><<<blocks, kApplyThreads, 0, st>>>(in, in_ld, nrhs, A.ctl, A.rcs, A.info, nsrc, out, out_ld, nout); break;
    default: break;
    }
}

// fpback (core:1808-1838) on independent lines of c: element i of line j at c[j*s_line + i*s_elem]
__global__ void grid_backsub_kernel(double *__restrict__ c, int nlines, long long s_line,
                                    long long s_elem, const double *__restrict__ a,
                                    const int *__restrict__ info, int n)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nlines) return;
    const int k = info[1];
    double *z = c + (size_t)j * s_line;
    z[(size_t)(n - 1) * s_elem] = z[(size_t)(n - 1) * s_elem] / a[(size_t)(n - 1) * kAs];
    for (int i = n - 2; i >= 0; --i) {
        double store = z[(size_t)i * s_elem];
        const int i1 = min(n - 1 - i, k - 1);
        for (int l = 1; l <= i1; ++l) store = sub(store, mul(z[(size_t)(i + l) * s_elem], a[(size_t)i * kAs + l]));
        z[(size_t)i * s_elem] = store / a[(size_t)i * kAs];
    }
}

constexpr int kResRows = 32, kResCols = 128;

// Squared residuals (core:7302-7358): tile of kResRows x kResCols cells per CTA, one thread per
// column.  colpart[tile_row][i2] and rowpart[tile_col][i1] are reduced later in a fixed order.
__global__ void __launch_bounds__(kResCols) grid_resid_kernel(
    const double *__restrict__ z, int mx, int my, const double *__restrict__ c, int nk1y,
    const double *__restrict__ spx, const int *__restrict__ nrx, int k1x,
    const double *__restrict__ spy, const int *__restrict__ nry, int k1y, double *__restrict__ rowpart,
    double *__restrict__ colpart)
{
    __shared__ double tile[kResRows][kResCols + 1];
    const int tx = threadIdx.x;
    const int i2 = blockIdx.x * kResCols + tx;
    const int r0 = blockIdx.y * kResRows;
    const bool live = i2 < my;
    double wy[6];
    int numy = 0;
    if (live) {
        numy = nry[i2];
#pragma unroll
        for (int q = 0; q < 6; ++q) wy[q] = (q < k1y) ? spy[(size_t)i2 * kSp + q] : 0.0;
    }
    double colsum = 0.0;
    for (int rr = 0; rr < kResRows; ++rr) {
        const int i1 = r0 + rr;
        double term = 0.0;
        if (live && i1 < mx) {
            const int numx = nrx[i1];
            const double *cc = c + (size_t)numx * nk1y + numy;
            double sp = 0.0;
            for (int cx = 0; cx < k1x; ++cx) {
                const double pw = spx[(size_t)i1 * kSp + cx];
                double dot = 0.0;
#pragma unroll
                for (int cy = 0; cy < 6; ++cy)
                    if (cy < k1y) dot = add(dot, mul(wy[cy], __ldg(&cc[(size_t)cx * nk1y + cy])));
                sp = add(sp, mul(pw, dot));
            }
            const double d = sub(z[(size_t)i1 * my + i2], sp);
            term = mul(d, d);
        }
        tile[rr][tx] = term;
        colsum = add(colsum, term);
    }
    if (live) colpart[(size_t)blockIdx.y * my + i2] = colsum;
    __syncthreads();
    const int warp = tx >> 5, lane = tx & 31;
    for (int rr = warp; rr < kResRows; rr += kResCols / 32) {
        double s = 0.0;
#pragma unroll
        for (int q = 0; q < kResCols / 32; ++q) s = add(s, tile[rr][lane + 32 * q]);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) s = add(s, __shfl_down_sync(0xffffffffu, s, off));
        if (lane == 0 && r0 + rr < mx) rowpart[(size_t)blockIdx.x * mx + r0 + rr] = s;
    }
}

// rowsum[i1] = sum over column tiles, colsum[i2] = sum over row tiles, both in tile order
__global__ void grid_reduce_kernel(const double *__restrict__ rowpart, int ntc, int mx,
                                   const double *__restrict__ colpart, int ntr, int my,
                                   double *__restrict__ sums)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < mx) {
        double s = 0.0;
        for (int q = 0; q < ntc; ++q) s = add(s, rowpart[(size_t)q * mx + i]);
        sums[i] = s;
    } else if (i < mx + my) {
        const int i2 = i - mx;
        double s = 0.0;
        for (int q = 0; q < ntr; ++q) s = add(s, colpart[(size_t)q * my + i2]);
        sums[i] = s;
    }
}

// ---- bispev ----
// basis tables of the evaluation grids (fpndsp, core:2190-2207): arg clamped to the support
__global__ void grid_eval_basis_kernel(const double *__restrict__ xg, int m, const double *__restrict__ t,
                                       int n, int k, double *__restrict__ w, int *__restrict__ lidx)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const int k1 = k + 1, nkd = n - k1;
    const double tb = t[k1 - 1], te = t[nkd];
    double arg = xg[i];
    if (arg < tb) arg = tb;
    if (arg > te) arg = te;
    const int l = grid_interval(t, k1, nkd, arg);
    double h[6];
    switch (k) {
#define FPB_CASE(KK)                                                  \
    case KK: {                                                        \
        double tw[2 * KK], hh[KK + 1];                                \
        for (int j = 0; j < 2 * KK; ++j) tw[j] = t[l - KK + j];       \
        bspl<KK>(tw, arg, hh);                                        \
        for (int j = 0; j <= KK; ++j) h[j] = hh[j];                   \
    } break;
        FPB_CASE(1) FPB_CASE(2) FPB_CASE(3) FPB_CASE(4) FPB_CASE(5)
#undef FPB_CASE
    default: return;
    }
    for (int j = 0; j < k1; ++j) w[(size_t)i * kSp + j] = h[j];
    lidx[i] = l - k1;
}

constexpr int kEvRows = 32, kEvCols = 128;

// z(i1,i2) = sum_cx wx(cx,i1) * dot(c(lx+cx, ly : ly+ky), wy(:,i2))   (core:2219-2222 order).
// One thread per output column, kEvRows rows per CTA.  The (kx+1) x (ky+1) coefficient support of
// a thread only changes when the row's x interval changes, so it is kept in registers across rows
// (the kernel is otherwise bound by the LSU: 16 coefficient loads per 8-byte output).
// KX1/KY1 = compile-time k+1 (0: run-time degrees, up to 6 x 6 with predicates).
// FMA = false: the reference's operation order without contraction (bit-identical values); FMA = true (FAST,
// fpb_engine_set_bispev_mode): the same sums as fused multiply-adds -- half the FP64 instructions, values within
// 1e-12 * max(1, max|c|) of exact.
template <int KX1, int KY1, bool FMA>
__global__ void __launch_bounds__(kEvCols) grid_bispev_kernel(
    const double *__restrict__ c, int nk1y, const double *__restrict__ wx, const int *__restrict__ lx,
    int k1x_rt, int mx, const double *__restrict__ wyt, const int *__restrict__ ly, int k1y_rt, int my,
    double *__restrict__ z)
{
    constexpr int NX = KX1 ? KX1 : 6, NY = KY1 ? KY1 : 6;
    const int k1x = KX1 ? KX1 : k1x_rt, k1y = KY1 ? KY1 : k1y_rt;
    __shared__ double wx_s[kEvRows][kSp];
    __shared__ int lx_s[kEvRows];
    const int r0 = blockIdx.y * kEvRows;
    for (int e = threadIdx.x; e < kEvRows * kSp; e += kEvCols) {
        const int i1 = r0 + e / kSp;
        wx_s[e / kSp][e % kSp] = (i1 < mx) ? wx[(size_t)i1 * kSp + e % kSp] : 0.0;
    }
    for (int e = threadIdx.x; e < kEvRows; e += kEvCols) lx_s[e] = (r0 + e < mx) ? lx[r0 + e] : -1;
    __syncthreads();
    const int i2 = blockIdx.x * kEvCols + threadIdx.x;
    if (i2 >= my) return;
    double wy[NY];
#pragma unroll
    for (int q = 0; q < NY; ++q) wy[q] = (q < k1y) ? wyt[(size_t)i2 * kSp + q] : 0.0;
    const int l2 = ly[i2];
    double cs[NX][NY];
    int lcur = -1;
    const int nrow = min(kEvRows, mx - r0);
    for (int rr = 0; rr < nrow; ++rr) {
        const int l1 = lx_s[rr];
        if (l1 != lcur) {  // uniform across the CTA
            lcur = l1;
            const double *cc = c + (size_t)l1 * nk1y + l2;
#pragma unroll
            for (int cx = 0; cx < NX; ++cx)
#pragma unroll
                for (int cy = 0; cy < NY; ++cy)
                    cs[cx][cy] = (cx < k1x && cy < k1y) ? __ldg(&cc[(size_t)cx * nk1y + cy]) : 0.0;
        }
        double sp = 0.0;
#pragma unroll
        for (int cx = 0; cx < NX; ++cx) {
            if (cx < k1x) {
                double dot = 0.0;
#pragma unroll
                for (int cy = 0; cy < NY; ++cy)
                    if (cy < k1y) dot = FMA ? __fma_rn(cs[cx][cy], wy[cy], dot) : add(dot, mul(cs[cx][cy], wy[cy]));
                sp = FMA ? __fma_rn(wx_s[rr][cx], dot, sp) : add(sp, mul(wx_s[rr][cx], dot));
            }
        }
        z[(size_t)(r0 + rr) * my + i2] = sp;
    }
}

// ------------------------------------------------------------------------------------------
// host side: control flow of fpregr (scalars and O(nest) arrays only)
// ------------------------------------------------------------------------------------------
inline bool h_equal(double a, double b) { return std::fabs(a - b) < DBL_MIN; }  // core:18644-18647

inline int h_x86_int(double v)  // gfortran int() on x86-64 (core:12236)
{
    if (!(v > -2147483649.0 && v < 2147483648.0)) return INT_MIN;
    return (int)v;
}

// fpchec, core:2507-2570
int h_fpchec(const double *x, int m, const double *t, int n, int k)
{
    auto X = [&](int i) { return x[i - 1]; };
    auto T = [&](int i) { return t[i - 1]; };
    const int k1 = k + 1, k2 = k1 + 1, nk1 = n - k1, nk2 = nk1 + 1;
    if (nk1 < k1 || nk1 > m) return FPB_INPUT_ERROR;
    int j = n;
    for (int i = 1; i <= k; ++i) {
        if (T(i) > T(i + 1)) return FPB_INPUT_ERROR;
        if (T(j) < T(j - 1)) return FPB_INPUT_ERROR;
        --j;
    }
    for (int i = k2; i <= nk2; ++i)
        if (T(i) <= T(i - 1)) return FPB_INPUT_ERROR;
    if (X(1) < T(k1) || X(m) > T(nk2)) return FPB_INPUT_ERROR;
    if (X(1) >= T(k2) || X(m) <= T(nk1)) return FPB_INPUT_ERROR;
    int i = 1, l = k2;
    const int nk3 = nk1 - 1;
    for (j = 2; j <= nk3; ++j) {
        const double tj = T(j);
        ++l;
        const double tl = T(l);
        while (i < m && X(i) <= tj) {
            ++i;
            if (i >= m) return FPB_INPUT_ERROR;
        }
        if (X(i) >= tl) return FPB_INPUT_ERROR;
    }
    return FPB_OK;
}

// fpknot, core:8199-8275 (first eligible interval seeds the search; no eligible interval => no-op)
void h_fpknot(const double *x, double *t, int &n, double *fpint, int *nrdata, int &nrint, int istart)
{
    auto X = [&](int i) { return x[i - 1]; };
    const int k = (n - nrint - 1) / 2;
    double fpmax = 0.0;
    int number = 0, maxpt = 0, maxbeg = 0, jbegin = istart;
    for (int j = 1; j <= nrint; ++j) {
        const int jpoint = nrdata[j - 1];
        if (jpoint != 0 && (number == 0 || fpint[j - 1] > fpmax)) {
            fpmax = fpint[j - 1];
            number = j;
            maxpt = jpoint;
            maxbeg = jbegin;
        }
        jbegin = jbegin + jpoint + 1;
    }
    if (number == 0) return;
    const int ihalf = maxpt / 2 + 1, nrx = maxbeg + ihalf, next = number + 1;
    for (int j = next; j <= nrint; ++j) {
        const int jj = next + nrint - j, jk = jj + k;
        fpint[jj] = fpint[jj - 1];
        nrdata[jj] = nrdata[jj - 1];
        t[jk] = t[jk - 1];
    }
    nrdata[number - 1] = ihalf - 1;
    nrdata[next - 1] = maxpt - ihalf;
    const double am = (double)maxpt;
    double an = (double)nrdata[number - 1];
    fpint[number - 1] = fpmax * an / am;
    an = (double)nrdata[next - 1];
    fpint[next - 1] = fpmax * an / am;
    t[next + k - 1] = X(nrx);
    n = n + 1;
    nrint = nrint + 1;
}

// fprati, core:11996-12025
double h_fprati(double &p1, double &f1, double p2, double f2, double &p3, double &f3)
{
    double p;
    if (p3 > 0.0) {
        const double h1 = f1 * (f2 - f3), h2 = f2 * (f3 - f1), h3 = f3 * (f1 - f2);
        p = -(p1 * p2 * h3 + p2 * p3 * h1 + p3 * p1 * h2) / (p1 * h1 + p2 * h2 + p3 * h3);
    } else {
        p = (p1 * (f1 - f3) * f2 - p2 * (f2 - f3) * f1) / ((f1 - f2) * f3);
    }
    if (f2 >= 0.0) {
        p1 = p2;
        f1 = f2;
    } else {
        p3 = p2;
        f3 = f2;
    }
    return p;
}

// root_finding_iterate, core:11641-11691
bool h_root_iterate(double &p1, double &f1, double &p2, double &f2, double &p3, double &f3, double &p,
                    double fpms, double acc, bool &check1, bool &check3)
{
    const double con1 = 0.1, con9 = 0.9, con4 = 0.04;
    p2 = p;
    f2 = fpms;
    if (!check3) {
        if ((f2 - f3) > acc) {
            check3 = (f2 < 0.0);
        } else {
            p3 = p2;
            f3 = f2;
            p = p * con4;
            if (p <= p1) p = p1 * con9 + p2 * con1;
            return true;
        }
    }
    if (!check1) {
        if ((f1 - f2) > acc) {
            check1 = (f2 > 0.0);
        } else {
            p1 = p2;
            f1 = f2;
            p = p / con4;
            if (p3 >= 0.0 && p >= p3) p = p2 * con1 + p3 * con9;
            return true;
        }
    }
    if (f2 >= f1 || f2 <= f3) return false;
    p = h_fprati(p1, f1, p2, f2, p3, f3);
    return true;
}

// buffers of the gridded path inside the engine's scratch set
enum GridBuf { GB_X = 0, GB_Y, GB_T, GB_SP, GB_NR, GB_A, GB_B, GB_META, GB_CS, GB_INFO, GB_Q, GB_C,
               GB_PART, GB_SUMS, GB_Z, GB_COUNT };

struct GridFit {
    fpb_engine *eng;
    cudaStream_t st;
    int m[2], k[2], nest[2], k1[2], k2[2];
    const double *hx[2];
    const double *z_dev;
    AxisDev ax[2];
    double *d_t[2];
    double *d_q, *d_c, *d_rowpart, *d_colpart, *d_sums;
    int ntr, ntc;
    double *h_sums;  // pinned [mx + my]
    long long launches = 0, fpgrre_calls = 0;

    bool setup()
    {
        DevBuf *B = eng->grid;
        const int mx = m[0], my = m[1];
        const size_t mtot = (size_t)mx + my, ntot = (size_t)nest[0] + nest[1];
        const int nkx = nest[0] - k1[0], nky = nest[1] - k1[1];
        ntr = (mx + kResRows - 1) / kResRows;
        ntc = (my + kResCols - 1) / kResCols;
        bool ok = B[GB_X].reserve(mtot * 8) && B[GB_T].reserve(ntot * 8) &&
                  B[GB_SP].reserve(mtot * kSp * 8) && B[GB_NR].reserve(mtot * 4) &&
                  B[GB_A].reserve(ntot * kAs * 8) && B[GB_B].reserve(ntot * kAs * 8) &&
                  B[GB_META].reserve((mtot + 8 * ntot + 64) * sizeof(int2)) &&
                  B[GB_CS].reserve((mtot + 8 * ntot + 64) * kCs * sizeof(double2)) && B[GB_INFO].reserve(64) &&
                  B[GB_Q].reserve((size_t)my * nkx * 8) && B[GB_C].reserve((size_t)nkx * nky * 8) &&
                  B[GB_PART].reserve(((size_t)ntc * mx + (size_t)ntr * my) * 8) &&
                  B[GB_SUMS].reserve(mtot * 8);
        if (!ok) return false;
        if (eng->grid_pinned_bytes < mtot * 8) {
            if (eng->grid_pinned) cudaFreeHost(eng->grid_pinned);
            eng->grid_pinned = nullptr;
            eng->grid_pinned_bytes = 0;
            if (!cuda_ok(cudaMallocHost((void **)&eng->grid_pinned, mtot * 8), "cudaMallocHost")) return false;
            eng->grid_pinned_bytes = mtot * 8;
        }
        h_sums = eng->grid_pinned;
        double *dx = B[GB_X].as<double>();
        if (!cuda_ok(cudaMemcpyAsync(dx, hx[0], (size_t)mx * 8, cudaMemcpyHostToDevice, st), "H2D x") ||
            !cuda_ok(cudaMemcpyAsync(dx + mx, hx[1], (size_t)my * 8, cudaMemcpyHostToDevice, st), "H2D y"))
            return false;
        size_t mo = 0, no = 0;
        for (int d = 0; d < 2; ++d) {
            AxisDev &A = ax[d];
            A.x = dx + mo;
            d_t[d] = B[GB_T].as<double>() + no;
            A.t = d_t[d];
            A.sp = B[GB_SP].as<double>() + mo * kSp;
            A.nr = B[GB_NR].as<int>() + mo;
            A.a = B[GB_A].as<double>() + no * kAs;
            A.b = B[GB_B].as<double>() + no * kAs;
            A.ctl = B[GB_META].as<int2>() + (mo + 8 * no + 32 * d);
            A.rcs = B[GB_CS].as<double2>() + (mo + 8 * no + 32 * d) * kCs;
            A.info = B[GB_INFO].as<int>() + 4 * d;
            A.m = m[d];
            A.k = k[d];
            A.n = 0;
            mo += m[d];
            no += nest[d];
        }
        d_q = B[GB_Q].as<double>();
        d_c = B[GB_C].as<double>();
        d_rowpart = B[GB_PART].as<double>();
        d_colpart = d_rowpart + (size_t)ntc * mx;
        d_sums = B[GB_SUMS].as<double>();
        return true;
    }

    // one fpgrre call: least-squares / smoothing solve for the current knots and p; returns fp and
    // the per-interval residual sums of both axes
    bool fpgrre(const int *n, double *const *t, double p, double &fp, std::vector<double> *fpint,
                std::vector<int> *nr_host)
    {
        const int mx = m[0], my = m[1];
        const int nk1x = n[0] - k1[0], nk1y = n[1] - k1[1];
        for (int d = 0; d < 2; ++d) {
            ax[d].n = n[d];
            if (!cuda_ok(cudaMemcpyAsync(d_t[d], t[d], (size_t)n[d] * 8, cudaMemcpyHostToDevice, st), "H2D t"))
                return false;
        }
        grid_plan_kernel<<<2, 256, 0, st>>>(ax[0], ax[1], p);
        launch_apply(k[0], z_dev, my, my, ax[0], mx, d_q, nk1x, nk1x, st);   // pass x: my columns of z
        launch_apply(k[1], d_q, nk1x, nk1x, ax[1], my, d_c, nk1y, nk1y, st);  // pass y: nk1x rows of q
        grid_backsub_kernel<<<(nk1x + 63) / 64, 64, 0, st>>>(d_c, nk1x, nk1y, 1, ax[1].a, ax[1].info, nk1y);
        grid_backsub_kernel<<<(nk1y + 63) / 64, 64, 0, st>>>(d_c, nk1y, 1, nk1y, ax[0].a, ax[0].info, nk1x);
        grid_resid_kernel<<<dim3(ntc, ntr), kResCols, 0, st>>>(z_dev, mx, my, d_c, nk1y, ax[0].sp, ax[0].nr,
                                                               k1[0], ax[1].sp, ax[1].nr, k1[1],
                                                               d_rowpart, d_colpart);
        grid_reduce_kernel<<<(mx + my + 127) / 128, 128, 0, st>>>(d_rowpart, ntc, mx, d_colpart, ntr, my, d_sums);
        launches += 7;
        ++fpgrre_calls;
        if (!cuda_ok(cudaGetLastError(), "regrid kernels")) return false;
        if (!cuda_ok(cudaMemcpyAsync(h_sums, d_sums, ((size_t)mx + my) * 8, cudaMemcpyDeviceToHost, st), "D2H sums") ||
            !cuda_ok(cudaStreamSynchronize(st), "regrid sync"))
            return false;
        // interval numbers on the host: the same comparisons as the kernel (core:7184-7188)
        for (int d = 0; d < 2; ++d) {
            std::vector<int> &nr = nr_host[d];
            nr.resize(m[d]);
            const int nk1 = n[d] - k1[d];
            int l = k1[d], l1 = k2[d], number = 0;
            for (int it = 0; it < m[d]; ++it) {
                const double arg = hx[d][it];
                while (arg >= t[d][l1 - 1] && l != nk1) {
                    l = l1;
                    l1 = l + 1;
                    ++number;
                }
                nr[it] = number;
            }
        }
        // fp and the per-interval sums with the half split at interval changes (core:7343-7357)
        fp = 0.0;
        for (int i = 0; i < mx; ++i) fp += h_sums[i];
        for (int d = 0; d < 2; ++d) {
            std::vector<double> &f = fpint[d];
            std::fill(f.begin(), f.end(), 0.0);
            const double *s = h_sums + (d == 0 ? 0 : mx);
            int nrold = 0;
            for (int i = 0; i < m[d]; ++i) {
                const int num = nr_host[d][i];
                f[num] += s[i];
                if (num != nrold) {
                    const double fac = s[i] * 0.5;
                    f[num] -= fac;
                    f[num - 1] += fac;
                }
                nrold = num;
            }
        }
        return true;
    }
};

// fpregr + regrid's checks, dims = 2.  z_dev is the grid data in HBM (row-major, x slowest).
int32_t regrid_run(fpb_engine *eng, int iopt, int mx, const double *x, int my, const double *y,
                   const double *z_dev, double xb, double xe, double yb, double ye, int kx, int ky,
                   double s, int nxest, int nyest, int *nx, double *tx, int *ny, double *ty, double *c,
                   double *fp_out, double *wrk, int lwrk, int *iwrk, int kwrk, int *ier_out)
{
    const double tol = 1.0e-03;
    const int maxit = 20;
    int m[2] = {mx, my}, k[2] = {kx, ky}, nest[2] = {nxest, nyest};
    const double lo[2] = {xb, yb}, hi[2] = {xe, ye};
    const double *xg[2] = {x, y};
    double *tt[2] = {tx, ty};
    int k1[2], k2[2], nmin[2], nmax[2], ne[2], n[2], nrint[2], npl[2];

    *ier_out = FPB_INPUT_ERROR;
    for (int d = 0; d < 2; ++d) {
        k1[d] = k[d] + 1;
        k2[d] = k[d] + 2;
        nmin[d] = 2 * k1[d];
    }
    // ---- regrid's data checks, core:16775-16816 ----
    for (int d = 0; d < 2; ++d)
        if (k[d] <= 0 || k[d] > 5) return FPB_STATUS_OK;
    if (iopt < -1 || iopt > 1) return FPB_STATUS_OK;
    for (int d = 0; d < 2; ++d)
        if (m[d] < k1[d] || nest[d] < nmin[d]) return FPB_STATUS_OK;
    {
        const long long maxm = std::max(mx, my), maxnest = std::max(nxest, nyest);
        const long long maxk1 = std::max(kx, ky) + 1, maxk2 = maxk1 + 1;
        const long long nkx = nxest - k1[0], nky = nyest - k1[1];
        const long long mm = std::max({maxnest, maxm, (long long)my, nkx, nky});
        const long long mynx = 2 * nkx * my;
        const long long lwest = 4 + maxnest * 2 + maxm * maxk1 * 2 + 2 * maxnest * maxk2 * 2 + mm + mynx;
        const long long kwest = 3 + maxm * 2 + maxnest * 2;
        if (lwrk < lwest || kwrk < kwest) return FPB_STATUS_OK;
    }
    for (int d = 0; d < 2; ++d) {
        if (lo[d] > xg[d][0] || hi[d] < xg[d][m[d] - 1]) return FPB_STATUS_OK;
        for (int i = 0; i < m[d] - 1; ++i)
            if (xg[d][i] >= xg[d][i + 1]) return FPB_STATUS_OK;
    }
    n[0] = *nx;
    n[1] = *ny;
    if (iopt < 0) {
        for (int d = 0; d < 2; ++d) {
            if (n[d] < nmin[d] || n[d] > nest[d]) return FPB_STATUS_OK;
            int j = n[d];
            for (int i = 1; i <= k1[d]; ++i) {
                tt[d][i - 1] = lo[d];
                tt[d][j - 1] = hi[d];
                --j;
            }
            const int ie = h_fpchec(xg[d], m[d], tt[d], n[d], k[d]);
            if (ie != FPB_OK) {
                *ier_out = ie;
                return FPB_STATUS_OK;
            }
        }
    } else {
        if (s < 0.0) return FPB_STATUS_OK;
        if (h_equal(s, 0.0) && (nest[0] < m[0] + k1[0] || nest[1] < m[1] + k1[1])) return FPB_STATUS_OK;
    }

    GridFit G;
    G.eng = eng;
    G.st = eng->stream;
    for (int d = 0; d < 2; ++d) {
        G.m[d] = m[d]; G.k[d] = k[d]; G.nest[d] = nest[d]; G.k1[d] = k1[d]; G.k2[d] = k2[d];
        G.hx[d] = xg[d];
    }
    G.z_dev = z_dev;
    if (!G.setup()) return FPB_ERR_ALLOC;

    // persistent state of an iopt=1 continuation (core:16764-16766)
    double &fp0 = wrk[0], &fpold = wrk[1];
    double *reduc = wrk + 2;
    int &lastdi = iwrk[0];
    int *nplus = iwrk + 1;
    std::vector<double> fpint[2] = {std::vector<double>(nest[0] + 1, 0.0), std::vector<double>(nest[1] + 1, 0.0)};
    std::vector<int> nrdat[2] = {std::vector<int>(nest[0] + 1, 0), std::vector<int>(nest[1] + 1, 0)};
    std::vector<int> nr_host[2];

    int ier = FPB_OK;
    double fp = 0.0, fpms = 0.0, p = -1.0;
    const double acc = tol * s;
    for (int d = 0; d < 2; ++d) {
        nmax[d] = m[d] + k1[d];
        ne[d] = std::min(nmax[d], nest[d]);
    }
    auto finish = [&]() {
        *nx = n[0];
        *ny = n[1];
        for (int d = 0; d < 2; ++d)
            for (int i = n[d]; i < nest[d]; ++i) tt[d][i] = 0.0;  // fp_regrid_c zero-fills beyond n
        *fp_out = fp;
        *ier_out = ier;
        eng->launches += G.launches;
        eng->grid_fpgrre_calls = G.fpgrre_calls;
        return FPB_STATUS_OK;
    };
    auto fetch_c = [&]() -> bool {
        const size_t ncof = (size_t)(n[0] - k1[0]) * (n[1] - k1[1]);
        return cuda_ok(cudaMemcpyAsync(c, G.d_c, ncof * 8, cudaMemcpyDeviceToHost, G.st), "D2H c") &&
               cuda_ok(cudaStreamSynchronize(G.st), "sync");
    };

    if (iopt >= 0) {
        if (s <= 0.0) {  // interpolating spline, core:12112-12134
            for (int d = 0; d < 2; ++d) n[d] = nmax[d];
            if (n[0] > nest[0] || n[1] > nest[1]) {
                ier = FPB_INSUFFICIENT_STORAGE;
                return finish();
            }
            for (int d = 0; d < 2; ++d) {
                const int mk1 = m[d] - k1[d], k3 = k[d] / 2;
                int i = k1[d] + 1, j = k3 + 2;
                for (int l = 1; l <= mk1; ++l) {
                    tt[d][i - 1] = (k3 * 2 != k[d]) ? xg[d][j - 1] : (xg[d][j - 1] + xg[d][j - 2]) * 0.5;
                    ++i;
                    ++j;
                }
            }
        } else if (iopt != 0 && fp0 > s) {  // continue from the previous knots, core:12141-12157
            for (int d = 0; d < 2; ++d) {
                int l = k2[d], j = 1;
                nrdat[d][0] = 0;
                for (int i = 2; i <= m[d] - 1; ++i) {
                    nrdat[d][j - 1] += 1;
                    if (xg[d][i - 1] >= tt[d][l - 1]) {
                        nrdat[d][j - 1] -= 1;
                        ++l;
                        ++j;
                        nrdat[d][j - 1] = 0;
                    }
                }
            }
        } else {  // start from the least-squares polynomial, core:12161-12168
            for (int d = 0; d < 2; ++d) {
                n[d] = nmin[d];
                nrdat[d][0] = m[d] - 2;
                nplus[d] = 0;
                reduc[d] = 0.0;
            }
            lastdi = 0;
            fp0 = 0.0;
            fpold = 0.0;
        }
    }

    // ---- part 1: knots, core:12176-12266 ----
    const int mpm = mx + my;
    int iter = 0;
    bool accepted = false;
    while (iter <= mpm) {
        ++iter;
        if (n[0] == nmin[0] && n[1] == nmin[1]) ier = FPB_LEASTSQUARES_OK;
        for (int d = 0; d < 2; ++d) {
            nrint[d] = n[d] - nmin[d] + 1;
            int j = n[d];
            for (int i = 1; i <= k1[d]; ++i) {
                tt[d][i - 1] = lo[d];
                tt[d][j - 1] = hi[d];
                --j;
            }
        }
        if (!G.fpgrre(n, tt, p, fp, fpint, nr_host)) return FPB_ERR_CUDA;
        if (ier == FPB_LEASTSQUARES_OK) fp0 = fp;
        fpms = fp - s;
        if (iopt < 0 || std::fabs(fpms) < acc) {
            if (!fetch_c()) return FPB_ERR_CUDA;
            return finish();
        }
        if (fpms < 0.0) {
            accepted = true;
            break;
        }
        if (n[0] == nmax[0] && n[1] == nmax[1]) {
            ier = FPB_INTERPOLATING_OK;
            fp = 0.0;
            if (!fetch_c()) return FPB_ERR_CUDA;
            return finish();
        }
        if (n[0] == ne[0] && n[1] == ne[1]) {
            ier = FPB_INSUFFICIENT_STORAGE;
            if (!fetch_c()) return FPB_ERR_CUDA;
            return finish();
        }
        ier = FPB_OK;
        if (lastdi != 0) reduc[lastdi - 1] = fpold - fp;
        fpold = fp;
        for (int d = 0; d < 2; ++d) {
            npl[d] = 1;
            if (n[d] != nmin[d]) {
                int npl1 = nplus[d] * 2;
                const double rn = (double)nplus[d];
                if (reduc[d] > acc) npl1 = h_x86_int(rn * fpms / reduc[d]);
                npl[d] = std::min(nplus[d] * 2, std::max({npl1, nplus[d] / 2, 1}));
            }
        }
        // new_knot_dimension_nd, core:12347-12362: smallest pending count among the uncapped axes,
        // ties to the higher axis
        int dir = 0;
        for (int d = 1; d <= 2; ++d) {
            if (n[d - 1] >= ne[d - 1]) continue;
            if (dir == 0) dir = d;
            else if (npl[d - 1] <= npl[dir - 1]) dir = d;
        }
        lastdi = dir;
        const int d = dir - 1;
        nplus[d] = npl[d];
        for (int l = 1; l <= nplus[d]; ++l) {
            h_fpknot(xg[d], tt[d], n[d], fpint[d].data(), nrdat[d].data(), nrint[d], 1);
            if (n[d] == ne[d]) break;
        }
    }
    (void)accepted;
    if (ier == FPB_LEASTSQUARES_OK) {
        if (!fetch_c()) return FPB_ERR_CUDA;
        return finish();
    }

    // ---- part 2: smoothing parameter, core:12270-12300 ----
    double p1 = 0.0, f1 = fp0 - s, p2 = 0.0, f2 = 0.0, p3 = -1.0, f3 = fpms;
    bool check1 = false, check3 = false;
    p = 1.0;
    for (iter = 1; iter <= maxit; ++iter) {
        if (!G.fpgrre(n, tt, p, fp, fpint, nr_host)) return FPB_ERR_CUDA;
        fpms = fp - s;
        if (std::fabs(fpms) < acc) {
            if (!fetch_c()) return FPB_ERR_CUDA;
            return finish();
        }
        if (!h_root_iterate(p1, f1, p2, f2, p3, f3, p, fpms, acc, check1, check3)) {
            ier = FPB_S_TOO_SMALL;
            if (!fetch_c()) return FPB_ERR_CUDA;
            return finish();
        }
    }
    ier = FPB_MAXIT;
    if (!fetch_c()) return FPB_ERR_CUDA;
    return finish();
}

// bispev with the output grid in HBM
int32_t bispev_run(fpb_engine *eng, const double *tx, int nx, const double *ty, int ny, const double *c,
                   int kx, int ky, const double *x, int mx, const double *y, int my, double *z_dev)
{
    DevBuf *B = eng->grid;
    cudaStream_t st = eng->stream;
    const size_t mtot = (size_t)mx + my, ntot = (size_t)nx + ny;
    const int nk1x = nx - kx - 1, nk1y = ny - ky - 1;
    if (!B[GB_X].reserve(mtot * 8) || !B[GB_T].reserve(ntot * 8) || !B[GB_SP].reserve(mtot * kSp * 8) ||
        !B[GB_NR].reserve(mtot * 4) || !B[GB_C].reserve((size_t)nk1x * nk1y * 8))
        return FPB_ERR_ALLOC;
    double *dx = B[GB_X].as<double>(), *dt = B[GB_T].as<double>(), *dw = B[GB_SP].as<double>();
    int *dl = B[GB_NR].as<int>();
    double *dc = B[GB_C].as<double>();
    bool ok = cuda_ok(cudaMemcpyAsync(dx, x, (size_t)mx * 8, cudaMemcpyHostToDevice, st), "H2D") &&
              cuda_ok(cudaMemcpyAsync(dx + mx, y, (size_t)my * 8, cudaMemcpyHostToDevice, st), "H2D") &&
              cuda_ok(cudaMemcpyAsync(dt, tx, (size_t)nx * 8, cudaMemcpyHostToDevice, st), "H2D") &&
              cuda_ok(cudaMemcpyAsync(dt + nx, ty, (size_t)ny * 8, cudaMemcpyHostToDevice, st), "H2D") &&
              cuda_ok(cudaMemcpyAsync(dc, c, (size_t)nk1x * nk1y * 8, cudaMemcpyHostToDevice, st), "H2D");
    if (!ok) return FPB_ERR_CUDA;
    grid_eval_basis_kernel<<<(mx + 127) / 128, 128, 0, st>>>(dx, mx, dt, nx, kx, dw, dl);
    grid_eval_basis_kernel<<<(my + 127) / 128, 128, 0, st>>>(dx + mx, my, dt + nx, ny, ky, dw + (size_t)mx * kSp, dl + mx);
    cudaEventRecord(eng->ev0, st);
    {
        const dim3 gridd((my + kEvCols - 1) / kEvCols, (mx + kEvRows - 1) / kEvRows);
        const double *dwy = dw + (size_t)mx * kSp;
        const int *dly = dl + mx;
#define FPB_BISPEV(A, B)                                                                                              \
    do {                                                                                                              \
        if (eng->bispev_fast)                                                                                         \
            grid_bispev_kernel<A, B, true><<<gridd, kEvCols, 0, st>>>(dc, nk1y, dw, dl, kx + 1, mx, dwy, dly, ky + 1, my, z_dev);  \
        else                                                                                                          \
            grid_bispev_kernel<A, B, false><<<gridd, kEvCols, 0, st>>>(dc, nk1y, dw, dl, kx + 1, mx, dwy, dly, ky + 1, my, z_dev); \
    } while (0)
        if (kx == ky && kx == 1) FPB_BISPEV(2, 2);
        else if (kx == ky && kx == 2) FPB_BISPEV(3, 3);
        else if (kx == ky && kx == 3) FPB_BISPEV(4, 4);
        else if (kx == ky && kx == 4) FPB_BISPEV(5, 5);
        else if (kx == ky && kx == 5) FPB_BISPEV(6, 6);
        else FPB_BISPEV(0, 0);
#undef FPB_BISPEV
    }
    cudaEventRecord(eng->ev1, st);
    eng->timed = true;
    eng->launches += 3;
    if (!cuda_ok(cudaGetLastError(), "bispev kernels")) return FPB_ERR_CUDA;
    return FPB_STATUS_OK;
}

struct DevGuard {
    int prev = -1;
    bool ok = true;
    explicit DevGuard(int dev)
    {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != dev) ok = cuda_ok(cudaSetDevice(dev), "cudaSetDevice");
    }
    ~DevGuard()
    {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

}  // namespace

}  // namespace fpb

using namespace fpb;

extern "C" {

int32_t fp_regrid_dev_c(fpb_engine *eng, FP_SIZE iopt, FP_SIZE mx, const FP_REAL *x, FP_SIZE my,
                        const FP_REAL *y, const FP_REAL *z_dev, FP_REAL xb, FP_REAL xe, FP_REAL yb,
                        FP_REAL ye, FP_SIZE kx, FP_SIZE ky, FP_REAL s, FP_SIZE nxest, FP_SIZE nyest,
                        FP_SIZE *nx, FP_REAL *tx, FP_SIZE *ny, FP_REAL *ty, FP_REAL *c, FP_REAL *fp,
                        FP_REAL *wrk, FP_SIZE lwrk, FP_SIZE *iwrk, FP_SIZE kwrk, FP_FLAG *ier)
{
    if (!eng || !x || !y || !z_dev || !nx || !tx || !ny || !ty || !c || !fp || !wrk || !iwrk || !ier) {
        set_last_error("fp_regrid_dev_c: null required pointer");
        return FPB_ERR_ARG;
    }
    DevGuard g(eng->device);
    if (!g.ok) return FPB_ERR_CUDA;
    cudaEventRecord(eng->ev0, eng->stream);
    const int32_t rc = regrid_run(eng, iopt, mx, x, my, y, z_dev, xb, xe, yb, ye, kx, ky, s, nxest, nyest,
                                  nx, tx, ny, ty, c, fp, wrk, lwrk, iwrk, kwrk, ier);
    cudaEventRecord(eng->ev1, eng->stream);
    eng->timed = true;
    return rc;
}

void fp_regrid_c(FP_SIZE iopt, FP_SIZE mx, const FP_REAL *x, FP_SIZE my, const FP_REAL *y,
                 const FP_REAL *z, FP_REAL xb, FP_REAL xe, FP_REAL yb, FP_REAL ye, FP_SIZE kx, FP_SIZE ky,
                 FP_REAL s, FP_SIZE nxest, FP_SIZE nyest, FP_SIZE *nx, FP_REAL *tx, FP_SIZE *ny,
                 FP_REAL *ty, FP_REAL *c, FP_REAL *fp, FP_REAL *wrk, FP_SIZE lwrk, FP_SIZE *iwrk,
                 FP_SIZE kwrk, FP_FLAG *ier)
{
    *ier = FPB_INPUT_ERROR;
    if (mx < 1 || my < 1) return;
    HostApiLock api_lock(-1);  // per-device serialisation (fpb_engine.h)
    fpb_engine *eng = default_engine(-1);
    if (!eng) {
        *ier = FPB_ERR_CUDA;
        return;
    }
    DevGuard g(eng->device);
    DevBuf &Z = eng->grid[GB_Z];
    const size_t bytes = (size_t)mx * my * 8;
    if (!Z.reserve(bytes)) {
        *ier = FPB_ERR_ALLOC;
        return;
    }
    if (!cuda_ok(cudaMemcpyAsync(Z.ptr, z, bytes, cudaMemcpyHostToDevice, eng->stream), "H2D z")) {
        *ier = FPB_ERR_CUDA;
        return;
    }
    const int32_t rc = fp_regrid_dev_c(eng, iopt, mx, x, my, y, Z.as<double>(), xb, xe, yb, ye, kx, ky, s,
                                       nxest, nyest, nx, tx, ny, ty, c, fp, wrk, lwrk, iwrk, kwrk, ier);
    if (rc != FPB_STATUS_OK) *ier = rc;
}

int32_t fp_bispev_dev_c(fpb_engine *eng, const FP_REAL *tx, FP_SIZE nx, const FP_REAL *ty, FP_SIZE ny,
                        const FP_REAL *c, FP_SIZE kx, FP_SIZE ky, const FP_REAL *x, FP_SIZE mx,
                        const FP_REAL *y, FP_SIZE my, FP_REAL *z_dev, FP_FLAG *ier)
{
    if (!eng || !tx || !ty || !c || !x || !y || !z_dev || !ier) {
        set_last_error("fp_bispev_dev_c: null required pointer");
        return FPB_ERR_ARG;
    }
    *ier = FPB_INPUT_ERROR;
    // bispev's checks, core:447-452
    if (mx < 1 || my < 1) return FPB_STATUS_OK;
    if (kx < 1 || kx > 5 || ky < 1 || ky > 5 || nx < 2 * (kx + 1) || ny < 2 * (ky + 1)) return FPB_STATUS_OK;
    for (int i = 1; i < mx; ++i)
        if (x[i] < x[i - 1]) return FPB_STATUS_OK;
    for (int i = 1; i < my; ++i)
        if (y[i] < y[i - 1]) return FPB_STATUS_OK;
    DevGuard g(eng->device);
    if (!g.ok) return FPB_ERR_CUDA;
    const int32_t rc = bispev_run(eng, tx, nx, ty, ny, c, kx, ky, x, mx, y, my, z_dev);
    if (rc == FPB_STATUS_OK) *ier = FPB_OK;
    return rc;
}

FP_FLAG fp_bispev_c(const FP_REAL *tx, FP_SIZE nx, const FP_REAL *ty, FP_SIZE ny, const FP_REAL *c,
                    FP_SIZE kx, FP_SIZE ky, const FP_REAL *x, FP_SIZE mx, const FP_REAL *y, FP_SIZE my,
                    FP_REAL *z, FP_REAL *wrk, FP_SIZE lwrk, FP_SIZE *iwrk, FP_SIZE kwrk)
{
    (void)wrk;
    (void)iwrk;
    const long long lwest = (long long)(kx + 1) * mx + (long long)(ky + 1) * my;
    if (lwrk < lwest || kwrk < (mx + my) || mx < 1 || my < 1) return FPB_INPUT_ERROR;
    HostApiLock api_lock(-1);  // per-device serialisation (fpb_engine.h)
    fpb_engine *eng = default_engine(-1);
    if (!eng) return FPB_ERR_CUDA;
    DevGuard g(eng->device);
    DevBuf &Z = eng->grid[GB_Z];
    const size_t bytes = (size_t)mx * my * 8;
    if (!Z.reserve(bytes)) return FPB_ERR_ALLOC;
    FP_FLAG ier = FPB_INPUT_ERROR;
    const int32_t rc = fp_bispev_dev_c(eng, tx, nx, ty, ny, c, kx, ky, x, mx, y, my, Z.as<double>(), &ier);
    if (rc != FPB_STATUS_OK) return rc;
    if (ier != FPB_OK) return ier;
    if (!cuda_ok(cudaMemcpyAsync(z, Z.ptr, bytes, cudaMemcpyDeviceToHost, eng->stream), "D2H z") ||
        !cuda_ok(cudaStreamSynchronize(eng->stream), "sync"))
        return FPB_ERR_CUDA;
    return FPB_OK;
}

int64_t fpb_engine_grid_fpgrre_calls(const fpb_engine *eng) { return eng ? eng->grid_fpgrre_calls : 0; }

}  // extern "C"
