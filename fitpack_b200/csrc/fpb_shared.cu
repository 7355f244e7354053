// fpb_shared.cu -- shared-abscissa least-squares fast path (curfit iopt=-1 with
// one x / w / knot set for the whole batch; BASELINE config 3).
//
// Reference: curfit core:1240-1299 (iopt=-1 branch, fpchec 2507-2570) and the
// LSQ pass of fpcurf core:4859-4906 (fpbspl, fp_rotate_row, fpgivs, fprota,
// fpback).  In that pass the Givens parameters depend only on (x, w, t): every
// curve of the batch would regenerate the SAME cos/sin.  So:
//   1. `shared_plan_kernel` runs the matrix half of the pass ONCE (the reference's
//      arithmetic, sequential over the m points) and records, per point, the band
//      row it starts at, the k+1 (cos, sin) pairs and which pivots were skipped;
//      it also leaves the final triangular band factor R.
//   2. `shared_apply_kernel` (the hot kernel, one thread per curve) streams y
//      point-major with coalesced loads and applies the recorded rotations to the
//      curve's own right-hand side only: 4 mul + 2 add per rotation, no sqrt, no
//      division.  The plan is staged tile by tile through shared memory and
//      read by all lanes as a broadcast; the k+1 live right-hand-side entries sit
//      in registers and retire to a shared-memory column as the band window
//      advances (uniform control flow: the window schedule is shared).
//   3. back-substitution with the shared R, then a transposed, coalesced store
//      of c (curve-major, as the caller holds it).
// Results are bit-identical to running curfit(iopt=-1) per curve: the same IEEE
// operations in the same order, no FMA.
#include "fpb_common.cuh"
#include "fpb_kernels.h"

namespace fpb {

namespace {

constexpr int kApplyThreads = 128;
constexpr int kPlanChunk = 64;  // points per staged plan tile

template <int K>
__global__ void shared_plan_kernel(const double *x, const double *w, int m, double xb, double xe,
                                   int n, double *t, SharedPlan plan)
{
    constexpr int K1 = K + 1, K2 = K + 2;
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const int nmin = 2 * K1;
    auto X = [&](int i) { return x[i - 1]; };
    auto Tt = [&](int i) -> double & { return t[i - 1]; };
    // ---- curfit checks, core:1265-1285 ----
    int ier = IER_INPUT;
    *plan.ier = ier;
    if (m < K1) return;
    if (xb > X(1) || xe < X(m)) return;
    for (int i = 1; i < m; ++i)
        if (X(i) > X(i + 1)) return;
    if (n < nmin) return;
    {
        int j = n;
        for (int i = 1; i <= K1; ++i) {
            Tt(i) = xb;
            Tt(j) = xe;
            --j;
        }
    }
    // ---- fpchec, core:2507-2570 ----
    const int nk1 = n - K1, nk2 = nk1 + 1;
    if (nk1 < K1 || nk1 > m) return;
    {
        int j = n;
        for (int i = 1; i <= K; ++i) {
            if (Tt(i) > Tt(i + 1)) return;
            if (Tt(j) < Tt(j - 1)) return;
            --j;
        }
        for (int i = K2; i <= nk2; ++i)
            if (Tt(i) <= Tt(i - 1)) return;
        if (X(1) < Tt(K1) || X(m) > Tt(nk2)) return;
        if (X(1) >= Tt(K2) || X(m) <= Tt(nk1)) return;
        int i = 1, l = K2;
        const int nk3 = nk1 - 1;
        for (j = 2; j <= nk3; ++j) {
            const double tj = Tt(j);
            ++l;
            const double tl = Tt(l);
            while (i < m && X(i) <= tj) {
                ++i;
                if (i >= m) return;
            }
            if (X(i) >= tl) return;
        }
    }
    // ---- matrix half of the LSQ pass, core:4867-4896 ----
    double *R = plan.R;
    for (int i = 0; i < nk1 * K1; ++i) R[i] = 0.0;
    int l = K1;
    for (int it = 1; it <= m; ++it) {
        const double xi = X(it);
        const double wi = w ? w[it - 1] : 1.0;
        while (l < nk1 && xi >= Tt(l + 1)) ++l;
        double tw[2 * K], h[K1];
#pragma unroll
        for (int j = 0; j < 2 * K; ++j) tw[j] = Tt(l - K + 1 + j);
        bspl<K>(tw, xi, h);
#pragma unroll
        for (int j = 0; j < K1; ++j) h[j] = mul(wi, h[j]);
        int skip = 0;
        const int row0 = l - K1;  // 0-based first band row
#pragma unroll
        for (int i = 0; i < K1; ++i) {
            const double piv = h[i];
            double cs = 1.0, sn = 0.0;
            if (fp_is_zero(piv)) {
                skip |= (1 << i);
            } else {
                double *Rrow = R + (size_t)(row0 + i) * K1;
                givens(piv, Rrow[0], cs, sn);
#pragma unroll
                for (int q = 1; q <= K - i; ++q) rota(cs, sn, h[i + q], Rrow[q]);
            }
            plan.cs_sn[((size_t)(it - 1) * K1 + i) * 2] = cs;
            plan.cs_sn[((size_t)(it - 1) * K1 + i) * 2 + 1] = sn;
        }
        plan.wgt[it - 1] = wi;
        plan.lrow[it - 1] = row0;
        plan.skip[it - 1] = skip;
    }
    // curfit reports the polynomial case (no interior knots) as ier=-2 (core:4852)
    *plan.ier = (n == nmin) ? IER_LSQ : IER_OK;
}

// meta word of a point: (band rows to retire before it) << 8 | skipped-pivot bits
__global__ void shared_meta_kernel(int m, const int *lrow, const int *skip, int *meta)
{
    for (int it = blockIdx.x * blockDim.x + threadIdx.x; it < m; it += gridDim.x * blockDim.x) {
        const int adv = lrow[it] - (it > 0 ? lrow[it - 1] : 0);
        meta[it] = (adv << 8) | (skip[it] & 0xff);
    }
}

template <int K, bool HAS_W>
__global__ void __launch_bounds__(kApplyThreads)
    shared_apply_kernel(int nbatch, int m, int n, const double *__restrict__ y, long long y_sb,
                        long long y_si, SharedPlan plan, const int *__restrict__ meta_g,
                        double *__restrict__ c, long long ldc, double *__restrict__ fp_out)
{
    constexpr int K1 = K + 1;
    constexpr int ZLD = kApplyThreads + 1;
    extern __shared__ __align__(16) double sm[];
    const int nk1 = n - K1;
    double *zs = sm;                                           // [nk1][ZLD]
    double *Rs = zs + (((size_t)nk1 * ZLD + 1) & ~(size_t)1);  // [nk1][K1]   (16 B aligned)
    double2 *pcs = reinterpret_cast<double2 *>(Rs + (((size_t)nk1 * K1 + 1) & ~(size_t)1));
    double *pw = reinterpret_cast<double *>(pcs + (size_t)kPlanChunk * K1);  // [kPlanChunk]
    int *pmeta = reinterpret_cast<int *>(pw + kPlanChunk);                   // [kPlanChunk]
    const int tid = threadIdx.x;
    if (*plan.ier > 0) return;  // ier<=0: success codes (0, or -2 for the polynomial)
    for (int i = tid; i < nk1 * K1; i += kApplyThreads) Rs[i] = plan.R[i];
    const double2 *cs_g = reinterpret_cast<const double2 *>(plan.cs_sn);

    const long long ntiles = ((long long)nbatch + kApplyThreads - 1) / kApplyThreads;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long b = tile * kApplyThreads + tid;
        const bool valid = b < nbatch;
        const double *yp = y + (valid ? b : 0) * y_sb;
        double Zw[K1];
#pragma unroll
        for (int r = 0; r < K1; ++r) Zw[r] = 0.0;
        double fp = 0.0;
        int cur = 0;  // band row held in Zw[0]
        // y is streamed through a register double buffer: the 8 loads of the next group are in
        // flight while the current group is rotated in (independent of the plan chunking)
        double nxt[8];
        const double *ynext = yp;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            nxt[u] = (u < m) ? *ynext : 0.0;
            ynext += y_si;
        }
        for (int base = 0; base < m; base += kPlanChunk) {
            const int cnt = min(kPlanChunk, m - base);
            __syncthreads();  // previous chunk fully consumed
            for (int i = tid; i < cnt * K1; i += kApplyThreads) pcs[i] = cs_g[(size_t)base * K1 + i];
            for (int i = tid; i < cnt; i += kApplyThreads) {
                if (HAS_W) pw[i] = plan.wgt[base + i];
                pmeta[i] = meta_g[base + i];
            }
            __syncthreads();
            for (int q0 = 0; q0 < cnt; q0 += 8) {
                double yv[8];
                const int itn = base + q0 + 8;  // first point of the next group
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    yv[u] = nxt[u];
                    nxt[u] = (itn + u < m) ? *ynext : 0.0;
                    ynext += y_si;
                }
                const double2 *pc = pcs + q0 * K1;
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int q = q0 + u;
                    if (q < cnt) {  // uniform
                        const int meta = pmeta[q];
                        if (meta >> 8) {  // the band window moves: retire finished rows (uniform)
                            for (int a = meta >> 8; a > 0; --a) {
                                zs[(size_t)cur * ZLD + tid] = Zw[0];
#pragma unroll
                                for (int r = 0; r < K; ++r) Zw[r] = Zw[r + 1];
                                Zw[K] = 0.0;
                                ++cur;
                            }
                        }
                        double yi = HAS_W ? mul(yv[u], pw[q]) : yv[u];
                        if ((meta & 0xff) == 0) {
#pragma unroll
                            for (int i = 0; i < K1; ++i) {
                                const double2 r2 = pc[u * K1 + i];
                                rota(r2.x, r2.y, yi, Zw[i]);
                            }
                        } else {
#pragma unroll
                            for (int i = 0; i < K1; ++i) {
                                if (!((meta >> i) & 1)) {
                                    const double2 r2 = pc[u * K1 + i];
                                    rota(r2.x, r2.y, yi, Zw[i]);
                                }
                            }
                        }
                        fp = add(fp, mul(yi, yi));
                    }
                }
            }
        }
        while (cur < nk1 - K1) {
            zs[(size_t)cur * ZLD + tid] = Zw[0];
#pragma unroll
            for (int r = 0; r < K; ++r) Zw[r] = Zw[r + 1];
            Zw[K] = 0.0;
            ++cur;
        }
#pragma unroll
        for (int r = 0; r < K1; ++r) zs[(size_t)(cur + r) * ZLD + tid] = Zw[r];
        // fpback, core:1808-1838 (own column of zs only: no barrier needed)
        double cw[K];
#pragma unroll
        for (int q = 0; q < K; ++q) cw[q] = 0.0;
        for (int i = nk1 - 1; i >= 0; --i) {
            double store = zs[(size_t)i * ZLD + tid];
            const int cnt = min(nk1 - 1 - i, K);
#pragma unroll
            for (int q = 1; q <= K; ++q)
                if (q <= cnt) store = sub(store, mul(cw[q - 1], Rs[i * K1 + q]));
            const double ci = store / Rs[i * K1];
            zs[(size_t)i * ZLD + tid] = ci;
#pragma unroll
            for (int q = K - 1; q >= 1; --q) cw[q] = cw[q - 1];
            cw[0] = ci;
        }
        if (valid) fp_out[b] = fp;
        __syncthreads();
        // transposed store: each warp writes its 32 curves, one curve per step, lanes along i
        {
            const int warp = tid >> 5, lane = tid & 31;
            for (int cb = 0; cb < 32; ++cb) {
                const long long bc = tile * kApplyThreads + warp * 32 + cb;
                if (bc >= nbatch) break;
                for (int i = lane; i < nk1; i += 32)
                    c[bc * ldc + i] = zs[(size_t)i * ZLD + warp * 32 + cb];
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Two curves per thread (the fast shape: point-major y, unit weights, 16-byte aligned pairs).
// The kernel above is issue bound (ncu: issue slots 61 %, FP64 pipe 53 %, 62 instructions per
// point of which 26 are FP64): every (cos, sin) broadcast load, meta word, branch and pointer
// update is paid per curve.  Here a thread owns curves 2j and 2j+1: one LDG.128 fetches both
// ordinates, one LDS.128 feeds two rotations, the bookkeeping is shared, and the two dependent
// rotation chains interleave.  FMA = false: the reference's operations (4 mul + 2 add per
// rotation, bit-identical to curfit(iopt=-1)); FMA = true ("fast" mode): 2 mul + 2 fma, results
// within rounding of the exact mode (tests state 1e-12 relative to max|c|) -- there is no discrete
// decision downstream of these rotations in an iopt=-1 fit.
// ---------------------------------------------------------------------------------------------
constexpr int kApply2Threads = 64;              // 128 curves per CTA tile
constexpr int kApply2Cols = 2 * kApply2Threads;

template <bool FMA>
__device__ __forceinline__ void rota2(double cs, double sn, double &a, double &b)
{
    if (FMA) {
        const double s1 = a, s2 = b;
        b = fma(cs, s2, mul(sn, s1));
        a = fma(cs, s1, -mul(sn, s2));
    } else {
        rota(cs, sn, a, b);
    }
}

// Retired right-hand-side rows do not stay in shared memory (a [nk1][curves] column block would
// cap the SM at ~640 resident curves = 10 warps): they collect in a 4-row buffer that the CTA
// flushes to the output array c as full 32-byte sectors (4 doubles per curve); the final
// back-substitution then runs in place on c (each thread re-reads its own two rows, L2 resident).
constexpr int kZRows = 4;

template <int K, bool FMA>
__global__ void __launch_bounds__(kApply2Threads)
    shared_apply2_kernel(int nbatch, int m, int n, const double *__restrict__ y, long long y_si,
                         SharedPlan plan, const int *__restrict__ meta_g, double *__restrict__ c,
                         long long ldc, double *__restrict__ fp_out)
{
    constexpr int K1 = K + 1;
    constexpr int ZLD = kApply2Cols + 1;
    constexpr int D = 8;  // points requested ahead per thread (16 B each)
    extern __shared__ __align__(16) double sm[];
    const int nk1 = n - K1;
    double *zb = sm;                                             // [kZRows][ZLD]
    double *Rs = zb + (((size_t)kZRows * ZLD + 1) & ~(size_t)1);  // [nk1][K1]   (16 B aligned)
    double2 *pcs = reinterpret_cast<double2 *>(Rs + (((size_t)nk1 * K1 + 1) & ~(size_t)1));
    int *pmeta = reinterpret_cast<int *>(pcs + (size_t)kPlanChunk * K1);  // [kPlanChunk]
    int *pgrp = pmeta + kPlanChunk;                                       // [kPlanChunk / 4]
    const int tid = threadIdx.x;
    if (*plan.ier > 0) return;
    for (int i = tid; i < nk1 * K1; i += kApply2Threads) Rs[i] = plan.R[i];
    const double2 *cs_g = reinterpret_cast<const double2 *>(plan.cs_sn);
    const long long ntiles = ((long long)nbatch + kApply2Cols - 1) / kApply2Cols;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long b0 = tile * kApply2Cols + 2 * tid;  // curves b0, b0+1 (nbatch is even)
        const bool valid = b0 < nbatch;
        const double2 *yp = reinterpret_cast<const double2 *>(y + (valid ? b0 : 0));
        const long long ystep = y_si / 2;  // in double2 units
        double ZA[K1], ZB[K1];
#pragma unroll
        for (int r = 0; r < K1; ++r) ZA[r] = ZB[r] = 0.0;
        double fpA = 0.0, fpB = 0.0;
        int cur = 0;  // band row held in Z?[0]
        // rows [row0, row0+cntz) of the buffer go to c[curve*ldc + row]; local curve lc sits in
        // column (lc & 1) * 64 + (lc >> 1) (thread lc>>1, half lc&1)
        auto flush = [&](int row0, int cntz) {
            __syncthreads();
            for (int e = tid; e < kApply2Cols * cntz; e += kApply2Threads) {
                const int lc = e / cntz, jr = e % cntz;
                const long long bc = tile * kApply2Cols + lc;
                if (bc < nbatch)
                    c[bc * ldc + row0 + jr] = zb[(size_t)jr * ZLD + (lc & 1) * kApply2Threads + (lc >> 1)];
            }
            __syncthreads();
        };
        auto retire = [&]() {  // uniform across the CTA
            zb[(size_t)(cur % kZRows) * ZLD + tid] = ZA[0];
            zb[(size_t)(cur % kZRows) * ZLD + kApply2Threads + tid] = ZB[0];
#pragma unroll
            for (int r = 0; r < K; ++r) {
                ZA[r] = ZA[r + 1];
                ZB[r] = ZB[r + 1];
            }
            ZA[K] = ZB[K] = 0.0;
            ++cur;
            if (cur % kZRows == 0) flush(cur - kZRows, kZRows);
        };
        double2 nxt[D];
        const double2 *ynext = yp;
#pragma unroll
        for (int u = 0; u < D; ++u) {
            nxt[u] = (u < m) ? *ynext : make_double2(0.0, 0.0);
            ynext += ystep;
        }
        for (int base = 0; base < m; base += kPlanChunk) {
            const int cnt = min(kPlanChunk, m - base);
            __syncthreads();  // previous chunk fully consumed
            for (int i = tid; i < cnt * K1; i += kApply2Threads) pcs[i] = cs_g[(size_t)base * K1 + i];
            for (int i = tid; i < cnt; i += kApply2Threads) pmeta[i] = meta_g[base + i];
            for (int i = tid; i < kPlanChunk / 4; i += kApply2Threads) {
                int orv = 0;
                for (int q = 4 * i; q < 4 * i + 4; ++q) orv |= (q < cnt) ? meta_g[base + q] : 1;
                pgrp[i] = orv;
            }
            __syncthreads();
            for (int q0 = 0; q0 < cnt; q0 += D) {
                double2 yv[D];
                const int itn = base + q0 + D;  // first point of the next group
#pragma unroll
                for (int u = 0; u < D; ++u) {
                    yv[u] = nxt[u];
                    nxt[u] = (itn + u < m) ? *ynext : make_double2(0.0, 0.0);
                    ynext += ystep;
                }
                const double2 *pc = pcs + q0 * K1;
                // groups of 4 points: when none of them moves the band window or skips a pivot
                // (the common case: pgrp == 0) the 16 rotations run as straight-line code, so the
                // (cos, sin) loads and the two curves' chains of neighbouring points interleave
#pragma unroll
                for (int h4 = 0; h4 < D / 4; ++h4) {
                    const int qh = q0 + 4 * h4;
                    if (qh + 4 <= cnt && pgrp[qh >> 2] == 0) {
#pragma unroll
                        for (int u = 4 * h4; u < 4 * h4 + 4; ++u) {
                            double ya = yv[u].x, yb = yv[u].y;
#pragma unroll
                            for (int i = 0; i < K1; ++i) {
                                const double2 r2 = pc[u * K1 + i];
                                rota2<FMA>(r2.x, r2.y, ya, ZA[i]);
                                rota2<FMA>(r2.x, r2.y, yb, ZB[i]);
                            }
                            if (FMA) {
                                fpA = fma(ya, ya, fpA);
                                fpB = fma(yb, yb, fpB);
                            } else {
                                fpA = add(fpA, mul(ya, ya));
                                fpB = add(fpB, mul(yb, yb));
                            }
                        }
                        continue;
                    }
#pragma unroll
                    for (int u = 4 * h4; u < 4 * h4 + 4; ++u) {
                        const int q = q0 + u;
                        if (q < cnt) {  // uniform
                            const int meta = pmeta[q];
                            if (meta >> 8) {  // the band window moves: retire finished rows (uniform)
                                for (int a = meta >> 8; a > 0; --a) retire();
                            }
                            double ya = yv[u].x, yb = yv[u].y;
#pragma unroll
                            for (int i = 0; i < K1; ++i) {
                                if (!((meta >> i) & 1)) {
                                    const double2 r2 = pc[u * K1 + i];
                                    rota2<FMA>(r2.x, r2.y, ya, ZA[i]);
                                    rota2<FMA>(r2.x, r2.y, yb, ZB[i]);
                                }
                            }
                            if (FMA) {
                                fpA = fma(ya, ya, fpA);
                                fpB = fma(yb, yb, fpB);
                            } else {
                                fpA = add(fpA, mul(ya, ya));
                                fpB = add(fpB, mul(yb, yb));
                            }
                        }
                    }
                }
            }
        }
        // the rows still in the register window (and skipped trailing rows) retire too
        while (cur < nk1) retire();
        if (cur % kZRows) flush(cur - cur % kZRows, cur % kZRows);
        // fpback, core:1808-1838, in place on my two rows of c
        if (valid) {
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                double *cz = c + (b0 + half) * ldc;
                double cw[K];
#pragma unroll
                for (int q = 0; q < K; ++q) cw[q] = 0.0;
                for (int i = nk1 - 1; i >= 0; --i) {
                    double store = cz[i];
                    const int cn = min(nk1 - 1 - i, K);
#pragma unroll
                    for (int q = 1; q <= K; ++q)
                        if (q <= cn) store = sub(store, mul(cw[q - 1], Rs[i * K1 + q]));
                    const double ci = store / Rs[i * K1];
                    cz[i] = ci;
#pragma unroll
                    for (int q = K - 1; q >= 1; --q) cw[q] = cw[q - 1];
                    cw[0] = ci;
                }
            }
            fp_out[b0] = fpA;
            fp_out[b0 + 1] = fpB;
        }
    }
}

}  // namespace

cudaError_t launch_shared_plan(const double *x, const double *w, int m, double xb, double xe, int k,
                               int n, double *t, const SharedPlan &plan, cudaStream_t stream)
{
    switch (k) {
    case 1: shared_plan_kernel<1><<<1, 32, 0, stream>>>(x, w, m, xb, xe, n, t, plan); break;
    case 2: shared_plan_kernel<2><<<1, 32, 0, stream>>>(x, w, m, xb, xe, n, t, plan); break;
    case 3: shared_plan_kernel<3><<<1, 32, 0, stream>>>(x, w, m, xb, xe, n, t, plan); break;
    case 4: shared_plan_kernel<4><<<1, 32, 0, stream>>>(x, w, m, xb, xe, n, t, plan); break;
    case 5: shared_plan_kernel<5><<<1, 32, 0, stream>>>(x, w, m, xb, xe, n, t, plan); break;
    default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

cudaError_t launch_shared_apply(int nbatch, int m, int k, int n, const double *y, long long y_sb,
                                long long y_si, const SharedPlan &plan, const double *w, int *meta,
                                double *c, long long ldc, double *fp, int nsm, cudaStream_t stream,
                                int fast)
{
    if (nbatch <= 0) return cudaSuccess;
    const int k1 = k + 1, nk1 = n - k1;
    shared_meta_kernel<<<(m + 255) / 256, 256, 0, stream>>>(m, plan.lrow, plan.skip, meta);
    // two-curves-per-thread shape: point-major y, unit weights, pairs of curves 16-byte aligned
    if (!w && y_sb == 1 && (nbatch & 1) == 0 && (y_si & 1) == 0 && ((uintptr_t)y & 15) == 0) {
        const size_t smem2 = sizeof(double) * ((((size_t)kZRows * (kApply2Cols + 1) + 1) & ~(size_t)1) +
                                               (((size_t)nk1 * k1 + 1) & ~(size_t)1) +
                                               (size_t)kPlanChunk * 2 * k1) +
                             sizeof(int) * (kPlanChunk + kPlanChunk / 4);
        if (smem2 <= 200 * 1024) {
            const long long tiles2 = ((long long)nbatch + kApply2Cols - 1) / kApply2Cols;
            long long res2 = (long long)((228 * 1024) / (smem2 + 1024));
            if (res2 > 8) res2 = 8;
            if (res2 < 1) res2 = 1;
            const long long grid2 = tiles2 < nsm * res2 ? tiles2 : nsm * res2;
#define FPB_LAUNCH_APPLY2(KK)                                                                      \
    do {                                                                                           \
        auto kern = fast ? shared_apply2_kernel<KK, true> : shared_apply2_kernel<KK, false>;       \
        cudaError_t e_ = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,   \
                                              (int)smem2);                                         \
        if (e_ != cudaSuccess) return e_;                                                          \
        kern<<<(unsigned)grid2, kApply2Threads, smem2, stream>>>(nbatch, m, n, y, y_si, plan, meta, \
                                                                 c, ldc, fp);                      \
    } while (0)
            switch (k) {
            case 1: FPB_LAUNCH_APPLY2(1); break;
            case 2: FPB_LAUNCH_APPLY2(2); break;
            case 3: FPB_LAUNCH_APPLY2(3); break;
            case 4: FPB_LAUNCH_APPLY2(4); break;
            case 5: FPB_LAUNCH_APPLY2(5); break;
            default: return cudaErrorInvalidValue;
            }
#undef FPB_LAUNCH_APPLY2
            return cudaGetLastError();
        }
    }
    const size_t smem = sizeof(double) * ((((size_t)nk1 * (kApplyThreads + 1) + 1) & ~(size_t)1) +
                                          (((size_t)nk1 * k1 + 1) & ~(size_t)1) +
                                          (size_t)kPlanChunk * 2 * k1 + kPlanChunk) +
                        sizeof(int) * kPlanChunk;
    if (smem > 220 * 1024) return cudaErrorInvalidValue;
    long long tiles = ((long long)nbatch + kApplyThreads - 1) / kApplyThreads;
    long long resident = (long long)((228 * 1024) / (smem + 1024));
    if (resident > 8) resident = 8;
    if (resident < 1) resident = 1;
    long long grid = tiles < nsm * resident ? tiles : nsm * resident;
#define FPB_LAUNCH_APPLY(KK)                                                                       \
    do {                                                                                           \
        auto kern = w ? shared_apply_kernel<KK, true> : shared_apply_kernel<KK, false>;            \
        cudaError_t e_ = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,   \
                                              (int)smem);                                          \
        if (e_ != cudaSuccess) return e_;                                                          \
        kern<<<(unsigned)grid, kApplyThreads, smem, stream>>>(nbatch, m, n, y, y_sb, y_si, plan,   \
                                                              meta, c, ldc, fp);                   \
    } while (0)
    switch (k) {
    case 1: FPB_LAUNCH_APPLY(1); break;
    case 2: FPB_LAUNCH_APPLY(2); break;
    case 3: FPB_LAUNCH_APPLY(3); break;
    case 4: FPB_LAUNCH_APPLY(4); break;
    case 5: FPB_LAUNCH_APPLY(5); break;
    default: return cudaErrorInvalidValue;
    }
#undef FPB_LAUNCH_APPLY
    return cudaGetLastError();
}

}  // namespace fpb
