// fpb_util.cu -- layout helpers: tile transpose between the caller's curve-major
// arrays ([curve][point], what a Fortran/C caller of curfit holds) and the
// point-major SoA layout ([point][curve]) the fit kernels stream with coalesced
// loads.  HBM-bound, 16 B/element of traffic; 32x32 tiles staged in shared memory
// (padded to 33 columns: conflict-free for 8-byte elements in both phases).
#include "fpb_common.cuh"
#include "fpb_kernels.h"

namespace fpb {

namespace {

__global__ void __launch_bounds__(256) transpose_kernel(const double *__restrict__ src,
                                                        long long ld_src, double *__restrict__ dst,
                                                        long long ld_dst, long long rows,
                                                        long long cols, long long tiles_c)
{
    __shared__ double tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
    const long long ntiles = tiles_c * ((rows + 31) / 32);
    for (long long tid = blockIdx.x; tid < ntiles; tid += gridDim.x) {
        const long long tr = tid / tiles_c, tc = tid % tiles_c;
        const long long r0 = tr * 32, c0 = tc * 32;
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
            const long long r = r0 + ty + j, c = c0 + tx;
            if (r < rows && c < cols) tile[ty + j][tx] = src[r * ld_src + c];
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
            const long long c = c0 + ty + j, r = r0 + tx;
            if (r < rows && c < cols) dst[c * ld_dst + r] = tile[tx][ty + j];
        }
        __syncthreads();
    }
}

__global__ void fill_i32_kernel(int *dst, int value, long long count)
{
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < count;
         i += (long long)gridDim.x * blockDim.x)
        dst[i] = value;
}

// Pack the meaningful part of every fitted curve -- t(1:n) and c(1:n-k-1) -- into two dense arrays for the
// device->host copy of the host-pointer entry points: at config 2 the mean knot count is 20 while the rows
// are nest = 359 doubles wide and the batch maximum is ~108, so the dense form is 5x smaller than the
// leading-columns 2-D copy it replaces.  One warp packs 32 curves: lengths are prefix-summed in the warp,
// one atomic per warp reserves the space (the order of the curves in the dense arrays is therefore not
// deterministic; off[b] records where curve b went), then the warp copies curve after curve with coalesced
// loads and stores.  Curves rejected with ier = 10 have length 0 (their t, c are not outputs).
__global__ void __launch_bounds__(256) pack_tc_kernel(int nb, int nest, int k, const int *__restrict__ n,
                                                      const int *__restrict__ ier, const double *__restrict__ t,
                                                      const double *__restrict__ c, unsigned long long *cursor,
                                                      long long *__restrict__ off, double *__restrict__ tp,
                                                      double *__restrict__ cp)
{
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long b0 = warp * 32; b0 < nb; b0 += nwarps * 32) {
        const long long b = b0 + lane;
        int len = 0;
        if (b < nb && ier[b] != 10) {
            len = n[b];
            if (len < 0 || len > nest) len = 0;
        }
        int incl = len;   // inclusive prefix sum over the warp
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += v;
        }
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        unsigned long long base = 0;
        if (lane == 0 && total > 0) base = atomicAdd(cursor, (unsigned long long)total);
        base = __shfl_sync(0xffffffffu, base, 0);
        const long long mine = (long long)base + incl - len;
        if (b < nb) off[b] = len > 0 ? mine : -1;
        for (int j = 0; j < 32; ++j) {
            const int lj = __shfl_sync(0xffffffffu, len, j);
            if (lj == 0) continue;
            const long long oj = __shfl_sync(0xffffffffu, mine, j);
            const double *tj = t + (b0 + j) * nest, *cj = c + (b0 + j) * nest;
            for (int i = lane; i < lj; i += 32) {
                tp[oj + i] = tj[i];
                cp[oj + i] = i < lj - k - 1 ? cj[i] : 0.0;
            }
        }
    }
}

}  // namespace

cudaError_t launch_pack_tc(int nb, int nest, int k, const int *n, const int *ier, const double *t, const double *c,
                           unsigned long long *cursor, long long *off, double *tp, double *cp, int nsm,
                           cudaStream_t stream)
{
    if (nb <= 0) return cudaSuccess;
    long long blocks = ((long long)nb + 255) / 256;
    if (blocks > (long long)nsm * 8) blocks = (long long)nsm * 8;
    pack_tc_kernel<<<(unsigned)blocks, 256, 0, stream>>>(nb, nest, k, n, ier, t, c, cursor, off, tp, cp);
    return cudaGetLastError();
}

cudaError_t launch_transpose(const double *src, long long ld_src, double *dst, long long ld_dst,
                             long long rows, long long cols, cudaStream_t stream)
{
    if (rows <= 0 || cols <= 0) return cudaSuccess;
    const long long tiles_c = (cols + 31) / 32;
    long long ntiles = tiles_c * ((rows + 31) / 32);
    const long long grid = ntiles < 148LL * 32 ? ntiles : 148LL * 32;
    transpose_kernel<<<(unsigned)grid, 256, 0, stream>>>(src, ld_src, dst, ld_dst, rows, cols,
                                                         tiles_c);
    return cudaGetLastError();
}

cudaError_t launch_fill_i32(int *dst, int value, long long count, cudaStream_t stream)
{
    if (count <= 0) return cudaSuccess;
    long long grid = (count + 255) / 256;
    if (grid > 148 * 8) grid = 148 * 8;
    fill_i32_kernel<<<(unsigned)grid, 256, 0, stream>>>(dst, value, count);
    return cudaGetLastError();
}

}  // namespace fpb

// ---- FP64 pipe probe (roofline denominator for the fit kernel; MEASURED_PEAKS.json has no
// FP64 figure).  8 independent dependent-chains per thread; fused=0 issues DMUL+DADD pairs
// (what the parity-exact kernels are allowed to use), fused=1 issues DFMA.
namespace fpb {
namespace {
template <bool FUSED>
__global__ void __launch_bounds__(256) fp64_probe_kernel(double *out, int iters, double a, double b)
{
    double v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = 1.0 + 1e-9 * (threadIdx.x + j);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (FUSED) v[j] = fma(v[j], a, b);
            else v[j] = __dadd_rn(__dmul_rn(v[j], a), b);
        }
    }
    double s = 0.0;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += v[j];
    if (s == 123.456) out[0] = s;  // keep the chains alive
}
}  // namespace

// mode 2: IEEE division, mode 3: IEEE sqrt (4 independent chains per thread)
template <int MODE>
__global__ void __launch_bounds__(256) fp64_divsqrt_probe_kernel(double *out, int iters, double a)
{
    double v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = 1.5 + 1e-3 * (threadIdx.x + j);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (MODE == 2) v[j] = a / v[j];
            else v[j] = sqrt(v[j]) + a;
        }
    }
    double s = 0.0;
#pragma unroll
    for (int j = 0; j < 4; ++j) s += v[j];
    if (s == 123.456) out[0] = s;
}

// modes 4..6: DEPENDENT-issue latency in SM cycles of a chain of DMUL/DADD pairs (4), of DFMA (5) and of
// MUFU.RCP64H + DFMA (6): one warp per SM, clock64() around `iters` dependent instructions; out[0] = cycles
// per instruction.  The fit kernels are chains of dependent FP64 instructions (ILP ~ 1 per warp), so their
// issue rate is bounded by (warps per scheduler) / (this latency).
template <int MODE>
__global__ void fp64_latency_probe_kernel(double *out, int iters, double a, double b)
{
    double v = 1.0 + 1e-9 * threadIdx.x;
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        if (MODE == 4) {
            v = __dmul_rn(v, a);
            v = __dadd_rn(v, b);
        } else if (MODE == 5) {
            v = __fma_rn(v, a, b);
            v = __fma_rn(v, a, b);
        } else {
            double r;
            asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(v));
            v = __fma_rn(r, a, b);
        }
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (double)(t1 - t0) / (2.0 * iters);
    if (v == 123.456) out[1] = v;
}

cudaError_t launch_fp64_probe(double *scratch, int blocks, int iters, int fused, cudaStream_t stream)
{
    if (fused >= 4 && fused <= 6) {
        const int n = 1 << 14;
        if (fused == 4) fp64_latency_probe_kernel<4><<<1, 32, 0, stream>>>(scratch, n, 0.999999, 1e-7);
        else if (fused == 5) fp64_latency_probe_kernel<5><<<1, 32, 0, stream>>>(scratch, n, 0.999999, 1e-7);
        else fp64_latency_probe_kernel<6><<<1, 32, 0, stream>>>(scratch, n, 0.999999, 1e-7);
        return cudaGetLastError();
    }
    if (fused == 2) {
        fp64_divsqrt_probe_kernel<2><<<blocks, 256, 0, stream>>>(scratch, iters, 1.75);
        return cudaGetLastError();
    }
    if (fused == 3) {
        fp64_divsqrt_probe_kernel<3><<<blocks, 256, 0, stream>>>(scratch, iters, 0.75);
        return cudaGetLastError();
    }
    if (fused) fp64_probe_kernel<true><<<blocks, 256, 0, stream>>>(scratch, iters, 0.999999, 1e-7);
    else fp64_probe_kernel<false><<<blocks, 256, 0, stream>>>(scratch, iters, 0.999999, 1e-7);
    return cudaGetLastError();
}
}  // namespace fpb

// ---- self-test of the hand-written division / square-root sequences (fpb_common.cuh) against the compiler's
// IEEE operators on random operands: counts bitwise mismatches of (0) div_refined(a, d, rcp_refined(d)) vs a / d
// inside the guarded ranges, (1) sqrt_fastpath(x) vs sqrt(x) for x in [1, 2], (2) givens_fast vs givens (cos, sin
// and the new diagonal) on operands spread over 600 binades incl. zeros, negatives and out-of-range values that
// must take the fallback.  tests/test_gpu_fastmath.py requires zero mismatches.
namespace fpb {
namespace {
__device__ __forceinline__ unsigned long long sm64(unsigned long long &s)
{
    unsigned long long z = (s += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
// random significand and sign bit, exponent uniform in [emin, emax]
__device__ __forceinline__ double rnd_double(unsigned long long &s, int emin, int emax, bool allow_neg)
{
    const unsigned long long r = sm64(s), r2 = sm64(s);
    const unsigned long long e = (unsigned long long)(1023 + emin + (int)(r2 % (unsigned long long)(emax - emin + 1)));
    unsigned long long b = (r & 0xFFFFFFFFFFFFFULL) | (e << 52);
    if (allow_neg && (r2 >> 40 & 1)) b |= 0x8000000000000000ULL;
    return __longlong_as_double((long long)b);
}
__global__ void __launch_bounds__(256) fastmath_selftest_kernel(unsigned long long seed, int iters,
                                                                unsigned long long *mismatch)
{
    unsigned long long s = seed + 0x1234567ULL * (blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x + 1);
    unsigned long long bad_div = 0, bad_sqrt = 0, bad_giv = 0;
    for (int it = 0; it < iters; ++it) {
        // (0) division inside the guarded ranges (plus numerator +0)
        {
            const double d = rnd_double(s, -400, 400, false);
            double a = rnd_double(s, -500, 500, false);
            if ((it & 63) == 0) a = 0.0;
            if ((it & 7) == 1) {   // quotients close to rounding boundaries: a = q*d nudged by an ulp
                const double q = rnd_double(s, -3, 3, false);
                a = __longlong_as_double(__double_as_longlong(__dmul_rn(q, d)) + ((it & 8) ? 1 : -1));
            }
            if (rcp_den_ok(d) && rcp_num_ok(a)) {
                const double q1 = div_refined(a, d, rcp_refined(d));
                const double q2 = a / d;
                if (__double_as_longlong(q1) != __double_as_longlong(q2)) ++bad_div;
            }
        }
        // (1) square root on [1, 2]
        {
            const double r = rnd_double(s, -60, 0, false);
            const double x = __dadd_rn(1.0, __dmul_rn(r, r));
            if (x >= 1.0 && x <= 2.0 &&
                __double_as_longlong(sqrt_fastpath(x)) != __double_as_longlong(sqrt(x)))
                ++bad_sqrt;
        }
        // (2) fpgivs: wide exponent spread, zeros and out-of-range operands (fallback path) included
        {
            double piv = rnd_double(s, -300, 300, true);
            double ww = rnd_double(s, -300, 300, false);
            const int sel = it & 31;
            if (sel == 0) ww = 0.0;
            else if (sel == 1) piv = rnd_double(s, -1000, -600, true);
            else if (sel == 2) ww = rnd_double(s, 450, 900, false);
            else if (sel == 3) ww = __dmul_rn(fabs(piv), 1.0 + 1e-15 * (double)(it & 255));
            double w1 = ww, w2 = ww, c1, s1, c2, s2;
            givens_fast(piv, w1, c1, s1);
            givens(piv, w2, c2, s2);
            if (__double_as_longlong(w1) != __double_as_longlong(w2) || __double_as_longlong(c1) != __double_as_longlong(c2) ||
                __double_as_longlong(s1) != __double_as_longlong(s2))
                ++bad_giv;
        }
    }
    if (bad_div) atomicAdd(&mismatch[0], bad_div);
    if (bad_sqrt) atomicAdd(&mismatch[1], bad_sqrt);
    if (bad_giv) atomicAdd(&mismatch[2], bad_giv);
}
}  // namespace

cudaError_t launch_fastmath_selftest(unsigned long long seed, int blocks, int iters, unsigned long long *mismatch,
                                     cudaStream_t stream)
{
    fastmath_selftest_kernel<<<blocks, 256, 0, stream>>>(seed, iters, mismatch);
    return cudaGetLastError();
}
}  // namespace fpb
