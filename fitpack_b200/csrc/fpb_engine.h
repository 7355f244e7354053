// fpb_engine.h -- per-GPU engine object behind the C-ABI (include/fitpack_b200.h).
#pragma once

#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>
#include <functional>
#include <nvtx3/nvToolsExt.h>
#include <string>
#include <vector>

#include "fpb_kernels.h"

namespace fpb {

void set_last_error(const std::string &msg);
bool cuda_ok(cudaError_t e, const char *what);

// grow-only device buffer
struct DevBuf {
    void *ptr = nullptr;
    size_t bytes = 0;
    bool reserve(size_t want);
    void release();
    template <class T> T *as() const { return static_cast<T *>(ptr); }
};

}  // namespace fpb

struct fpb_engine {
    int device = 0;
    int nsm = 148;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool timed = false;
    int64_t launches = 0;
    fpb::DevBuf ws, wsi, counter;     // fit kernel workspace
    fpb::DevBuf az_pool, az_off;      // (R, z) rows handed from the accepting pass to part 2
    fpb::DevBuf ws2, wsi2;            // second workspace: part 2 of the early finishers overlaps the tail passes
    cudaStream_t s_p2 = nullptr;      // stream of that early part-2 launch
    cudaEvent_t ev_p2[2] = {nullptr, nullptr};
    fpb::DevBuf lists, st_i, st_d, st_nrd;  // pass pipeline: curve lists and per-curve state
    int *h_count = nullptr;           // pinned host word for list-length read-back
    fpb::DevBuf plan_cs, plan_w, plan_row, plan_skip, plan_R, plan_ier, plan_meta;  // shared-x plan
    fpb::DevBuf scratch[16];          // staging for the host-pointer entry points
    fpb::DevBuf scratch2[16];         // second staging set (copy/compute overlap)
    // pinned host staging of the packed knots/coefficients, one per staging set (grow-only)
    double *h_pack[2] = {nullptr, nullptr};
    size_t h_pack_bytes[2] = {0, 0};
    long long *h_off[2] = {nullptr, nullptr};
    size_t h_off_bytes[2] = {0, 0};
    cudaStream_t s_in = nullptr, s_out = nullptr;
    cudaEvent_t ev_h2d[2] = {nullptr, nullptr};   // timing of the copy stream (wavefront host path)
    std::vector<cudaEvent_t> ev_chunk;   // wavefront host path: one 'chunk has landed' event per chunk
    cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_fit[2] = {nullptr, nullptr};
    double h_aux[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}};
    double *h_solo = nullptr;         // pinned staging of the single-curve fit (fp_curfit_c)
    size_t h_solo_bytes = 0;
    // gridded surfaces (fpb_regrid.cu): device tables, pinned read-back of the row/column sums
    fpb::DevBuf grid[16];
    double *grid_pinned = nullptr;
    size_t grid_pinned_bytes = 0;
    int64_t grid_fpgrre_calls = 0;
    int shared_fast = 0;              // shared-abscissa LSQ: 0 exact rotations, 1 FMA (fast)
    int bispev_fast = 0;              // bispev: 0 exact (reference order, no FMA), 1 fast (FMA sums)
    int curev_fast = 0;               // curev: 0 exact (reference arithmetic), 1 fast (local polynomials + Horner)
    fpb::DevBuf curev_n;              // curev fast path: knot counts with rejected curves zeroed
    fpb::DevBuf calc[6];              // spline calculus (fpb_calculus.cu)
};

namespace fpb {
// default engine of a device (created on first use, lives until process exit)
fpb_engine *default_engine(int device);
// Thread-safety contract of the host-pointer entry points: the reference's
// curfit/splev/regrid/parcur/percur are `pure` (core:1240, 17826, 16732, 15201, 15784), i.e. safe to call
// from several threads on distinct buffers.  Here every host-pointer entry point stages through the
// device's default engine (scratch buffers, stream), so each of them holds this per-device lock for
// the duration of the call: concurrent callers are correct, and serialised per GPU.  Recursive, because
// the batch entry points take it per shard and then call the single-device paths.
// NVTX range (tracing hooks) -- shows the pipeline stages in Nsight timelines; a no-op
// when no tool is attached.
struct NvtxRange {
    explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange &) = delete;
    NvtxRange &operator=(const NvtxRange &) = delete;
};

struct HostApiLock {
    explicit HostApiLock(int device);  // device < 0: the calling thread's current device
    ~HostApiLock();
    HostApiLock(const HostApiLock &) = delete;
    HostApiLock &operator=(const HostApiLock &) = delete;
    int slot = -1;
};
// the curfit / parcur pass pipeline (fpb_engine.cu): `p` carries the caller's arrays and scalars
// `arrivals` (optional): the batch reaches the device in chunks -- chunk i = curves [c0, c0+nc) is usable once
// `ready` has fired; its curves join the running pipeline at the next launch after that (host-pointer path)
struct FitArrival {
    int c0, nc;
    cudaEvent_t ready;
};
int32_t run_fit_pipeline(fpb_engine *eng, FitParams p, const std::vector<FitArrival> *arrivals = nullptr);
// A few parametric curves from host memory on the warp-per-curve kernel (fp_parcur_c, small fp_parcur_batch_c slices;
// fpb_curfit_solo.cu), on `eng` (the caller holds its HostApiLock).  Returns false when the call is not one for that
// kernel (idim other than 2 or 3, too large for its shared memory, too many curves, switched off): the caller then
// takes the batch path.  Otherwise ier[b] is the reference's flag, or *status / ier[] an FPB_ERR_* code.
bool parcur_solo_host(fpb_engine *eng, long long nb, int32_t iopt, int32_t ipar, int32_t idim, int32_t m, double *u,
                      const double *x, const double *w, double *ub, double *ue, int32_t k, double s, int32_t nest,
                      int32_t *n, double *t, double *c, double *fp, int32_t *ier, double *fpint, int32_t *nrdata,
                      int32_t *status);
// contiguous slices [b0, b1) of `total` units over the first `ngpus` devices (<= 0: all), one host
// thread + default engine per device; fn(engine, b0, b1) runs on each (no collective)
int32_t shard_units_over_gpus(long long total, int ngpus, long long align,
                              const std::function<int32_t(fpb_engine *, long long, long long)> &fn);
}
