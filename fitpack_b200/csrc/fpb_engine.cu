// fpb_engine.cu -- host side of the C-ABI: engine lifetime, argument checks that
// curfit/splev make before any arithmetic (core:1259-1270, 17844-17845), workspace
// sizing, kernel launches, host<->device staging for the host-pointer entry points
// and per-GPU sharding of a batch.  There is no CPU compute path in this file: every
// entry point either runs the CUDA kernels or reports FPB_ERR_CUDA.
#include "fpb_engine.h"

#include <algorithm>
#include <atomic>
#include <cfloat>
#include <cmath>
#include <cstdlib>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <future>
#include <mutex>
#include <thread>
#include <vector>

#include "../../include/fitpack_b200.h"

namespace fpb {

static thread_local std::string g_last_error;

void set_last_error(const std::string &msg) { g_last_error = msg; }

bool cuda_ok(cudaError_t e, const char *what)
{
    if (e == cudaSuccess) return true;
    set_last_error(std::string(what) + ": " + cudaGetErrorString(e));
    return false;
}

bool DevBuf::reserve(size_t want)
{
    if (want <= bytes) return true;
    if (ptr) cudaFree(ptr);
    ptr = nullptr;
    bytes = 0;
    // round up so that slowly growing batches do not reallocate every call
    size_t cap = (want + ((size_t)1 << 20) - 1) & ~(((size_t)1 << 20) - 1);
    if (!cuda_ok(cudaMalloc(&ptr, cap), "cudaMalloc")) {
        ptr = nullptr;
        return false;
    }
    bytes = cap;
    return true;
}

void DevBuf::release()
{
    if (ptr) cudaFree(ptr);
    ptr = nullptr;
    bytes = 0;
}

static std::mutex g_engines_mu;
static fpb_engine *g_engines[256] = {nullptr};   // slot device + 32*w: engine of host worker w (0..7)

fpb_engine *default_engine(int device)
{
    if (device < 0) {
        if (cudaGetDevice(&device) != cudaSuccess) {
            set_last_error("no CUDA device");
            return nullptr;
        }
    }
    if (device >= 256) return nullptr;
    std::lock_guard<std::mutex> lk(g_engines_mu);
    if (!g_engines[device]) {
        // slots 32.. hold further engines for devices 0..31 (copy/compute overlap workers)
        fpb_engine *e = nullptr;
        if (fpb_engine_create(device & 31, &e) != FPB_STATUS_OK) return nullptr;
        g_engines[device] = e;
    }
    return g_engines[device];
}

static std::recursive_mutex g_host_api_mu[32];

HostApiLock::HostApiLock(int device)
{
    if (device < 0 && cudaGetDevice(&device) != cudaSuccess) device = 0;  // no device: callers fail later
    slot = device & 31;
    g_host_api_mu[slot].lock();
}

HostApiLock::~HostApiLock()
{
    if (slot >= 0) g_host_api_mu[slot].unlock();
}

struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev)
    {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != dev) ok = cuda_ok(cudaSetDevice(dev), "cudaSetDevice");
    }
    ~DeviceGuard()
    {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

static void begin_timing(fpb_engine *eng)
{
    cudaEventRecord(eng->ev0, eng->stream);
}
static void end_timing(fpb_engine *eng)
{
    cudaEventRecord(eng->ev1, eng->stream);
    eng->timed = true;
}

}  // namespace fpb

using namespace fpb;

extern "C" {

const char *fpb_last_error(void) { return g_last_error.c_str(); }
const char *fpb_version(void) { return "fitpack_b200 0.1 (sm_100a)"; }

int32_t fpb_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

int32_t fpb_engine_create(int32_t device, fpb_engine **out)
{
    if (!out) return FPB_ERR_ARG;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        set_last_error("no CUDA device visible: fitpack_b200 has no CPU fallback");
        return FPB_ERR_CUDA;
    }
    if (device < 0 && !cuda_ok(cudaGetDevice(&device), "cudaGetDevice")) return FPB_ERR_CUDA;
    if (device >= ndev) {
        set_last_error("device index out of range");
        return FPB_ERR_ARG;
    }
    DeviceGuard g(device);
    if (!g.ok) return FPB_ERR_CUDA;
    fpb_engine *e = new fpb_engine();
    e->device = device;
    cudaDeviceProp prop;
    if (!cuda_ok(cudaGetDeviceProperties(&prop, device), "cudaGetDeviceProperties")) {
        delete e;
        return FPB_ERR_CUDA;
    }
    e->nsm = prop.multiProcessorCount;
    if (!cuda_ok(cudaStreamCreateWithFlags(&e->own_stream, cudaStreamNonBlocking), "stream") ||
        !cuda_ok(cudaEventCreate(&e->ev0), "event") || !cuda_ok(cudaEventCreate(&e->ev1), "event")) {
        delete e;
        return FPB_ERR_CUDA;
    }
    e->stream = e->own_stream;
    *out = e;
    return FPB_STATUS_OK;
}

void fpb_engine_destroy(fpb_engine *eng)
{
    if (!eng) return;
    DeviceGuard g(eng->device);
    cudaStreamSynchronize(eng->stream);
    eng->ws.release();
    eng->wsi.release();
    eng->ws2.release();
    eng->wsi2.release();
    eng->az_pool.release();
    eng->az_off.release();
    eng->curev_n.release();
    if (eng->s_p2) cudaStreamDestroy(eng->s_p2);
    for (int i = 0; i < 2; ++i)
        if (eng->ev_p2[i]) cudaEventDestroy(eng->ev_p2[i]);
    eng->counter.release();
    eng->lists.release();
    eng->st_i.release();
    eng->st_d.release();
    eng->st_nrd.release();
    if (eng->h_count) cudaFreeHost(eng->h_count);
    for (int i = 0; i < 2; ++i) {
        if (eng->h_pack[i]) cudaFreeHost(eng->h_pack[i]);
        if (eng->h_off[i]) cudaFreeHost(eng->h_off[i]);
    }
    for (auto &b : eng->scratch2) b.release();
    for (auto e : eng->ev_chunk) cudaEventDestroy(e);
    for (int i = 0; i < 2; ++i)
        if (eng->ev_h2d[i]) cudaEventDestroy(eng->ev_h2d[i]);
    if (eng->s_in) cudaStreamDestroy(eng->s_in);
    if (eng->s_out) cudaStreamDestroy(eng->s_out);
    for (int i = 0; i < 2; ++i) {
        if (eng->ev_in[i]) cudaEventDestroy(eng->ev_in[i]);
        if (eng->ev_fit[i]) cudaEventDestroy(eng->ev_fit[i]);
    }
    eng->plan_cs.release();
    eng->plan_w.release();
    eng->plan_row.release();
    eng->plan_skip.release();
    eng->plan_R.release();
    eng->plan_meta.release();
    eng->plan_ier.release();
    for (auto &b : eng->scratch) b.release();
    for (auto &b : eng->grid) b.release();
    for (auto &b : eng->calc) b.release();
    if (eng->grid_pinned) cudaFreeHost(eng->grid_pinned);
    if (eng->ev0) cudaEventDestroy(eng->ev0);
    if (eng->ev1) cudaEventDestroy(eng->ev1);
    if (eng->own_stream) cudaStreamDestroy(eng->own_stream);
    delete eng;
}

int32_t fpb_engine_set_stream(fpb_engine *eng, void *cuda_stream)
{
    if (!eng) return FPB_ERR_ARG;
    // taken literally: NULL is CUDA's default (legacy) stream, which is what torch uses
    eng->stream = static_cast<cudaStream_t>(cuda_stream);
    return FPB_STATUS_OK;
}

int32_t fpb_engine_reset_stream(fpb_engine *eng)
{
    if (!eng) return FPB_ERR_ARG;
    eng->stream = eng->own_stream;
    return FPB_STATUS_OK;
}

int32_t fpb_engine_synchronize(fpb_engine *eng)
{
    if (!eng) return FPB_ERR_ARG;
    DeviceGuard g(eng->device);
    return cuda_ok(cudaStreamSynchronize(eng->stream), "cudaStreamSynchronize") ? FPB_STATUS_OK
                                                                                : FPB_ERR_CUDA;
}

int64_t fpb_engine_launch_count(const fpb_engine *eng) { return eng ? eng->launches : 0; }

int64_t fpb_default_engine_launch_count(int32_t device)
{
    if (device < 0 || device >= 32) return 0;
    std::lock_guard<std::mutex> lk(g_engines_mu);
    int64_t total = 0;  // all default engines of the device (engines 1..3 are the overlap workers)
    for (int w = 0; w < 8; ++w)
        if (g_engines[device + 32 * w]) total += g_engines[device + 32 * w]->launches;
    return total;
}

int32_t fpb_engine_set_shared_mode(fpb_engine *eng, int32_t mode)
{
    if (!eng || (mode != 0 && mode != 1)) return FPB_ERR_ARG;
    eng->shared_fast = mode;
    return FPB_STATUS_OK;
}

int32_t fpb_engine_set_bispev_mode(fpb_engine *eng, int32_t mode)
{
    if (!eng || (mode != 0 && mode != 1)) return FPB_ERR_ARG;
    eng->bispev_fast = mode;
    return FPB_STATUS_OK;
}

int32_t fpb_engine_set_curev_mode(fpb_engine *eng, int32_t mode)
{
    if (!eng || (mode != 0 && mode != 1)) return FPB_ERR_ARG;
    eng->curev_fast = mode;
    return FPB_STATUS_OK;
}

double fpb_engine_last_kernel_ms(fpb_engine *eng)
{
    if (!eng || !eng->timed) return -1.0;
    DeviceGuard g(eng->device);
    float ms = -1.0f;
    if (cudaEventSynchronize(eng->ev1) != cudaSuccess) return -1.0;
    if (cudaEventElapsedTime(&ms, eng->ev0, eng->ev1) != cudaSuccess) return -1.0;
    return (double)ms;
}

FP_BOOL FITPACK_SUCCESS_c(FP_FLAG ierr) { return ierr <= 0; }

void fitpack_message_c(FP_FLAG ierr, char *msg)
{
    // strings of FITPACK_MESSAGE, reference core:315-339
    const char *s;
    switch (ierr) {
    case 0: s = "Success!"; break;
    case -1: s = "Success! (interpolation)"; break;
    case -2: s = "Success! (least-squares)"; break;
    case 1: s = "Insufficient Storage"; break;
    case 2: s = "Smoothing parameter is too small"; break;
    case 3: s = "Infinite loop detected"; break;
    case 4: s = "More knots than data points"; break;
    case 5: s = "Overlapping knots found"; break;
    case 6: s = "Invalid variable range"; break;
    case 10: s = "Invalid input"; break;
    case 11: s = "Test(s) failed"; break;
    case 12: s = "Invalid constraint(s) provided"; break;
    case 13: s = "Not enough knots (n>=10)"; break;
    case 14: s = "concon: maxbin too small"; break;
    case 15: s = "concon: maxtr too small"; break;
    case 16: s = "concon: QP solver failed"; break;
    default: s = "UNKNOWN ERROR"; break;
    }
    if (msg) std::strcpy(msg, s);
}

// ------------------------------------------------------------------ device API

} // extern "C"

namespace {
// a few curves: the warp-per-curve / curve-per-CTA kernels (1) or the batch kernels (0); fpb_set_single_curve_kernel
std::atomic<int> g_solo_on{(std::getenv("FPB_SOLO") && std::atoi(std::getenv("FPB_SOLO")) == 0) ? 0 : 1};
constexpr int kPerSoloMaxBatch = 1024;
}  // namespace

namespace fpb {
// The pass pipeline shared by curfit (idim = 1) and parcur: `p` arrives with the caller's arrays,
// strides and scalars set; this fills in the engine-owned workspace / state and runs the passes.
int32_t run_fit_pipeline(fpb_engine *eng, FitParams p, const std::vector<FitArrival> *arrivals)
{
    NvtxRange nvtx_pipeline(p.periodic ? "fpb:percur pipeline" : (p.par ? "fpb:parcur pipeline" : "fpb:curfit pipeline"));
    const int nbatch = p.nbatch, iopt = p.iopt, m = p.m, k = p.k, nest = p.nest;
    int *n = p.n;
    const FitLayout L = fit_layout(m, k, nest, p.idim, p.periodic);
    long long blocks_needed = ((long long)nbatch + kFitThreads - 1) / kFitThreads;
    // the workspace is sized for the larger of the two kernels' resident grids
    const int max_warps = p.periodic ? percur_resident_warps(k, eng->nsm) : fit_max_resident_warps(k, p.idim, eng->nsm);
    const long long grid_ws = std::min<long long>(blocks_needed, max_warps / (kFitThreads / 32));
    long long grid = std::min<long long>(
        blocks_needed, p.periodic ? (long long)max_warps / (kFitThreads / 32)
                                  : (long long)eng->nsm * fit_blocks_per_sm(k, idim_class(p.idim)));
    const long long grid_p2 = std::min<long long>(
        blocks_needed, (p.periodic ? max_warps : fit_p2_resident_warps(k, p.idim, eng->nsm)) / (kFitThreads / 32));
    const size_t warps = (size_t)grid_ws * (kFitThreads / 32);
    const size_t nb = (size_t)nbatch;
    // counters: [0] tile queue head, [4] list A count, [5] list B count, [6] part-2 count,
    // [16..17] cursor of the hand-over pool; from [64] on: one histogram of n per part-2 section, then the
    // sections' counts and tile-queue heads
    const size_t hist_off = 64;  // ints
    const size_t hist_len = (size_t)L.nestp + 8;
    constexpr int kMaxSec = 24;   // part-2 sections on the side stream (+ the final one on the main stream)
    const size_t sec_off = hist_off + (size_t)(kMaxSec + 1) * hist_len;   // per section: count, tile-queue head
    const size_t cnt_ints = sec_off + 2 * (kMaxSec + 1) + 8;
    if (!eng->ws.reserve(warps * (size_t)L.ws_per_lane * 32 * sizeof(double)) ||
        !eng->wsi.reserve(warps * (size_t)L.nestp * 32 * sizeof(int)) ||
        !eng->counter.reserve(cnt_ints * sizeof(int)) ||
        !eng->lists.reserve(4 * nb * sizeof(int)) ||
        !eng->st_i.reserve(2 * nb * sizeof(int)) || !eng->st_d.reserve(2 * nb * sizeof(double)) ||
        !eng->st_nrd.reserve(nb * (size_t)L.nestp * sizeof(int)))
        return FPB_ERR_ALLOC;
    if (!eng->h_count) {
        if (!cuda_ok(cudaHostAlloc((void **)&eng->h_count, 1024, cudaHostAllocDefault), "cudaHostAlloc"))
            return FPB_ERR_ALLOC;
    }
    int *cnt = eng->counter.as<int>();
    cudaStream_t st = eng->stream;
    if (!cuda_ok(cudaMemsetAsync(cnt, 0, cnt_ints * sizeof(int), st), "memset"))
        return FPB_ERR_CUDA;
    p.nestp = L.nestp;
    {
        const int cls = idim_class(p.idim);
        const int blocks = std::max(fit_blocks_per_sm(k, cls), fit_p2_blocks_per_sm(k, cls));
        p.sdepth = fit_stream_depth(k, p.idim, p.w != nullptr, blocks);
    }
    p.ws = eng->ws.as<double>(); p.wsi = eng->wsi.as<int>();
    p.counter = reinterpret_cast<unsigned *>(cnt);
    p.off_t = L.off_t; p.off_c = L.off_c; p.off_z = L.off_z; p.off_fpint = L.off_fpint;
    p.off_a = L.off_a; p.off_g = L.off_g; p.off_b = L.off_b; p.off_q = L.off_q;
    p.off_x = L.off_x; p.off_y = L.off_y; p.off_w = L.off_w;
    p.off_a2 = L.off_a2; p.off_g2 = L.off_g2;
    p.ws_per_lane = L.ws_per_lane;
    p.st_nplus = eng->st_i.as<int>(); p.st_iter = p.st_nplus + nb;
    p.st_fpold = eng->st_d.as<double>(); p.st_fp0 = p.st_fpold + nb;
    p.st_nrd = eng->st_nrd.as<int>();
    // pool for the (R, z) rows of accepted knot sets: room for 96 rows per curve on average (config 2 uses
    // 17), at most 8 GB; curves that do not fit regenerate the rows in part 2
    p.az_pool = nullptr; p.az_cursor = nullptr; p.az_cap = 0; p.az_off = nullptr;
    static const bool no_handover = std::getenv("FPB_FIT_NO_HANDOVER") != nullptr;
    if (!p.periodic && iopt >= 0 && !no_handover) {
        const long long rowlen = (k + 1) + p.idim;
        long long cap = (long long)nb * std::min<long long>(L.nestp, 96) * rowlen;
        cap = std::min<long long>(cap, ((long long)8 << 30) / 8);
        if (eng->az_pool.reserve((size_t)cap * sizeof(double)) && eng->az_off.reserve(nb * sizeof(long long))) {
            p.az_pool = eng->az_pool.as<double>();
            p.az_off = eng->az_off.as<long long>();
            p.az_cap = cap;
            p.az_cursor = reinterpret_cast<unsigned long long *>(cnt + 16);   // zeroed with the counters
        }
    }
    auto sec_hist = [&](int i) { return cnt + hist_off + (size_t)i * hist_len; };
    auto sec_count = [&](int i) { return cnt + sec_off + i; };
    auto sec_head = [&](int i) { return cnt + sec_off + (kMaxSec + 1) + i; };
    p.hist = sec_hist(0);
    int *list[2] = {eng->lists.as<int>(), eng->lists.as<int>() + nb};
    int *p2_list = eng->lists.as<int>() + 2 * nb, *p2_sorted = eng->lists.as<int>() + 3 * nb;
    int *lcount[2] = {cnt + 4, cnt + 5};
    int *p2_count = cnt + 6;

    // Early part 2 (curfit / parcur, large batches): once the curves that still need a finer knot set no
    // longer fill the GPU, the passes that remain are a chain of sparsely populated, latency-bound launches
    // (about 10 of ~20 launches, 6 % of the time at config 2).  Part 2 of every curve accepted so far starts
    // right then on a second stream with its own workspace and one CTA slot per SM left free for the tail;
    // the rest of the part-2 list follows after the last pass.
    const long long capacity = grid * (long long)kFitThreads;   // curves one resident wave of the pass kernel holds
    static const bool no_overlap = std::getenv("FPB_FIT_NO_OVERLAP") != nullptr;
    static const bool no_overlap_per = std::getenv("FPB_PER_NO_OVERLAP") != nullptr;
    const bool may_overlap = !no_overlap && !(p.periodic && no_overlap_per) && iopt >= 0 && (long long)nbatch >= 4 * capacity;
    // tuning knobs (profiles/r02_ab.md): cut-off point in units of capacity/4, CTAs per SM left to the tail
    static const int early_at4 = std::getenv("FPB_FIT_EARLY_AT4") ? std::atoi(std::getenv("FPB_FIT_EARLY_AT4")) : 4;
    static const int early_free = std::getenv("FPB_FIT_EARLY_FREE") ? std::atoi(std::getenv("FPB_FIT_EARLY_FREE")) : 1;
    int isec = 0, sec_start = 0;   // side-stream sections launched so far; first part-2 list entry not yet in a section
    bool tail_done = false;          // the cut at the start of the tail has been made

    begin_timing(eng);
    // Staged mode (host-pointer path): the chunks of the batch join the pipeline as their data arrives; every
    // launch then works on whatever is active -- curves of early chunks may be several knot sets ahead of the
    // late ones (a list entry with the top bit set has not had its first pass).  Otherwise: pass 1 over the
    // whole batch (identity list), then passes over the shrinking list of curves that still need a finer
    // knot set; the host reads the list length every other pass.
    // A few periodic curves: the whole of fpperi per curve in ONE launch, the curve's workspace in shared memory
    // (fpb_percur_solo.cu), instead of ~2R launches of threads working out of global memory with a host read-back
    // between them.
    if (p.periodic && nbatch <= kPerSoloMaxBatch && g_solo_on.load() &&
        percur_solo_smem_bytes(m, k, L.ws_per_lane, L.nestp) <= kSoloSmemMax) {
        if (!cuda_ok(launch_percur_solo(p, st), "percur_solo_kernel")) return FPB_ERR_CUDA;
        eng->launches++;
        end_timing(eng);
        return FPB_STATUS_OK;
    }
    const bool staged = arrivals != nullptr && !arrivals->empty() && iopt >= 0 && !p.periodic;
    p.staged = staged ? 1 : 0;
    size_t next_arr = 0;
    int cur = 0, remaining = 0;
    if (!staged) {
        if (!cuda_ok(launch_curfit_pass(p, (int)grid, st, nullptr, nullptr, list[0], lcount[0], p2_list,
                                        p2_count, 1), "curfit pass kernel"))
            return FPB_ERR_CUDA;
        eng->launches++;
        remaining = (iopt < 0) ? 0 : nbatch;  // iopt=-1: a single pass settles every curve
    }
    // admit every chunk that has arrived (at least one when nothing else is active)
    auto admit = [&]() -> bool {
        while (next_arr < arrivals->size()) {
            const FitArrival &a = (*arrivals)[next_arr];
            if (remaining > 0 && cudaEventQuery(a.ready) != cudaSuccess) break;
            if (!cuda_ok(cudaStreamWaitEvent(st, a.ready, 0), "wait") ||
                !cuda_ok(launch_fit_admit(list[cur], lcount[cur], remaining, a.c0, a.nc, st), "admit"))
                return false;
            eng->launches++;
            remaining += a.nc;
            ++next_arr;
        }
        return true;
    };
    const bool all_in = true;
    (void)all_in;
    const int max_rounds = 2 * (m + 4) + 8 + (staged ? 4 * (int)arrivals->size() : 0);  // fpcurf's own bound on knot sets, with one restart
    for (int round = 1; round <= max_rounds && (remaining > 0 || (staged && next_arr < arrivals->size())); ++round) {
        if (staged && next_arr < arrivals->size() && !admit()) return FPB_ERR_CUDA;
        const bool arrived = !staged || next_arr == arrivals->size();
        const bool near_tail = may_overlap && isec == 0 && arrived && remaining <= std::max<long long>(4, early_at4) * capacity;
        // staged: the host needs the exact list length before every launch (the next admission appends behind it);
        // the read comes after the launch below, so round 1 has nothing to read
        if (!staged && round >= 3 && ((round & 1) || near_tail)) {
            // [4],[5],[6]: both list lengths and the part-2 count in one copy
            if (!cuda_ok(cudaMemcpyAsync(eng->h_count, cnt + 4, 3 * sizeof(int), cudaMemcpyDeviceToHost, st),
                         "count D2H") ||
                !cuda_ok(cudaStreamSynchronize(st), "sync"))
                return FPB_ERR_CUDA;
            remaining = eng->h_count[cur];
            if (remaining <= 0) break;
        }
        {
            const int n2_now = eng->h_count[2];   // staged: read after the previous launch (below)
            const bool fresh_counts = staged ? round >= 2 : (round >= 3 && ((round & 1) || near_tail));
            // non-staged: ONE cut, when the curves still in part 1 no longer fill the GPU.  staged: the GPU follows
            // the arriving data with a small active population, so part 2 of the curves accepted so far is started
            // whenever two resident waves of them have piled up -- it runs beside the passes the whole time
            // instead of after the last chunk -- plus the final cut as in the non-staged case.
            const bool tail_cut = arrived && !tail_done && remaining <= capacity * early_at4 / 4 &&
                                  (long long)(n2_now - sec_start) >= (staged ? capacity / 2 : 2 * capacity);
            // (staged mode: cutting a section whenever two resident waves of accepted curves have piled up -- part 2
            // beside the passes all the way -- was measured and lost: 184.5 vs 168.7 ms end to end, the passes crawl
            // on the one CTA slot per SM the section leaves them; FPB_FIT_PILE_CUTS=1 re-enables it)
            static const bool pile_cuts = std::getenv("FPB_FIT_PILE_CUTS") != nullptr;
            const bool pile_cut = pile_cuts && staged && !tail_done && (long long)(n2_now - sec_start) >= 2 * capacity;
            if (may_overlap && fresh_counts && remaining > 0 && isec < kMaxSec && (pile_cut || tail_cut)) {
                if (tail_cut) tail_done = true;
                // ---- cut the part-2 list here and start the section [sec_start, n2_now) on the second stream ----
                if (!eng->s_p2) {
                    if (!cuda_ok(cudaStreamCreateWithFlags(&eng->s_p2, cudaStreamNonBlocking), "stream") ||
                        !cuda_ok(cudaEventCreateWithFlags(&eng->ev_p2[0], cudaEventDisableTiming), "event") ||
                        !cuda_ok(cudaEventCreateWithFlags(&eng->ev_p2[1], cudaEventDisableTiming), "event"))
                        return FPB_ERR_CUDA;
                }
                const int blocks_p2 = p.periodic ? max_warps / (kFitThreads / 32) / eng->nsm
                                                 : fit_p2_blocks_per_sm(k, idim_class(p.idim));
                const int blocks_early = std::max(1, blocks_p2 - early_free);
                const int nsec = n2_now - sec_start;
                const long long slots_early = (long long)eng->nsm * blocks_early;
                const long long grid_early = std::min<long long>(slots_early, ((long long)nsec + kFitThreads - 1) / kFitThreads);
                const size_t warps2 = (size_t)std::min<long long>(slots_early, blocks_needed) * (kFitThreads / 32);
                if (!eng->ws2.reserve(warps2 * (size_t)L.ws_per_lane * 32 * sizeof(double)) ||
                    !eng->wsi2.reserve(warps2 * (size_t)L.nestp * 32 * sizeof(int)))
                    return FPB_ERR_ALLOC;
                eng->h_count[32 + isec] = nsec;
                if (!cuda_ok(cudaMemcpyAsync(sec_count(isec), eng->h_count + 32 + isec, sizeof(int), cudaMemcpyHostToDevice, st),
                             "section count") ||
                    !cuda_ok(cudaEventRecord(eng->ev_p2[0], st), "event") ||
                    !cuda_ok(cudaStreamWaitEvent(eng->s_p2, eng->ev_p2[0], 0), "wait"))
                    return FPB_ERR_CUDA;
                FitParams pe = p;
                pe.ws = eng->ws2.as<double>();
                pe.wsi = eng->wsi2.as<int>();
                pe.counter = reinterpret_cast<unsigned *>(sec_head(isec));
                if (!cuda_ok(launch_p2_sort(p2_list + sec_start, sec_count(isec), n, sec_hist(isec), L.nestp + 1,
                                            p2_sorted + sec_start, eng->nsm, eng->s_p2), "part-2 sort") ||
                    !cuda_ok(launch_curfit_part2(pe, (int)grid_early, eng->s_p2, p2_sorted + sec_start, sec_count(isec)),
                             "curfit part-2 kernel (early)") ||
                    !cuda_ok(cudaEventRecord(eng->ev_p2[1], eng->s_p2), "event"))
                    return FPB_ERR_CUDA;
                eng->launches += 3;
                sec_start = n2_now;
                ++isec;
                p.hist = sec_hist(isec);   // curves accepted from now on are counted for the next section
            }
        }
        if (remaining <= 0) continue;   // staged: nothing active, wait for the next chunk
        // reset the queue head and the output list length
        if (!cuda_ok(cudaMemsetAsync(cnt, 0, sizeof(int), st), "memset") ||
            !cuda_ok(cudaMemsetAsync(lcount[1 - cur], 0, sizeof(int), st), "memset"))
            return FPB_ERR_CUDA;
        long long g2 = std::min<long long>(grid, ((long long)remaining + kFitThreads - 1) / kFitThreads);
        if (g2 < 1) g2 = 1;
        if (!cuda_ok(launch_curfit_pass(p, (int)g2, st, list[cur], lcount[cur], list[1 - cur],
                                        lcount[1 - cur], p2_list, p2_count, 0), "curfit pass kernel"))
            return FPB_ERR_CUDA;
        eng->launches++;
        cur = 1 - cur;
        if (staged) {   // exact survivor count (and the part-2 count) for the next admission / launch
            if (!cuda_ok(cudaMemcpyAsync(eng->h_count, cnt + 4, 3 * sizeof(int), cudaMemcpyDeviceToHost, st),
                         "count D2H") ||
                !cuda_ok(cudaStreamSynchronize(st), "sync"))
                return FPB_ERR_CUDA;
            remaining = eng->h_count[cur];
            static const bool trace = std::getenv("FPB_HOST_TRACE") != nullptr;
            if (trace) {
                static const auto t0 = std::chrono::steady_clock::now();
                std::fprintf(stderr, "[fpb staged] %9.2f ms  round %d  chunks in %zu/%zu  active %d  part-2 list %d  side sections %d\n",
                             std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(),
                             round, next_arr, arrivals->size(), remaining, eng->h_count[2], isec);
            }
        }
    }
    // part 2 on the curves whose knots were accepted with fp < s
    if (iopt >= 0) {
        if (!cuda_ok(cudaMemcpyAsync(eng->h_count, p2_count, sizeof(int), cudaMemcpyDeviceToHost, st),
                     "count D2H") ||
            !cuda_ok(cudaStreamSynchronize(st), "sync"))
            return FPB_ERR_CUDA;
        const int n2 = eng->h_count[0];
        const int first = sec_start;   // entries [first, n2) of the part-2 list are still to do
        if (n2 > first) {
            const int *count_ptr = p2_count;
            if (isec > 0) {
                eng->h_count[32 + isec] = n2 - first;
                if (!cuda_ok(cudaMemcpyAsync(sec_count(isec), eng->h_count + 32 + isec, sizeof(int), cudaMemcpyHostToDevice, st),
                             "rest count"))
                    return FPB_ERR_CUDA;
                count_ptr = sec_count(isec);
            }
            if (!cuda_ok(launch_p2_sort(p2_list + first, count_ptr, n, p.hist, L.nestp + 1, p2_sorted + first,
                                        eng->nsm, st), "part-2 sort") ||
                !cuda_ok(cudaMemsetAsync(cnt, 0, sizeof(int), st), "memset"))
                return FPB_ERR_CUDA;
            eng->launches += 2;
            long long g2 = std::min<long long>(grid_p2, ((long long)(n2 - first) + kFitThreads - 1) / kFitThreads);
            if (!cuda_ok(launch_curfit_part2(p, (int)g2, st, p2_sorted + first, count_ptr), "curfit part-2 kernel"))
                return FPB_ERR_CUDA;
            eng->launches++;
        }
    }
    if (isec > 0 && !cuda_ok(cudaStreamWaitEvent(st, eng->ev_p2[1], 0), "join"))   // the caller's stream sees all of it
        return FPB_ERR_CUDA;
    end_timing(eng);
    return FPB_STATUS_OK;
}

}  // namespace fpb

extern "C" {


int32_t fp_curfit_batch_dev_c(fpb_engine *eng, FP_SIZE nbatch, FP_SIZE iopt, FP_SIZE m,
                              const FP_REAL *x, int64_t x_sb, int64_t x_si, const FP_REAL *y,
                              int64_t y_sb, int64_t y_si, const FP_REAL *w, int64_t w_sb,
                              int64_t w_si, const FP_REAL *xb, const FP_REAL *xe, FP_SIZE xbe_stride,
                              FP_SIZE k, const FP_REAL *s, FP_SIZE s_stride, FP_SIZE nest,
                              FP_SIZE *n, FP_REAL *t, FP_REAL *c, FP_REAL *fp, FP_FLAG *ier,
                              FP_REAL *fpint, FP_SIZE *nrdata, FP_SIZE *stats)
{
    if (!eng) {
        set_last_error("null engine");
        return FPB_ERR_ARG;
    }
    if (nbatch <= 0) return FPB_STATUS_OK;
    if (!x || !y || !s || !n || !t || !c || !fp || !ier || (xb && !xe) || (!xb && xe)) {
        set_last_error("fp_curfit_batch_dev_c: null required pointer");
        return FPB_ERR_ARG;
    }
    if (iopt == 1 && (!fpint || !nrdata)) {
        set_last_error("fp_curfit_batch_dev_c: iopt=1 needs the fpint/nrdata state arrays");
        return FPB_ERR_ARG;
    }
    DeviceGuard g(eng->device);
    if (!g.ok) return FPB_ERR_CUDA;
    // call-level part of the curfit data check (core:1265-1268): every curve gets ier=10
    if (k <= 0 || k > 5 || iopt < -1 || iopt > 1 || m < k + 1 || nest < 2 * (k + 1)) {
        if (!cuda_ok(launch_fill_i32(ier, FPB_INPUT_ERROR, nbatch, eng->stream), "fill ier"))
            return FPB_ERR_CUDA;
        eng->launches++;
        return FPB_STATUS_OK;
    }
    FitParams p;
    p.nbatch = nbatch; p.iopt = iopt; p.m = m; p.k = k; p.nest = nest; p.nestp = 0;
    p.x = x; p.y = y; p.w = w;
    p.x_sb = x_sb; p.x_si = x_si; p.y_sb = y_sb; p.y_si = y_si; p.w_sb = w_sb; p.w_si = w_si;
    p.xb = xb; p.xe = xe; p.xbe_stride = xbe_stride;
    p.s = s; p.s_stride = s_stride;
    p.n = n; p.t = t; p.c = c; p.fp = fp; p.ier = ier;
    p.fpint = fpint; p.nrdata = nrdata; p.stats = stats;
    p.idim = 1; p.par = 0; p.periodic = 0; p.y_sd = 0; p.c_sb = nest;
    return run_fit_pipeline(eng, p);
}

int32_t fp_curfit_shared_dev_c(fpb_engine *eng, FP_SIZE nbatch, FP_SIZE m, const FP_REAL *x,
                               const FP_REAL *w, const FP_REAL *y, int64_t y_sb, int64_t y_si,
                               FP_REAL xb, FP_REAL xe, FP_SIZE k, FP_SIZE n, FP_REAL *t, FP_REAL *c,
                               int64_t ldc, FP_REAL *fp, FP_FLAG *ier)
{
    if (!eng || !x || !y || !t || !c || !fp || !ier) {
        set_last_error("fp_curfit_shared_dev_c: null required pointer");
        return FPB_ERR_ARG;
    }
    DeviceGuard g(eng->device);
    if (!g.ok) return FPB_ERR_CUDA;
    if (k <= 0 || k > 5 || m < k + 1 || n < 2 * (k + 1) || ldc < n - k - 1) {
        if (!cuda_ok(launch_fill_i32(ier, FPB_INPUT_ERROR, 1, eng->stream), "fill ier"))
            return FPB_ERR_CUDA;
        eng->launches++;
        return FPB_STATUS_OK;
    }
    const int k1 = k + 1, nk1 = n - k1;
    if (!eng->plan_cs.reserve((size_t)m * k1 * 2 * sizeof(double)) ||
        !eng->plan_w.reserve((size_t)m * sizeof(double)) ||
        !eng->plan_row.reserve((size_t)m * sizeof(int)) ||
        !eng->plan_skip.reserve((size_t)m * sizeof(int)) ||
        !eng->plan_R.reserve((size_t)std::max(nk1, 1) * k1 * sizeof(double)) ||
        !eng->plan_meta.reserve((size_t)m * sizeof(int)) ||
        !eng->plan_ier.reserve(256))
        return FPB_ERR_ALLOC;
    SharedPlan plan;
    plan.cs_sn = eng->plan_cs.as<double>();
    plan.wgt = eng->plan_w.as<double>();
    plan.lrow = eng->plan_row.as<int>();
    plan.skip = eng->plan_skip.as<int>();
    plan.R = eng->plan_R.as<double>();
    plan.ier = ier;
    if (!cuda_ok(launch_shared_plan(x, w, m, xb, xe, k, n, t, plan, eng->stream), "plan kernel"))
        return FPB_ERR_CUDA;
    eng->launches++;
    begin_timing(eng);
    if (nbatch > 0) {
        if (!cuda_ok(launch_shared_apply(nbatch, m, k, n, y, y_sb, y_si, plan, w,
                                         eng->plan_meta.as<int>(), c, ldc, fp, eng->nsm, eng->stream,
                                         eng->shared_fast),
                     "shared apply kernel"))
            return FPB_ERR_CUDA;
        eng->launches += 2;
    }
    end_timing(eng);
    return FPB_STATUS_OK;
}

int32_t fp_splev_batch_dev_c(fpb_engine *eng, FP_SIZE nspl, const FP_REAL *t, const FP_SIZE *n,
                             const FP_REAL *c, FP_SIZE ldt, FP_SIZE k, const FP_REAL *x, FP_REAL *y,
                             FP_SIZE m, FP_FLAG e, FP_SIZE mode, FP_SIZE nshared, FP_FLAG *ier,
                             FP_SIZE *lidx)
{
    if (!eng || !t || !n || !c || !x || !y) {
        set_last_error("fp_splev_batch_dev_c: null required pointer");
        return FPB_ERR_ARG;
    }
    if (nspl <= 0) return FPB_STATUS_OK;
    DeviceGuard g(eng->device);
    if (!g.ok) return FPB_ERR_CUDA;
    if (m < 1) {  // core:17844-17845
        if (ier) {
            if (!cuda_ok(launch_fill_i32(ier, FPB_INPUT_ERROR, nspl, eng->stream), "fill ier"))
                return FPB_ERR_CUDA;
            eng->launches++;
        }
        return FPB_STATUS_OK;
    }
    if (k < 1 || k > 5 || ldt < 2 * (k + 1) || (mode != 0 && mode != 1) || e < 0 || e > 3) {
        set_last_error("fp_splev_batch_dev_c: k must be 1..5, ldt >= 2k+2, mode 0/1, e 0..3");
        return FPB_ERR_ARG;
    }
    NvtxRange nvtx_splev("fpb:splev batch");
    SplevParams p;
    p.nspl = nspl; p.ldt = ldt; p.k = k; p.m = m; p.e = e; p.mode = mode; p.nshared = nshared;
    p.t = t; p.c = c; p.x = x; p.n = n; p.y = y; p.ier = ier; p.lidx = lidx;
    begin_timing(eng);
    if (!cuda_ok(launch_splev(p, eng->nsm, eng->stream), "splev kernel launch")) return FPB_ERR_CUDA;
    end_timing(eng);
    eng->launches++;
    return FPB_STATUS_OK;
}

int32_t fpb_transpose_dev(fpb_engine *eng, const FP_REAL *src, int64_t ld_src, FP_REAL *dst,
                          int64_t ld_dst, int64_t rows, int64_t cols)
{
    if (!eng || !src || !dst) return FPB_ERR_ARG;
    DeviceGuard g(eng->device);
    if (!g.ok) return FPB_ERR_CUDA;
    if (!cuda_ok(launch_transpose(src, ld_src, dst, ld_dst, rows, cols, eng->stream), "transpose"))
        return FPB_ERR_CUDA;
    eng->launches++;
    return FPB_STATUS_OK;
}

int32_t fpb_selftest_fastmath(fpb_engine *eng, int64_t samples, int64_t seed, int64_t *mismatches)
{
    if (!eng || !mismatches || samples <= 0) return FPB_ERR_ARG;
    DeviceGuard g(eng->device);
    if (!g.ok || !eng->counter.reserve(256)) return FPB_ERR_CUDA;
    unsigned long long *d = eng->counter.as<unsigned long long>();
    const int blocks = eng->nsm * 8;
    const long long per_thread = (samples + (long long)blocks * 256 - 1) / ((long long)blocks * 256);
    unsigned long long h[3] = {0, 0, 0};
    if (!cuda_ok(cudaMemsetAsync(d, 0, 24, eng->stream), "memset") ||
        !cuda_ok(launch_fastmath_selftest((unsigned long long)seed, blocks, (int)std::min<long long>(per_thread, 1 << 24), d,
                                          eng->stream), "fastmath self-test") ||
        !cuda_ok(cudaMemcpyAsync(h, d, 24, cudaMemcpyDeviceToHost, eng->stream), "D2H") ||
        !cuda_ok(cudaStreamSynchronize(eng->stream), "sync"))
        return FPB_ERR_CUDA;
    eng->launches++;
    for (int i = 0; i < 3; ++i) mismatches[i] = (int64_t)h[i];
    return FPB_STATUS_OK;
}

double fpb_probe_fp64_gops(fpb_engine *eng, int32_t fused)
{
    if (!eng) return -1.0;
    DeviceGuard g(eng->device);
    if (!g.ok || !eng->counter.reserve(256)) return -1.0;
    if (fused >= 4 && fused <= 6) {   // dependent-issue latency in cycles per instruction (see fpb_util.cu)
        double *d = eng->counter.as<double>() + 8;
        double h = -1.0;
        if (!cuda_ok(launch_fp64_probe(d, 1, 0, fused, eng->stream), "latency probe") ||
            !cuda_ok(cudaMemcpyAsync(&h, d, 8, cudaMemcpyDeviceToHost, eng->stream), "D2H") ||
            !cuda_ok(cudaStreamSynchronize(eng->stream), "sync"))
            return -1.0;
        eng->launches++;
        return h;
    }
    const int blocks = eng->nsm * 8, iters = (fused >= 2) ? (1 << 11) : (1 << 14);
    double best = -1.0;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(eng->ev0, eng->stream);
        if (!cuda_ok(launch_fp64_probe(eng->counter.as<double>() + 8, blocks, iters, fused,
                                       eng->stream), "fp64 probe"))
            return -1.0;
        cudaEventRecord(eng->ev1, eng->stream);
        eng->launches++;
        float ms = 0.f;
        if (cudaEventSynchronize(eng->ev1) != cudaSuccess ||
            cudaEventElapsedTime(&ms, eng->ev0, eng->ev1) != cudaSuccess)
            return -1.0;
        // one FMA counts as one instruction here; mul+add as two
        // one FMA counts as one instruction; mul+add as two; modes 2/3: one div / sqrt per op
        const double ops = (fused >= 2) ? (double)blocks * 256.0 * iters * 4.0
                                        : (double)blocks * 256.0 * iters * 8.0 * (fused ? 1.0 : 2.0);
        best = std::max(best, ops / (ms * 1e-3) * 1e-9);
    }
    eng->timed = false;
    return best;
}

}  // extern "C"

// ------------------------------------------------------------------ host API

namespace {

// One GPU's share of a host-pointer batch fit.  The slice is cut into a few large chunks that
// are software-pipelined over three streams of one engine: while chunk i is fitted on the
// engine's stream, chunk i+1 is copied in (s_in) and the knots/coefficients of chunk i-1 are
// copied out (s_out).  Two staging sets alternate.  Pinned host memory is needed for the copies
// to be asynchronous; with pageable memory the result is the same, only not overlapped.
// Calls for one device are serialised by a per-device mutex.
struct FitStage {
    long long c0 = 0, nb = 0;
    int nmax_used = 0;
    bool pending = false;          // t/c copy-out in flight on s_out
    bool packed = false;           // dense copy-out (iopt >= 0): scatter on the host when it has landed
    long long total = 0;           // doubles in each of the two dense arrays
    std::vector<long long> bad;    // rows rejected with ier=10: the caller's t/c must survive (2-D copy only)
    std::vector<double> keep_t, keep_c;
};

bool reserve_pinned(void **ptr, size_t *have, size_t want)
{
    if (want <= *have) return true;
    if (*ptr) cudaFreeHost(*ptr);
    *ptr = nullptr;
    *have = 0;
    const size_t cap = (want + want / 4 + ((size_t)1 << 20)) & ~(((size_t)1 << 20) - 1);
    if (!cuda_ok(cudaHostAlloc(ptr, cap, cudaHostAllocDefault), "cudaHostAlloc")) return false;
    *have = cap;
    return true;
}

// One host worker's share of a host-pointer batch fit: the chunks `chunk_ids` of the slice (boundaries in
// `cuts`), software-pipelined over three streams of the worker's engine: while chunk i is fitted on the
// engine's stream, chunk i+1 is copied in (s_in) and the knots/coefficients of chunk i-1 are copied out
// (s_out).  Two staging sets alternate.  Knots and coefficients travel back DENSE (pack_tc_kernel: t(1:n),
// c(1:n-k-1) of every fitted curve, 5x fewer bytes than the leading columns of the [nb][nest] arrays at
// config 2) into pinned staging and are scattered into the caller's arrays by this thread.  Pinned caller
// memory makes the input copies asynchronous; with pageable memory the result is the same, only not overlapped.
int32_t fit_slice_pipelined(fpb_engine *eng, const std::vector<long long> &cuts, const std::vector<int> &chunk_ids,
                            long long b0, FP_SIZE iopt,
                            FP_SIZE m, const FP_REAL *x, const FP_REAL *y, const FP_REAL *w,
                            FP_SIZE shared_x, const FP_REAL *xb, const FP_REAL *xe, FP_SIZE k,
                            FP_REAL s, FP_SIZE nest, FP_SIZE *n, FP_REAL *t, FP_REAL *c, FP_REAL *fp,
                            FP_FLAG *ier, FP_REAL *fpint, FP_SIZE *nrdata)
{
    NvtxRange nvtx_slice("fpb:host slice (H2D | transpose+fit | D2H)");
    DeviceGuard g(eng->device);
    if (!g.ok) return FPB_ERR_CUDA;
    if (!eng->s_in) {
        if (!cuda_ok(cudaStreamCreateWithFlags(&eng->s_in, cudaStreamNonBlocking), "stream") ||
            !cuda_ok(cudaStreamCreateWithFlags(&eng->s_out, cudaStreamNonBlocking), "stream"))
            return FPB_ERR_CUDA;
        for (int i = 0; i < 2; ++i)
            if (!cuda_ok(cudaEventCreateWithFlags(&eng->ev_in[i], cudaEventDisableTiming), "event") ||
                !cuda_ok(cudaEventCreateWithFlags(&eng->ev_fit[i], cudaEventDisableTiming), "event"))
                return FPB_ERR_CUDA;
    }
    cudaStream_t st = eng->stream, s_in = eng->s_in, s_out = eng->s_out;
    // FPB_HOST_TRACE=1: wall-clock timeline of the staging pipeline on stderr (tuning aid)
    static const bool trace = std::getenv("FPB_HOST_TRACE") != nullptr;
    static const auto t_epoch = std::chrono::steady_clock::now();
    auto stamp = [&](const char *what, long long ci) {
        if (!trace) return;
        const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_epoch).count();
        std::fprintf(stderr, "[fpb host %p] %9.2f ms  chunk %lld  %s\n", (void *)eng, ms, ci, what);
    };
    const bool state = (fpint && nrdata);
    const bool pack = (iopt >= 0);   // iopt = -1: t is an input with clamped ends, the plain 2-D copy stays
    enum { XC, YC, WC, XP, YP, WP, TT, CC, NN, FP_, IER, AUX, TP, CP, OFF, CUR };
    FitStage stage[2];
    std::future<void> scatter[2];   // host-side scatter of a retired stage's dense t/c
    auto scatter_done = [&](int set) {
        if (scatter[set].valid()) scatter[set].get();
    };
    const long long nchunks = (long long)chunk_ids.size();
    auto bufs = [&](int set) { return set ? eng->scratch2 : eng->scratch; };

    auto copy_in = [&](int set, long long c0, long long nb) -> int32_t {
        DevBuf *B = bufs(set);
        const size_t pts = (size_t)nb * m * sizeof(double);
        const size_t aux_bytes = 64 + (state ? (size_t)nb * nest * 12 : 0);
        if (!B[YC].reserve(pts) || !B[YP].reserve(pts) ||
            !B[XC].reserve(shared_x ? (size_t)m * 8 : pts) || !B[XP].reserve(shared_x ? 8 : pts) ||
            !B[WC].reserve(w ? (shared_x ? (size_t)m * 8 : pts) : 8) ||
            !B[WP].reserve((w && !shared_x) ? pts : 8) || !B[TT].reserve((size_t)nb * nest * 8) ||
            !B[CC].reserve((size_t)nb * nest * 8) || !B[NN].reserve((size_t)nb * 4) ||
            !B[FP_].reserve((size_t)nb * 8) || !B[IER].reserve((size_t)nb * 4) ||
            !B[AUX].reserve(aux_bytes) || !B[OFF].reserve((size_t)nb * 8) || !B[CUR].reserve(64))
            return FPB_ERR_ALLOC;
        double *aux = B[AUX].as<double>();  // [0]=s [1]=xb [2]=xe, then optional state
        eng->h_aux[set][0] = s;
        eng->h_aux[set][1] = xb ? *xb : 0.0;
        eng->h_aux[set][2] = xe ? *xe : 0.0;
        bool ok = cuda_ok(cudaMemcpyAsync(aux, eng->h_aux[set], 24, cudaMemcpyHostToDevice, s_in),
                          "H2D s");
        ok = ok && cuda_ok(cudaMemcpyAsync(B[YC].ptr, y + (size_t)c0 * m, pts, cudaMemcpyHostToDevice,
                                           s_in), "H2D y");
        if (shared_x) {
            ok = ok && cuda_ok(cudaMemcpyAsync(B[XC].ptr, x, (size_t)m * 8, cudaMemcpyHostToDevice,
                                               s_in), "H2D x");
            if (w)
                ok = ok && cuda_ok(cudaMemcpyAsync(B[WC].ptr, w, (size_t)m * 8,
                                                   cudaMemcpyHostToDevice, s_in), "H2D w");
        } else {
            ok = ok && cuda_ok(cudaMemcpyAsync(B[XC].ptr, x + (size_t)c0 * m, pts,
                                               cudaMemcpyHostToDevice, s_in), "H2D x");
            if (w)
                ok = ok && cuda_ok(cudaMemcpyAsync(B[WC].ptr, w + (size_t)c0 * m, pts,
                                                   cudaMemcpyHostToDevice, s_in), "H2D w");
        }
        if (iopt != 0) {  // n, t (and the continuation state) are inputs
            ok = ok && cuda_ok(cudaMemcpyAsync(B[NN].ptr, n + c0, (size_t)nb * 4,
                                               cudaMemcpyHostToDevice, s_in), "H2D n");
            ok = ok && cuda_ok(cudaMemcpyAsync(B[TT].ptr, t + (size_t)c0 * nest,
                                               (size_t)nb * nest * 8, cudaMemcpyHostToDevice, s_in),
                               "H2D t");
            if (iopt == -1)  // a curve rejected by fpchec must get its caller's c back unchanged
                ok = ok && cuda_ok(cudaMemcpyAsync(B[CC].ptr, c + (size_t)c0 * nest,
                                                   (size_t)nb * nest * 8, cudaMemcpyHostToDevice,
                                                   s_in), "H2D c");
            if (iopt == 1 && state) {
                double *dfpint = aux + 8;
                int *dnrd = reinterpret_cast<int *>(dfpint + (size_t)nb * nest);
                ok = ok && cuda_ok(cudaMemcpyAsync(dfpint, fpint + (size_t)c0 * nest,
                                                   (size_t)nb * nest * 8, cudaMemcpyHostToDevice,
                                                   s_in), "H2D fpint");
                ok = ok && cuda_ok(cudaMemcpyAsync(dnrd, nrdata + (size_t)c0 * nest,
                                                   (size_t)nb * nest * 4, cudaMemcpyHostToDevice,
                                                   s_in), "H2D nrdata");
            }
        }
        // outputs the kernel may leave untouched (ier=10) must round-trip unchanged
        ok = ok && cuda_ok(cudaMemcpyAsync(B[FP_].ptr, fp + c0, (size_t)nb * 8,
                                           cudaMemcpyHostToDevice, s_in), "H2D fp");
        if (iopt == 0)
            ok = ok && cuda_ok(cudaMemcpyAsync(B[NN].ptr, n + c0, (size_t)nb * 4,
                                               cudaMemcpyHostToDevice, s_in), "H2D n");
        ok = ok && cuda_ok(cudaEventRecord(eng->ev_in[set], s_in), "event");
        return ok ? FPB_STATUS_OK : FPB_ERR_CUDA;
    };

    // finish the copy-out of a stage: wait for s_out; dense form: scatter into the caller's rows;
    // 2-D form: restore rejected rows
    auto retire = [&](int set) -> bool {
        FitStage &S = stage[set];
        if (!S.pending) return true;
        if (!cuda_ok(cudaStreamSynchronize(s_out), "sync")) return false;
        stamp("t/c copied out (stage retired)", S.c0);
        if (S.packed) {
            // the scatter into the caller's rows (two small strided copies per curve) runs on a helper thread:
            // this thread goes on to launch the next fit
            const double *tp = eng->h_pack[set], *cp = eng->h_pack[set] + S.total;
            const long long *off = eng->h_off[set];
            const long long c0s = S.c0, nbs = S.nb;
            scatter[set] = std::async(std::launch::async, [=]() {
                for (long long b = 0; b < nbs; ++b) {
                    if (off[b] < 0) continue;
                    const int nn = n[c0s + b];
                    std::memcpy(t + (size_t)(c0s + b) * nest, tp + off[b], (size_t)nn * 8);
                    if (nn - k - 1 > 0) std::memcpy(c + (size_t)(c0s + b) * nest, cp + off[b], (size_t)(nn - k - 1) * 8);
                }
            });
        } else {
            for (size_t i = 0; i < S.bad.size(); ++i) {
                std::memcpy(t + (size_t)(S.c0 + S.bad[i]) * nest, &S.keep_t[i * S.nmax_used],
                            (size_t)S.nmax_used * 8);
                std::memcpy(c + (size_t)(S.c0 + S.bad[i]) * nest, &S.keep_c[i * S.nmax_used],
                            (size_t)S.nmax_used * 8);
            }
        }
        S.pending = false;
        return true;
    };

    auto chunk_lo = [&](long long ci) { return b0 + cuts[chunk_ids[ci]]; };
    auto chunk_nb = [&](long long ci) { return cuts[chunk_ids[ci] + 1] - cuts[chunk_ids[ci]]; };
    if (nchunks > 0) {
        const int32_t rc = copy_in(0, chunk_lo(0), chunk_nb(0));
        if (rc != FPB_STATUS_OK) return rc;
    }
    for (long long ci = 0; ci < nchunks; ++ci) {
        const int set = (int)(ci & 1);
        const long long c0 = chunk_lo(ci);
        const long long nb = chunk_nb(ci);
        DevBuf *B = bufs(set);
        // prefetch the next chunk into the other staging set while this one is fitted
        if (ci + 1 < nchunks) {
            if (!retire(1 - set)) return FPB_ERR_CUDA;
            const int32_t rc = copy_in(1 - set, chunk_lo(ci + 1), chunk_nb(ci + 1));
            if (rc != FPB_STATUS_OK) return rc;
        }
        double *aux = B[AUX].as<double>();
        if (trace) {
            cudaEventSynchronize(eng->ev_in[set]);
            stamp("inputs on device", ci);
        }
        bool ok = cuda_ok(cudaStreamWaitEvent(st, eng->ev_in[set], 0), "wait");
        const double *dx, *dw = nullptr;
        long long x_sb, x_si, w_sb = 0, w_si = 0;
        if (shared_x) {
            dx = B[XC].as<double>(); x_sb = 0; x_si = 1;
            if (w) { dw = B[WC].as<double>(); w_sb = 0; w_si = 1; }
        } else {
            ok = ok && cuda_ok(launch_transpose(B[XC].as<double>(), m, B[XP].as<double>(), nb, nb, m,
                                                st), "transpose x");
            eng->launches++;
            dx = B[XP].as<double>(); x_sb = 1; x_si = nb;
            if (w) {
                ok = ok && cuda_ok(launch_transpose(B[WC].as<double>(), m, B[WP].as<double>(), nb,
                                                    nb, m, st), "transpose w");
                eng->launches++;
                dw = B[WP].as<double>(); w_sb = 1; w_si = nb;
            }
        }
        ok = ok && cuda_ok(launch_transpose(B[YC].as<double>(), m, B[YP].as<double>(), nb, nb, m, st),
                           "transpose y");
        eng->launches++;
        if (!ok) return FPB_ERR_CUDA;
        double *dfpint = nullptr;
        int *dnrd = nullptr;
        if (state) {
            dfpint = aux + 8;
            dnrd = reinterpret_cast<int *>(dfpint + (size_t)nb * nest);
        }
        int32_t rc = fp_curfit_batch_dev_c(eng, (FP_SIZE)nb, iopt, m, dx, x_sb, x_si,
                                           B[YP].as<double>(), 1, nb, dw, w_sb, w_si,
                                           xb ? aux + 1 : nullptr, xe ? aux + 2 : nullptr, 0, k, aux,
                                           0, nest, B[NN].as<int>(), B[TT].as<double>(),
                                           B[CC].as<double>(), B[FP_].as<double>(),
                                           B[IER].as<int>(), dfpint, dnrd, nullptr);
        if (rc != FPB_STATUS_OK) return rc;
        FitStage &S = stage[set];
        scatter_done(set);   // the pinned staging of this set is about to be overwritten
        if (pack) {
            // dense t/c: the point-major copies of the inputs (XP/YP) are dead once the fit has run and are
            // always large enough (2 m >= 2 nest doubles per curve is not guaranteed, so reserve explicitly)
            if (!B[TP].reserve((size_t)nb * std::min<long long>(nest, m + k + 1) * 8) ||
                !B[CP].reserve((size_t)nb * std::min<long long>(nest, m + k + 1) * 8) ||
                !reserve_pinned((void **)&eng->h_off[set], &eng->h_off_bytes[set], (size_t)nb * 8))
                return FPB_ERR_ALLOC;
            ok = cuda_ok(cudaMemsetAsync(B[CUR].ptr, 0, 8, st), "memset") &&
                 cuda_ok(launch_pack_tc((int)nb, nest, k, B[NN].as<int>(), B[IER].as<int>(), B[TT].as<double>(),
                                        B[CC].as<double>(), B[CUR].as<unsigned long long>(), B[OFF].as<long long>(),
                                        B[TP].as<double>(), B[CP].as<double>(), eng->nsm, st), "pack kernel") &&
                 cuda_ok(cudaMemcpyAsync(eng->h_off[set], B[OFF].ptr, (size_t)nb * 8, cudaMemcpyDeviceToHost, st),
                         "D2H offsets");
            eng->launches++;
            if (!ok) return FPB_ERR_CUDA;
        }
        ok = cuda_ok(cudaMemcpyAsync(n + c0, B[NN].ptr, (size_t)nb * 4, cudaMemcpyDeviceToHost, st),
                     "D2H n");
        ok = ok && cuda_ok(cudaMemcpyAsync(fp + c0, B[FP_].ptr, (size_t)nb * 8,
                                           cudaMemcpyDeviceToHost, st), "D2H fp");
        ok = ok && cuda_ok(cudaMemcpyAsync(ier + c0, B[IER].ptr, (size_t)nb * 4,
                                           cudaMemcpyDeviceToHost, st), "D2H ier");
        if (!ok || !cuda_ok(cudaStreamSynchronize(st), "sync")) return FPB_ERR_CUDA;
        stamp("fit done, n/fp/ier on host", ci);
        S.c0 = c0;
        S.nb = nb;
        S.bad.clear();
        S.packed = pack;
        if (pack) {
            long long total = 0;
            for (long long b = 0; b < nb; ++b)
                if (eng->h_off[set][b] >= 0) total += n[c0 + b];
            S.total = total;
            if (total > 0) {
                if (!reserve_pinned((void **)&eng->h_pack[set], &eng->h_pack_bytes[set], (size_t)total * 16))
                    return FPB_ERR_ALLOC;
                ok = cuda_ok(cudaMemcpyAsync(eng->h_pack[set], B[TP].ptr, (size_t)total * 8, cudaMemcpyDeviceToHost,
                                             s_out), "D2H t (dense)") &&
                     cuda_ok(cudaMemcpyAsync(eng->h_pack[set] + total, B[CP].ptr, (size_t)total * 8,
                                             cudaMemcpyDeviceToHost, s_out), "D2H c (dense)");
                if (state)
                    ok = ok && cuda_ok(cudaMemcpyAsync(fpint + (size_t)c0 * nest, dfpint, (size_t)nb * nest * 8,
                                                       cudaMemcpyDeviceToHost, s_out), "D2H fpint") &&
                         cuda_ok(cudaMemcpyAsync(nrdata + (size_t)c0 * nest, dnrd, (size_t)nb * nest * 4,
                                                 cudaMemcpyDeviceToHost, s_out), "D2H nrdata");
                if (!ok) return FPB_ERR_CUDA;
                S.pending = true;
            }
            continue;
        }
        // iopt = -1: the leading max(n) columns as a 2-D copy
        int nmax_used = nest;  // clamped end knots may have been written even on ier=10
        S.nmax_used = nmax_used;
        ok = cuda_ok(cudaMemcpy2DAsync(t + (size_t)c0 * nest, (size_t)nest * 8, B[TT].ptr,
                                       (size_t)nest * 8, (size_t)nmax_used * 8, (size_t)nb,
                                       cudaMemcpyDeviceToHost, s_out), "D2H t");
        ok = ok && cuda_ok(cudaMemcpy2DAsync(c + (size_t)c0 * nest, (size_t)nest * 8, B[CC].ptr,
                                             (size_t)nest * 8, (size_t)nmax_used * 8, (size_t)nb,
                                             cudaMemcpyDeviceToHost, s_out), "D2H c");
        if (!ok) return FPB_ERR_CUDA;
        S.pending = true;
    }
    const bool ok_end = retire(0) && retire(1);
    scatter_done(0);
    scatter_done(1);
    stamp("all stages scattered", nchunks);
    return ok_end ? FPB_STATUS_OK : FPB_ERR_CUDA;
}

// How many slices copy through this host at the same time: the ranks the launcher put on this node (torchrun's
// LOCAL_WORLD_SIZE, Open MPI's and Slurm's equivalents, or FPB_LOCAL_RANKS) or, for one process sharding a call
// over its GPUs, the number of slices in flight.  Two decisions hang on it.  (1) Helper threads per slice = cores /
// sharers.  (2) Which host form runs: with more than 2 sharers the host side starts to be the bound (8 GPUs behind one
// host: 23 GB/s H2D and 9 GB/s D2H per GPU instead of 55 / 50) and the wavefront form -- which pays its tail, its
// copy-out and its scatter after the last byte has landed -- loses to the chunk-per-engine form, whose copy-out
// overlaps later chunks: per step of 2^20 curves 168.1 vs 182.5 ms at one GPU, 188.8 vs 182.9 ms at 4, 298.8 vs
// 267.7 ms at 8.  A rate measured on the copy stream was tried as the signal and flapped between forms (ranks are not
// in step during warm-up: 23..54 GB/s on the same box); the launcher's count is deterministic.
std::atomic<int> g_slices_in_flight{0};
int host_sharers()
{
    static const int env_ranks = []() {
        for (const char *name : {"FPB_LOCAL_RANKS", "LOCAL_WORLD_SIZE", "OMPI_COMM_WORLD_LOCAL_SIZE", "SLURM_NTASKS_PER_NODE"})
            if (const char *e = std::getenv(name))
                if (std::atoi(e) > 0) return std::atoi(e);
        return 1;
    }();
    return std::max(env_ranks, std::max(1, g_slices_in_flight.load()));
}

// Does a large fresh fit of `total` curves per slice take the wavefront form (true) or the chunk-per-engine form?
// (fit_slice_host adds the conditions on the call itself: iopt = 0, own abscissae, no continuation state, memory.)
bool wavefront_preferred(long long total)
{
    static const char *mode = std::getenv("FPB_HOST_MODE");
    if (total < 4 * 65536 || total > 0x3fffffff) return false;
    if (mode && std::strcmp(mode, "workers") == 0) return false;
    if (mode && std::strcmp(mode, "wave") == 0) return true;
    return host_sharers() <= 2;   // see host_sharers()
}

// Wavefront form of the host-pointer batch fit (large fresh fits, iopt = 0): ONE pipeline over the whole slice
// on one engine.  The inputs stream in as chunks on a copy stream (H2D into two alternating staging sets, tile
// transpose into the slice-wide point-major arrays); each chunk's curves join the running pipeline at the first
// launch after they have landed (run_fit_pipeline, staged mode), so the GPU works from the first chunk on and
// the latency-bound tail of the fit (the sparsely populated last passes, part 2 of the curves with ~100 knots)
// is paid ONCE per slice instead of once per chunk.  Results come back dense (pack_tc_kernel) and are scattered
// into the caller's rows by a few helper threads.
int32_t fit_slice_wavefront(fpb_engine *eng, long long b0, long long b1, FP_SIZE m, const FP_REAL *x,
                            const FP_REAL *y, const FP_REAL *w, const FP_REAL *xb, const FP_REAL *xe, FP_SIZE k,
                            FP_REAL s, FP_SIZE nest, FP_SIZE *n, FP_REAL *t, FP_REAL *c, FP_REAL *fp, FP_FLAG *ier)
{
    NvtxRange nvtx_slice("fpb:host slice, wavefront (H2D chunks | one pipeline | dense D2H)");
    DeviceGuard g(eng->device);
    if (!g.ok) return FPB_ERR_CUDA;
    if (!eng->s_in) {
        if (!cuda_ok(cudaStreamCreateWithFlags(&eng->s_in, cudaStreamNonBlocking), "stream") ||
            !cuda_ok(cudaStreamCreateWithFlags(&eng->s_out, cudaStreamNonBlocking), "stream"))
            return FPB_ERR_CUDA;
        for (int i = 0; i < 2; ++i)
            if (!cuda_ok(cudaEventCreateWithFlags(&eng->ev_in[i], cudaEventDisableTiming), "event") ||
                !cuda_ok(cudaEventCreateWithFlags(&eng->ev_fit[i], cudaEventDisableTiming), "event"))
                return FPB_ERR_CUDA;
    }
    cudaStream_t st = eng->stream, s_in = eng->s_in;
    static const bool trace = std::getenv("FPB_HOST_TRACE") != nullptr;
    const auto t_epoch = std::chrono::steady_clock::now();
    auto stamp = [&](const char *what) {
        if (!trace) return;
        const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_epoch).count();
        std::fprintf(stderr, "[fpb wavefront %p] %9.2f ms  %s\n", (void *)eng, ms, what);
    };
    const long long total = b1 - b0;
    // chunks: small first (the GPU starts early), then ~1/8 of the slice each
    int nchunk = 10;
    if (const char *e = std::getenv("FPB_HOST_WAVE_CHUNKS")) nchunk = std::max(2, std::atoi(e));
    std::vector<long long> cuts;
    cuts.push_back(0);
    {
        const long long unit = std::max<long long>(128, (total / nchunk + 127) / 128 * 128);
        long long first = std::max<long long>(128, (unit / 4 + 127) / 128 * 128);
        cuts.push_back(std::min(total, first));
        cuts.push_back(std::min(total, first + unit / 2 / 128 * 128 + 128));
        while (cuts.back() < total) cuts.push_back(std::min(total, cuts.back() + unit));
    }
    const int nc = (int)cuts.size() - 1;
    long long maxchunk = 0;
    for (int i = 0; i < nc; ++i) maxchunk = std::max(maxchunk, cuts[i + 1] - cuts[i]);
    enum { XC0, YC0, WC0, XC1, YC1, WC1, XP, YP, WP, TT, CC, NN, FP_, IER, AUX, OFF };   // eng->scratch
    enum { TP, CP, CUR };                                                                  // eng->scratch2
    DevBuf *B = eng->scratch, *B2 = eng->scratch2;
    const size_t cpts = (size_t)maxchunk * m * 8, pts = (size_t)total * m * 8;
    const long long dense_cap = total * std::min<long long>(nest, m + k + 1);
    if (!B[XC0].reserve(cpts) || !B[YC0].reserve(cpts) || !B[XC1].reserve(cpts) || !B[YC1].reserve(cpts) ||
        !B[WC0].reserve(w ? cpts : 8) || !B[WC1].reserve(w ? cpts : 8) || !B[XP].reserve(pts) || !B[YP].reserve(pts) ||
        !B[WP].reserve(w ? pts : 8) || !B[TT].reserve((size_t)total * nest * 8) || !B[CC].reserve((size_t)total * nest * 8) ||
        !B[NN].reserve((size_t)total * 4) || !B[FP_].reserve((size_t)total * 8) || !B[IER].reserve((size_t)total * 4) ||
        !B[AUX].reserve(64) || !B[OFF].reserve((size_t)total * 8) || !B2[TP].reserve((size_t)dense_cap * 8) ||
        !B2[CP].reserve((size_t)dense_cap * 8) || !B2[CUR].reserve(64) ||
        !reserve_pinned((void **)&eng->h_off[0], &eng->h_off_bytes[0], (size_t)total * 8))
        return FPB_ERR_ALLOC;
    // events, one per chunk (grow-only pool on the engine)
    while ((int)eng->ev_chunk.size() < nc) {
        cudaEvent_t e = nullptr;
        if (!cuda_ok(cudaEventCreateWithFlags(&e, cudaEventDisableTiming), "event")) return FPB_ERR_CUDA;
        eng->ev_chunk.push_back(e);
    }
    // everything the copy stream will do, enqueued up front (asynchronous for pinned caller memory)
    double *aux = B[AUX].as<double>();
    eng->h_aux[0][0] = s;
    eng->h_aux[0][1] = xb ? *xb : 0.0;
    eng->h_aux[0][2] = xe ? *xe : 0.0;
    bool ok = cuda_ok(cudaMemcpyAsync(aux, eng->h_aux[0], 24, cudaMemcpyHostToDevice, st), "H2D s");
    // outputs the kernel may leave untouched (ier = 10) round-trip unchanged
    ok = ok && cuda_ok(cudaMemcpyAsync(B[FP_].ptr, fp + b0, (size_t)total * 8, cudaMemcpyHostToDevice, st), "H2D fp");
    ok = ok && cuda_ok(cudaMemcpyAsync(B[NN].ptr, n + b0, (size_t)total * 4, cudaMemcpyHostToDevice, st), "H2D n");
    ok = ok && cuda_ok(cudaEventRecord(eng->ev_in[0], st), "event");
    ok = ok && cuda_ok(cudaStreamWaitEvent(s_in, eng->ev_in[0], 0), "wait");   // the copy stream starts after earlier work on st
    if (!eng->ev_h2d[0]) {
        ok = ok && cuda_ok(cudaEventCreate(&eng->ev_h2d[0]), "event") && cuda_ok(cudaEventCreate(&eng->ev_h2d[1]), "event");
    }
    ok = ok && cuda_ok(cudaEventRecord(eng->ev_h2d[0], s_in), "event");
    std::vector<FitArrival> arrivals;
    for (int ci = 0; ok && ci < nc; ++ci) {
        const long long c0 = cuts[ci], nb = cuts[ci + 1] - cuts[ci];
        const int set = ci & 1;
        double *xc = B[set ? XC1 : XC0].as<double>(), *yc = B[set ? YC1 : YC0].as<double>(),
               *wc = B[set ? WC1 : WC0].as<double>();
        const size_t bytes = (size_t)nb * m * 8;
        ok = ok && cuda_ok(cudaMemcpyAsync(xc, x + (size_t)(b0 + c0) * m, bytes, cudaMemcpyHostToDevice, s_in), "H2D x");
        ok = ok && cuda_ok(cudaMemcpyAsync(yc, y + (size_t)(b0 + c0) * m, bytes, cudaMemcpyHostToDevice, s_in), "H2D y");
        if (w) ok = ok && cuda_ok(cudaMemcpyAsync(wc, w + (size_t)(b0 + c0) * m, bytes, cudaMemcpyHostToDevice, s_in), "H2D w");
        // curve-major chunk -> columns [c0, c0+nb) of the slice-wide point-major arrays (ld = total)
        ok = ok && cuda_ok(launch_transpose(xc, m, B[XP].as<double>() + c0, total, nb, m, s_in), "transpose x");
        ok = ok && cuda_ok(launch_transpose(yc, m, B[YP].as<double>() + c0, total, nb, m, s_in), "transpose y");
        if (w) ok = ok && cuda_ok(launch_transpose(wc, m, B[WP].as<double>() + c0, total, nb, m, s_in), "transpose w");
        eng->launches += w ? 3 : 2;
        ok = ok && cuda_ok(cudaEventRecord(eng->ev_chunk[ci], s_in), "event");
        arrivals.push_back(FitArrival{(int)c0, (int)nb, eng->ev_chunk[ci]});
    }
    ok = ok && cuda_ok(cudaEventRecord(eng->ev_h2d[1], s_in), "event");
    if (!ok) return FPB_ERR_CUDA;
    stamp("copies enqueued");
    FitParams p;
    std::memset(&p, 0, sizeof p);
    p.nbatch = (int)total; p.iopt = 0; p.m = m; p.k = k; p.nest = nest;
    p.x = B[XP].as<double>(); p.y = B[YP].as<double>(); p.w = w ? B[WP].as<double>() : nullptr;
    p.x_sb = 1; p.x_si = total; p.y_sb = 1; p.y_si = total; p.w_sb = w ? 1 : 0; p.w_si = w ? total : 0;
    p.xb = xb ? aux + 1 : nullptr; p.xe = xe ? aux + 2 : nullptr; p.xbe_stride = 0;
    p.s = aux; p.s_stride = 0;
    p.n = B[NN].as<int>(); p.t = B[TT].as<double>(); p.c = B[CC].as<double>(); p.fp = B[FP_].as<double>();
    p.ier = B[IER].as<int>();
    p.idim = 1; p.c_sb = nest;
    const int32_t rc = run_fit_pipeline(eng, p, &arrivals);
    if (rc != FPB_STATUS_OK) return rc;
    stamp("pipeline enqueued");
    // dense copy-out
    ok = cuda_ok(cudaMemsetAsync(B2[CUR].ptr, 0, 8, st), "memset") &&
         cuda_ok(launch_pack_tc((int)total, nest, k, B[NN].as<int>(), B[IER].as<int>(), B[TT].as<double>(),
                                B[CC].as<double>(), B2[CUR].as<unsigned long long>(), B[OFF].as<long long>(),
                                B2[TP].as<double>(), B2[CP].as<double>(), eng->nsm, st), "pack kernel") &&
         cuda_ok(cudaMemcpyAsync(eng->h_off[0], B[OFF].ptr, (size_t)total * 8, cudaMemcpyDeviceToHost, st), "D2H offsets") &&
         cuda_ok(cudaMemcpyAsync(n + b0, B[NN].ptr, (size_t)total * 4, cudaMemcpyDeviceToHost, st), "D2H n") &&
         cuda_ok(cudaMemcpyAsync(fp + b0, B[FP_].ptr, (size_t)total * 8, cudaMemcpyDeviceToHost, st), "D2H fp") &&
         cuda_ok(cudaMemcpyAsync(ier + b0, B[IER].ptr, (size_t)total * 4, cudaMemcpyDeviceToHost, st), "D2H ier") &&
         cuda_ok(cudaStreamSynchronize(st), "sync");
    eng->launches++;
    if (!ok) return FPB_ERR_CUDA;
    stamp("fit done, n/fp/ier/offsets on host");
    if (trace) {   // what the copy stream achieved (all chunks have landed: the pipeline consumed them)
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, eng->ev_h2d[0], eng->ev_h2d[1]) == cudaSuccess && ms > 0.f)
            std::fprintf(stderr, "[fpb wavefront %p] copy stream: %.2f ms, %.1f GB/s\n", (void *)eng, ms,
                         (double)total * m * 8.0 * (w ? 3 : 2) / (ms * 1e-3) / 1e9);
    }
    const long long *off = eng->h_off[0];
    long long dense = 0;
    for (long long b = 0; b < total; ++b)
        if (off[b] >= 0) dense += n[b0 + b];
    if (dense > 0) {
        if (!reserve_pinned((void **)&eng->h_pack[0], &eng->h_pack_bytes[0], (size_t)dense * 16)) return FPB_ERR_ALLOC;
        double *tp = eng->h_pack[0], *cp = eng->h_pack[0] + dense;
        // the dense arrays come back in pieces; a helper thread scatters each piece's curves as soon as it has
        // landed -- but the curves of a piece are not contiguous in b (allocation order of the pack kernel), so
        // every helper walks its own range of b and waits for the piece that holds the curve's data
        const int npiece = 8;
        std::vector<cudaEvent_t> evp(npiece);
        std::vector<long long> pend(npiece + 1, 0);
        for (int q = 0; q <= npiece; ++q) pend[q] = dense * q / npiece;
        for (int q = 0; ok && q < npiece; ++q) {
            ok = cuda_ok(cudaEventCreateWithFlags(&evp[q], cudaEventDisableTiming | cudaEventBlockingSync), "event");
            const long long lo = pend[q], cnt_ = pend[q + 1] - pend[q];
            if (cnt_ > 0) {
                ok = ok && cuda_ok(cudaMemcpyAsync(tp + lo, B2[TP].as<double>() + lo, (size_t)cnt_ * 8, cudaMemcpyDeviceToHost, st), "D2H t (dense)");
                ok = ok && cuda_ok(cudaMemcpyAsync(cp + lo, B2[CP].as<double>() + lo, (size_t)cnt_ * 8, cudaMemcpyDeviceToHost, st), "D2H c (dense)");
            }
            ok = ok && cuda_ok(cudaEventRecord(evp[q], st), "event");
        }
        if (!ok) return FPB_ERR_CUDA;
        // helper threads: one per core this slice may fairly claim (cores / host_sharers()), 2..16; they take blocks
        // of 2048 curves in order from a shared counter, so all of them work on the pieces that landed first
        int nthr = (int)std::thread::hardware_concurrency() / host_sharers();
        nthr = std::max(2, std::min(16, nthr));
        if (const char *e = std::getenv("FPB_HOST_SCATTER_THREADS")) nthr = std::max(1, std::min(32, std::atoi(e)));
        std::vector<std::thread> th;
        const int dev = eng->device;
        std::atomic<long long> next_block{0};
        const long long blk = 2048;
        for (int wk = 0; wk < nthr; ++wk) {
            th.emplace_back([&]() {
                cudaSetDevice(dev);
                int reached = -1;   // pieces known to have landed
                for (;;) {
                    const long long lo = next_block.fetch_add(blk);
                    if (lo >= total) break;
                    const long long hi = std::min(total, lo + blk);
                    for (long long b = lo; b < hi; ++b) {
                        if (off[b] < 0) continue;
                        const int nn = n[b0 + b];
                        const long long last = off[b] + nn - 1;
                        int need = 0;
                        while (need < npiece - 1 && last >= pend[need + 1]) ++need;
                        while (reached < need) cudaEventSynchronize(evp[++reached]);
                        std::memcpy(t + (size_t)(b0 + b) * nest, tp + off[b], (size_t)nn * 8);
                        if (nn - k - 1 > 0) std::memcpy(c + (size_t)(b0 + b) * nest, cp + off[b], (size_t)(nn - k - 1) * 8);
                    }
                }
            });
        }
        for (auto &tt : th) tt.join();
        ok = cuda_ok(cudaStreamSynchronize(st), "sync");
        for (auto e : evp) cudaEventDestroy(e);
        if (!ok) return FPB_ERR_CUDA;
    }
    stamp("t/c scattered");
    return FPB_STATUS_OK;
}

// Chunking of one GPU's slice.  The fit pipeline is most efficient on big batches (its tail -- the last
// knot sets and the longest p-iterations -- is a chain of latency-bound launches), while the first bytes
// should reach the GPU quickly and the last copy-out should be short.  The slice is cut into `nchunks`
// pieces dealt round-robin to `workers` host threads, each with its own engine (stream, workspace,
// staging): the copies of one chunk and the sparsely populated tail of its fit overlap the other workers'
// compute.  FPB_HOST_WORKERS (1..4) / FPB_HOST_CHUNKS override the defaults.
// fpchec, core:2507-2570 (host form: the flat entry point has the knots and abscissae in host memory)
int host_fpchec(const double *x, int m, const double *t, int n, int k)
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

// A FEW curves from host memory -- one (fp_curfit_c and, through it, the handle API) up to kSoloMaxBatch
// (fp_curfit_batch_c) -- on the warp-per-curve kernel (fpb_curfit_solo.cu), one CTA of one warp per curve: where the
// one-thread-per-curve pipeline would leave the GPU to a handful of threads.  curfit's data checks (core:1265-1285)
// run on the host, where the data is.  One packed block for a single curve (one H2D, one launch, one D2H); for more,
// x / y / w go up straight from the caller's arrays.  Returns false when the call is not one for this kernel (too
// large for its shared memory, shared abscissae, ...): the caller then takes the batch path.
constexpr long long kSoloMaxBatch = 4096;   // measured: 2048 curves 7.4 ms here, 22 ms on the batch kernels (profiles/r02_solo.md)

bool fit_solo_host(fpb_engine *eng, long long nb, FP_SIZE iopt, FP_SIZE m, const FP_REAL *x, const FP_REAL *y,
                   const FP_REAL *w, const FP_REAL *xbp, const FP_REAL *xep, FP_SIZE k, FP_REAL s, FP_SIZE nest,
                   FP_SIZE *n, FP_REAL *t, FP_REAL *c, FP_REAL *fp, FP_FLAG *ier, FP_REAL *fpint, FP_SIZE *nrdata,
                   int32_t *status)
{
    if (!g_solo_on.load() || nb < 1 || nb > kSoloMaxBatch || m < 1 || solo_smem_bytes(m, k, nest, 1) > kSoloSmemMax)
        return false;
    if (iopt == 1 && (!fpint || !nrdata)) return false;
    *status = FPB_STATUS_OK;
    const int k1 = k + 1, nmin = 2 * k1, nmax = m + k1;
    // ---- packed block, in doubles: x y w | fp n ier | t | c | fpint nrdata ----
    const size_t pts = (size_t)nb * m, kn = (size_t)nb * nest, half = ((size_t)nb + 1) / 2;
    const size_t o_fp = 3 * pts, o_n = o_fp + nb, o_ier = o_n + half, o_t = o_ier + half, o_c = o_t + kn,
                 o_f = o_c + kn, o_nrd = o_f + kn, o_end = o_nrd + (kn + 1) / 2;
    DeviceGuard g(eng->device);
    auto fail = [&](int32_t rc) {
        *status = rc;
        for (long long b = 0; b < nb; ++b) ier[b] = rc;
        return true;
    };
    if (!g.ok) return fail(FPB_ERR_CUDA);
    // the pinned mirror of the block starts where the host needs it: a single curve is packed whole; for more, x and y
    // (and w) go up straight from the caller's arrays
    const size_t h0 = (nb == 1) ? 0 : (w ? o_fp : 2 * pts);
    if (!eng->scratch[0].reserve(o_end * 8) ||
        !reserve_pinned((void **)&eng->h_solo, &eng->h_solo_bytes, (o_end - h0) * 8))
        return fail(FPB_ERR_ALLOC);
    double *hd = eng->h_solo - h0, *dd = eng->scratch[0].as<double>();   // hd + offset is valid for offsets >= h0
    int *hn = reinterpret_cast<int *>(hd + o_n), *hier = reinterpret_cast<int *>(hd + o_ier),
        *hnrd = reinterpret_cast<int *>(hd + o_nrd);
    // ---- curfit's checks on the data, core:1271-1285, curve by curve ----
    long long nbad = 0;
    for (long long b = 0; b < nb; ++b) {
        const double *xc = x + (size_t)b * m;
        const double xb = xbp ? *xbp : xc[0], xe = xep ? *xep : xc[m - 1];
        bool bad = (xb > xc[0] || xe < xc[m - 1]);
        for (int i = 0; i + 1 < m && !bad; ++i) bad = xc[i] > xc[i + 1];
        int n_in = 0;
        int flag = FPB_OK;
        if (!bad && iopt >= 0) {
            bad = s < 0.0 || (std::fabs(s) < DBL_MIN && nest < m + k1);
            if (iopt == 1) {
                n_in = n[b];
                if (!(n_in >= nmin && n_in <= nest && n_in <= nmax)) n_in = nmin;   // garbage n: a fresh start
            }
        } else if (!bad) {
            n_in = n[b];
            bad = n_in < nmin || n_in > nest;
            if (!bad) {
                double *tc = t + (size_t)b * nest;
                for (int i = 0, j = n_in - 1; i < k1; ++i, --j) {   // visible even if fpchec fails (core:1278-1284)
                    tc[i] = xb;
                    tc[j] = xe;
                }
                flag = host_fpchec(xc, m, tc, n_in, k);
            }
        }
        if (bad) flag = FPB_INPUT_ERROR;
        hier[b] = flag;           // a curve that arrives flagged is skipped by the kernel
        hn[b] = n_in;
        hd[o_fp + b] = fp[b];
        if (flag != FPB_OK) ++nbad;
        if (flag == FPB_OK && iopt != 0) std::memcpy(hd + o_t + (size_t)b * nest, t + (size_t)b * nest, (size_t)n_in * 8);
        if (flag == FPB_OK && iopt == 1) {
            std::memcpy(hd + o_f + (size_t)b * nest, fpint + (size_t)b * nest, (size_t)n_in * 8);
            std::memcpy(hnrd + (size_t)b * nest, nrdata + (size_t)b * nest, (size_t)n_in * 4);
        }
    }
    if (nbad == nb) {   // nothing to fit
        for (long long b = 0; b < nb; ++b) ier[b] = hier[b];
        return true;
    }
    cudaStream_t st = eng->stream;
    bool ok = true;
    if (nb == 1) {
        std::memcpy(hd, x, pts * 8);
        std::memcpy(hd + pts, y, pts * 8);
        if (w) std::memcpy(hd + 2 * pts, w, pts * 8);
        else std::fill(hd + 2 * pts, hd + 3 * pts, 1.0);
        ok = cuda_ok(cudaMemcpyAsync(dd, hd, (iopt == 1 ? o_end : o_c) * 8, cudaMemcpyHostToDevice, st), "H2D curve");
    } else {
        ok = cuda_ok(cudaMemcpyAsync(dd, x, pts * 8, cudaMemcpyHostToDevice, st), "H2D x") &&
             cuda_ok(cudaMemcpyAsync(dd + pts, y, pts * 8, cudaMemcpyHostToDevice, st), "H2D y");
        if (w) {
            ok = ok && cuda_ok(cudaMemcpyAsync(dd + 2 * pts, w, pts * 8, cudaMemcpyHostToDevice, st), "H2D w");
        } else {
            std::fill(hd + 2 * pts, hd + 3 * pts, 1.0);
            ok = ok && cuda_ok(cudaMemcpyAsync(dd + 2 * pts, hd + 2 * pts, pts * 8, cudaMemcpyHostToDevice, st), "H2D w");
        }
        ok = ok && cuda_ok(cudaMemcpyAsync(dd + o_fp, hd + o_fp, ((iopt != 0 ? o_c : o_t) - o_fp) * 8,
                                           cudaMemcpyHostToDevice, st), "H2D state");
        if (iopt == 1)
            ok = ok && cuda_ok(cudaMemcpyAsync(dd + o_f, hd + o_f, (o_end - o_f) * 8, cudaMemcpyHostToDevice, st), "H2D state");
    }
    SoloParams p;
    std::memset(&p, 0, sizeof p);
    p.nbatch = (int)nb; p.iopt = iopt; p.m = m; p.k = k; p.nest = nest;
    p.idim = 1; p.c_stride = nest;
    p.xb = xbp ? *xbp : 0.0; p.xe = xep ? *xep : 0.0; p.s = s;
    p.ends_from_x = (xbp && xep) ? 0 : 1;
    if (!xbp != !xep) {   // one end given, the other from the data: not a case the batch API produces
        p.ends_from_x = 0;
        if (nb != 1) return false;
        p.xb = xbp ? *xbp : x[0];
        p.xe = xep ? *xep : x[m - 1];
    }
    p.x = dd; p.y = dd + pts; p.w = dd + 2 * pts;
    p.fp = dd + o_fp; p.n = reinterpret_cast<int *>(dd + o_n); p.ier = reinterpret_cast<int *>(dd + o_ier);
    p.t = dd + o_t; p.c = dd + o_c; p.fpint = dd + o_f; p.nrdata = reinterpret_cast<int *>(dd + o_nrd);
    const bool want_state = fpint && nrdata;
    ok = ok && cuda_ok(launch_curfit_solo(p, st), "curfit_solo_kernel") &&
         cuda_ok(cudaMemcpyAsync(hd + o_fp, dd + o_fp, ((want_state ? o_end : o_f) - o_fp) * 8, cudaMemcpyDeviceToHost, st),
                 "D2H fit") &&
         cuda_ok(cudaStreamSynchronize(st), "sync");
    eng->launches++;
    if (!ok) return fail(FPB_ERR_CUDA);
    const bool untouched_all = (iopt >= 0 && s <= 0.0 && nmax > nest);   // ier = 1 before anything was computed
    for (long long b = 0; b < nb; ++b) {
        const int ier_out = hier[b], n_out = hn[b];
        ier[b] = ier_out;
        if (ier_out == FPB_INPUT_ERROR) continue;            // rejected: n, t, c, fp stay the caller's
        n[b] = n_out;
        if (untouched_all || n_out < nmin || n_out > nest) continue;
        fp[b] = hd[o_fp + b];
        std::memcpy(t + (size_t)b * nest, hd + o_t + (size_t)b * nest, (size_t)n_out * 8);
        std::memcpy(c + (size_t)b * nest, hd + o_c + (size_t)b * nest, (size_t)(n_out - k1) * 8);
        if (want_state) {
            std::memcpy(fpint + (size_t)b * nest, hd + o_f + (size_t)b * nest, (size_t)n_out * 8);
            std::memcpy(nrdata + (size_t)b * nest, hnrd + (size_t)b * nest, (size_t)n_out * 4);
        }
    }
    return true;
}

int32_t fit_slice_host(int device, long long b0, long long b1, FP_SIZE iopt, FP_SIZE m,
                       const FP_REAL *x, const FP_REAL *y, const FP_REAL *w, FP_SIZE shared_x,
                       const FP_REAL *xb, const FP_REAL *xe, FP_SIZE k, FP_REAL s, FP_SIZE nest,
                       FP_SIZE *n, FP_REAL *t, FP_REAL *c, FP_REAL *fp, FP_FLAG *ier,
                       FP_REAL *fpint, FP_SIZE *nrdata)
{
    if (device < 0 && !cuda_ok(cudaGetDevice(&device), "cudaGetDevice")) return FPB_ERR_CUDA;
    if (device >= 32) return FPB_ERR_ARG;
    HostApiLock lk(device);
    fpb_engine *e0 = default_engine(device);
    if (!e0) return FPB_ERR_CUDA;
    const long long total = b1 - b0;
    if (total <= kSoloMaxBatch && !shared_x && iopt != 1) {   // a few curves: the warp-per-curve kernel
        int32_t status = FPB_STATUS_OK;
        if (fit_solo_host(e0, total, iopt, m, x + (size_t)b0 * m, y + (size_t)b0 * m, w ? w + (size_t)b0 * m : nullptr,
                          xb, xe, k, s, nest, n + b0, t + (size_t)b0 * nest, c + (size_t)b0 * nest, fp + b0, ier + b0,
                          fpint ? fpint + (size_t)b0 * nest : nullptr, nrdata ? nrdata + (size_t)b0 * nest : nullptr,
                          &status))
            return status;
    } else if (total == 1 && !shared_x) {   // continuation of one curve (fp_curfit_c with iopt = 1)
        int32_t status = FPB_STATUS_OK;
        if (fit_solo_host(e0, 1, iopt, m, x + (size_t)b0 * m, y + (size_t)b0 * m, w ? w + (size_t)b0 * m : nullptr,
                          xb, xe, k, s, nest, n + b0, t + (size_t)b0 * nest, c + (size_t)b0 * nest, fp + b0, ier + b0,
                          fpint ? fpint + (size_t)b0 * nest : nullptr, nrdata ? nrdata + (size_t)b0 * nest : nullptr,
                          &status))
            return status;
    }
    // large fresh fits: one pipeline over the slice, chunks joining as they arrive (FPB_HOST_MODE=workers: the
    // chunk-per-engine form below)
    {
        const size_t per_curve_w = (size_t)m * 8 * 2 * (w ? 3 : 2) + (size_t)nest * 8 * 4 + 64;   // slice-wide device bytes per curve
        if (wavefront_preferred(total) && iopt == 0 && !shared_x && !fpint && !nrdata &&
            (size_t)total * per_curve_w <= ((size_t)40 << 30))
            return fit_slice_wavefront(e0, b0, b1, m, x, y, w, xb, xe, k, s, nest, n, t, c, fp, ier);
    }
    int workers = 4;   // measured (profiles/r02_e2e.md): 2 -> 209 ms, 3 -> 213 ms, 4 -> 184 ms, 6 -> 190 ms per 2^20 curves
    if (const char *e = std::getenv("FPB_HOST_WORKERS")) workers = std::max(1, std::min(8, std::atoi(e)));
    // bytes per curve on the device: x,y,w curve-major + point-major, outputs t,c (+state, dense copies)
    const size_t per_curve = (size_t)m * 8 * 6 + (size_t)nest * 8 * 4 + (size_t)nest * 12 + 64;
    const long long cap = std::max<long long>(128, (long long)(((size_t)10 << 30) / per_curve) / 128 * 128);  // <= 10 GB per staging set
    // Plan: one SMALL chunk per worker first (2/32, 3/32, 4/32, 5/32 of the slice: the first bytes reach the
    // GPU after a few ms and every engine has work early), then one large chunk per worker with the rest
    // (more if a chunk would exceed the staging cap).  FPB_HOST_PLAN="f1,f2,..." overrides the fractions.
    std::vector<double> frac;
    if (const char *e = std::getenv("FPB_HOST_PLAN")) {
        for (const char *q = e; *q;) {
            char *end = nullptr;
            const double v = std::strtod(q, &end);
            if (end == q) break;
            if (v > 0) frac.push_back(v);
            q = (*end == ',') ? end + 1 : end;
        }
    }
    if (total < 4 * 65536) {
        workers = 1;
        frac.assign(1, 1.0);
    } else if (frac.empty()) {
        double used = 0.0;
        for (int wk = 0; wk < workers; ++wk) {
            frac.push_back((2.0 + wk) / 32.0);
            used += frac.back();
        }
        for (int wk = 0; wk < workers; ++wk) frac.push_back((1.0 - used) / workers);
    }
    std::vector<long long> cuts;
    cuts.push_back(0);
    for (size_t i = 0; i < frac.size() && cuts.back() < total; ++i) {
        long long want = (i + 1 == frac.size()) ? total - cuts.back() : (long long)(frac[i] * (double)total);
        want = std::max<long long>(128, (want + 127) / 128 * 128);
        while (want > 0 && cuts.back() < total) {   // split what exceeds the staging cap
            const long long piece = std::min(std::min(want, cap), total - cuts.back());
            cuts.push_back(cuts.back() + piece);
            want -= piece;
        }
    }
    while (cuts.back() < total) cuts.push_back(std::min(total, cuts.back() + cap));
    const int nc = (int)cuts.size() - 1;
    workers = std::min(workers, nc);
    std::vector<std::vector<int>> ids(workers);
    for (int ci = 0; ci < nc; ++ci) ids[ci % workers].push_back(ci);
    if (workers == 1)
        return fit_slice_pipelined(e0, cuts, ids[0], b0, iopt, m, x, y, w, shared_x, xb, xe, k, s, nest, n, t, c, fp,
                                   ier, fpint, nrdata);
    std::vector<fpb_engine *> engs(workers);
    for (int wk = 0; wk < workers; ++wk) {
        engs[wk] = default_engine(device + 32 * wk);  // engines 1..3 of the same device
        if (!engs[wk]) return FPB_ERR_CUDA;
    }
    std::vector<int32_t> rc(workers, FPB_STATUS_OK);
    std::vector<std::string> err(workers);
    std::vector<std::thread> th;
    for (int wk = 0; wk < workers; ++wk) {
        th.emplace_back([&, wk]() {
            rc[wk] = fit_slice_pipelined(engs[wk], cuts, ids[wk], b0, iopt, m, x, y, w, shared_x, xb, xe, k, s, nest,
                                         n, t, c, fp, ier, fpint, nrdata);
            if (rc[wk] != FPB_STATUS_OK) err[wk] = fpb_last_error();
        });
    }
    for (auto &tt : th) tt.join();
    for (int wk = 0; wk < workers; ++wk)
        if (rc[wk] != FPB_STATUS_OK) {
            set_last_error(err[wk]);
            return rc[wk];
        }
    return FPB_STATUS_OK;
}

int32_t splev_slice_host(fpb_engine *eng, long long j0, long long j1, const FP_REAL *t,
                         const FP_SIZE *n, const FP_REAL *c, FP_SIZE ldt, FP_SIZE k,
                         const FP_REAL *x, FP_REAL *y, FP_SIZE m, FP_FLAG e, FP_SIZE mode,
                         FP_FLAG *ier)
{
    DeviceGuard g(eng->device);
    if (!g.ok) return FPB_ERR_CUDA;
    cudaStream_t st = eng->stream;
    const size_t per_spl = (size_t)m * 16 + (size_t)ldt * 16 + 8;
    long long chunk = std::max<long long>((long long)(((size_t)8 << 30) / per_spl), 1);
    enum { TT, CC, NN, XX, YY, IER };
    for (long long c0 = j0; c0 < j1; c0 += chunk) {
        const long long ns = std::min(chunk, j1 - c0);
        DevBuf *B = eng->scratch;
        if (!B[TT].reserve((size_t)ns * ldt * 8) || !B[CC].reserve((size_t)ns * ldt * 8) ||
            !B[NN].reserve((size_t)ns * 4) || !B[XX].reserve((size_t)ns * m * 8) ||
            !B[YY].reserve((size_t)ns * m * 8) || !B[IER].reserve((size_t)ns * 4))
            return FPB_ERR_ALLOC;
        bool ok = cuda_ok(cudaMemcpyAsync(B[TT].ptr, t + (size_t)c0 * ldt, (size_t)ns * ldt * 8,
                                          cudaMemcpyHostToDevice, st), "H2D t");
        ok = ok && cuda_ok(cudaMemcpyAsync(B[CC].ptr, c + (size_t)c0 * ldt, (size_t)ns * ldt * 8,
                                           cudaMemcpyHostToDevice, st), "H2D c");
        ok = ok && cuda_ok(cudaMemcpyAsync(B[NN].ptr, n + c0, (size_t)ns * 4, cudaMemcpyHostToDevice,
                                           st), "H2D n");
        ok = ok && cuda_ok(cudaMemcpyAsync(B[XX].ptr, x + (size_t)c0 * m, (size_t)ns * m * 8,
                                           cudaMemcpyHostToDevice, st), "H2D x");
        // splev leaves y untouched where it does not evaluate (e=2 early return): round-trip it
        if (e == FPB_OUTSIDE_NOT_ALLOWED)
            ok = ok && cuda_ok(cudaMemcpyAsync(B[YY].ptr, y + (size_t)c0 * m, (size_t)ns * m * 8,
                                               cudaMemcpyHostToDevice, st), "H2D y");
        if (!ok) return FPB_ERR_CUDA;
        int32_t rc = fp_splev_batch_dev_c(eng, (FP_SIZE)ns, B[TT].as<double>(), B[NN].as<int>(),
                                          B[CC].as<double>(), ldt, k, B[XX].as<double>(),
                                          B[YY].as<double>(), m, e, mode, 0, B[IER].as<int>(),
                                          nullptr);
        if (rc != FPB_STATUS_OK) return rc;
        ok = cuda_ok(cudaMemcpyAsync(y + (size_t)c0 * m, B[YY].ptr, (size_t)ns * m * 8,
                                     cudaMemcpyDeviceToHost, st), "D2H y");
        if (ier)
            ok = ok && cuda_ok(cudaMemcpyAsync(ier + c0, B[IER].ptr, (size_t)ns * 4,
                                               cudaMemcpyDeviceToHost, st), "D2H ier");
        if (!ok || !cuda_ok(cudaStreamSynchronize(st), "sync")) return FPB_ERR_CUDA;
    }
    return FPB_STATUS_OK;
}

template <class F> int32_t shard_over_gpus(long long total, int ngpus, long long align, F &&fn)
{
    int ndev = fpb_device_count();
    if (ndev == 0) {
        set_last_error("no CUDA device visible: fitpack_b200 has no CPU fallback");
        return FPB_ERR_CUDA;
    }
    if (ngpus <= 0 || ngpus > ndev) ngpus = ndev;
    if (total < (long long)ngpus * align) ngpus = (int)std::max<long long>(1, total / std::max<long long>(align, 1));
    if (ngpus <= 1) {
        HostApiLock lk(-1);
        fpb_engine *e = default_engine(-1);
        if (!e) return FPB_ERR_CUDA;
        return fn(e, 0LL, total);
    }
    std::vector<int32_t> rc(ngpus, FPB_STATUS_OK);
    std::vector<std::string> err(ngpus);
    std::vector<std::thread> th;
    long long per = (total + ngpus - 1) / ngpus;
    per = (per + align - 1) / align * align;
    g_slices_in_flight.fetch_add(ngpus);   // before any slice decides how much of the host it may claim
    for (int d = 0; d < ngpus; ++d) {
        const long long b0 = std::min(total, per * d), b1 = std::min(total, per * (d + 1));
        th.emplace_back([&, d, b0, b1]() {
            if (b0 >= b1) return;
            HostApiLock lk(d);
            fpb_engine *e = default_engine(d);
            if (!e) {
                rc[d] = FPB_ERR_CUDA;
                err[d] = fpb_last_error();
                return;
            }
            rc[d] = fn(e, b0, b1);
            if (rc[d] != FPB_STATUS_OK) err[d] = fpb_last_error();
        });
    }
    for (auto &t : th) t.join();
    g_slices_in_flight.fetch_sub(ngpus);
    for (int d = 0; d < ngpus; ++d)
        if (rc[d] != FPB_STATUS_OK) {
            set_last_error(err[d]);
            return rc[d];
        }
    return FPB_STATUS_OK;
}

}  // namespace

namespace fpb {
bool parcur_solo_host(fpb_engine *eng, long long nb, int32_t iopt, int32_t ipar, int32_t idim, int32_t m, double *u,
                      const double *x, const double *w, double *ub, double *ue, int32_t k, double s, int32_t nest,
                      int32_t *n, double *t, double *c, double *fp, int32_t *ier, double *fpint, int32_t *nrdata,
                      int32_t *status)
{
    if (!g_solo_on.load() || nb < 1 || nb > kSoloMaxBatch || !solo_supports(k, idim, 1) ||
        solo_smem_bytes(m, k, nest, idim) > kSoloSmemMax)
        return false;
    if (iopt == 1 && (!fpint || !nrdata)) return false;
    *status = FPB_STATUS_OK;
    auto fail = [&](int32_t rc) {
        *status = rc;
        for (long long b = 0; b < nb; ++b) ier[b] = rc;
        return true;
    };
    DeviceGuard g(eng->device);
    if (!g.ok) return fail(FPB_ERR_CUDA);
    const int k1 = k + 1, nmin = 2 * k1, nmax = m + k1;
    const bool chord = (ipar == 0 && iopt <= 0), want_state = fpint && nrdata;
    // ---- block, in doubles: u | x | w | ub,ue | fp | n | ier | t | c | fpint | nrdata ----
    const size_t pts = (size_t)nb * m, kn = (size_t)nb * nest, half = ((size_t)nb + 1) / 2;
    const size_t o_x = pts, o_w = o_x + pts * idim, o_be = o_w + pts, o_fp = o_be + 2 * (size_t)nb, o_n = o_fp + nb,
                 o_ier = o_n + half, o_t = o_ier + half, o_c = o_t + kn, o_f = o_c + kn * idim, o_nrd = o_f + kn,
                 o_end = o_nrd + (kn + 1) / 2;
    // pinned mirror from where the host needs it: everything for one curve; for more, u / x / w go up from the caller's
    // arrays (w = NULL: ones from the mirror)
    const size_t h0 = (nb == 1) ? 0 : (w ? o_be : o_w);
    if (!eng->scratch[0].reserve(o_end * 8) ||
        !reserve_pinned((void **)&eng->h_solo, &eng->h_solo_bytes, (o_end - h0) * 8))
        return fail(FPB_ERR_ALLOC);
    double *hd = eng->h_solo - h0, *dd = eng->scratch[0].as<double>();   // hd + offset is valid for offsets >= h0
    int *hn = reinterpret_cast<int *>(hd + o_n), *hier = reinterpret_cast<int *>(hd + o_ier),
        *hnrd = reinterpret_cast<int *>(hd + o_nrd);
    for (long long b = 0; b < nb; ++b) {
        const int n_in = (iopt != 0) ? n[b] : 0;
        const bool n_ok = n_in >= nmin && n_in <= nest;
        hd[o_be + 2 * b] = ub[b];
        hd[o_be + 2 * b + 1] = ue[b];
        hd[o_fp + b] = fp[b];
        hn[b] = n_in;
        hier[b] = 0;
        if (iopt != 0 && n_ok) std::memcpy(hd + o_t + (size_t)b * nest, t + (size_t)b * nest, (size_t)n_in * 8);
        if (iopt == 1 && n_ok) {
            std::memcpy(hd + o_f + (size_t)b * nest, fpint + (size_t)b * nest, (size_t)n_in * 8);
            std::memcpy(hnrd + (size_t)b * nest, nrdata + (size_t)b * nest, (size_t)n_in * 4);
        }
    }
    cudaStream_t st = eng->stream;
    bool ok = true;
    if (nb == 1) {
        if (!chord) std::memcpy(hd, u, pts * 8);
        std::memcpy(hd + o_x, x, pts * idim * 8);
        if (w) std::memcpy(hd + o_w, w, pts * 8);
        else std::fill(hd + o_w, hd + o_be, 1.0);
        ok = cuda_ok(cudaMemcpyAsync(dd, hd, (iopt == 1 ? o_end : o_c) * 8, cudaMemcpyHostToDevice, st), "H2D curve");
    } else {
        if (!chord) ok = cuda_ok(cudaMemcpyAsync(dd, u, pts * 8, cudaMemcpyHostToDevice, st), "H2D u");
        ok = ok && cuda_ok(cudaMemcpyAsync(dd + o_x, x, pts * idim * 8, cudaMemcpyHostToDevice, st), "H2D x");
        if (w) {
            ok = ok && cuda_ok(cudaMemcpyAsync(dd + o_w, w, pts * 8, cudaMemcpyHostToDevice, st), "H2D w");
        } else {
            std::fill(hd + o_w, hd + o_be, 1.0);
            ok = ok && cuda_ok(cudaMemcpyAsync(dd + o_w, hd + o_w, pts * 8, cudaMemcpyHostToDevice, st), "H2D w");
        }
        ok = ok && cuda_ok(cudaMemcpyAsync(dd + o_be, hd + o_be, ((iopt != 0 ? o_c : o_t) - o_be) * 8, cudaMemcpyHostToDevice, st),
                           "H2D state");
        if (iopt == 1)
            ok = ok && cuda_ok(cudaMemcpyAsync(dd + o_f, hd + o_f, (o_end - o_f) * 8, cudaMemcpyHostToDevice, st), "H2D state");
    }
    SoloParams p;
    std::memset(&p, 0, sizeof p);
    p.nbatch = (int)nb; p.iopt = iopt; p.m = m; p.k = k; p.nest = nest;
    p.idim = idim; p.par = 1; p.chord = chord ? 1 : 0; p.s = s;
    p.x = dd; p.u_out = dd; p.y = dd + o_x; p.w = dd + o_w; p.ube = dd + o_be; p.fp = dd + o_fp;
    p.n = reinterpret_cast<int *>(dd + o_n); p.ier = reinterpret_cast<int *>(dd + o_ier);
    p.t = dd + o_t; p.c = dd + o_c; p.c_stride = (long long)nest * idim;
    p.fpint = dd + o_f; p.nrdata = reinterpret_cast<int *>(dd + o_nrd);
    ok = ok && cuda_ok(launch_curfit_solo(p, st), "curfit_solo_kernel (parcur)");
    if (chord) ok = ok && cuda_ok(cudaMemcpyAsync(u, dd, pts * 8, cudaMemcpyDeviceToHost, st), "D2H u");
    ok = ok && cuda_ok(cudaMemcpyAsync(hd + o_be, dd + o_be, ((want_state ? o_end : o_f) - o_be) * 8, cudaMemcpyDeviceToHost, st),
                       "D2H fit") &&
         cuda_ok(cudaStreamSynchronize(st), "sync");
    eng->launches++;
    if (!ok) return fail(FPB_ERR_CUDA);
    const bool untouched_all = (iopt >= 0 && s <= 0.0 && nmax > nest);   // ier = 1 before anything was computed
    for (long long b = 0; b < nb; ++b) {
        const int n_out = hn[b], ier_out = hier[b];
        ier[b] = ier_out;
        if (chord) {   // the parameters are computed before the checks (core:15238-15253): visible either way
            ub[b] = hd[o_be + 2 * b];
            ue[b] = hd[o_be + 2 * b + 1];
        }
        if (ier_out == FPB_INPUT_ERROR) {   // the clamped end knots of a rejected iopt = -1 curve (core:15264-15270)
            const int n_in = (iopt != 0) ? n[b] : 0;
            if (iopt < 0 && n_in >= nmin && n_in <= nest && n_in <= nmax)
                std::memcpy(t + (size_t)b * nest, hd + o_t + (size_t)b * nest, (size_t)n_in * 8);
            continue;
        }
        n[b] = n_out;
        if (untouched_all || n_out < nmin || n_out > nest) continue;
        fp[b] = hd[o_fp + b];
        std::memcpy(t + (size_t)b * nest, hd + o_t + (size_t)b * nest, (size_t)n_out * 8);
        for (int d = 0; d < idim; ++d)
            std::memcpy(c + (size_t)b * nest * idim + (size_t)d * n_out, hd + o_c + (size_t)b * nest * idim + (size_t)d * n_out,
                        (size_t)(n_out - k1) * 8);
        if (want_state) {
            std::memcpy(fpint + (size_t)b * nest, hd + o_f + (size_t)b * nest, (size_t)n_out * 8);
            std::memcpy(nrdata + (size_t)b * nest, hnrd + (size_t)b * nest, (size_t)n_out * 4);
        }
    }
    return true;
}

int32_t shard_units_over_gpus(long long total, int ngpus, long long align,
                              const std::function<int32_t(fpb_engine *, long long, long long)> &fn)
{
    return shard_over_gpus(total, ngpus, align, fn);
}
}  // namespace fpb

extern "C" {

int32_t fpb_set_single_curve_kernel(int32_t on) { return g_solo_on.exchange(on ? 1 : 0); }

int32_t fpb_host_sharers(void) { return host_sharers(); }

const char *fpb_host_form(int64_t nbatch)
{
    if (nbatch <= kSoloMaxBatch && g_solo_on.load()) return "warp-per-curve";
    return wavefront_preferred(nbatch) ? "wavefront" : "chunks";
}


int32_t fp_curfit_batch_c(FP_SIZE nbatch, FP_SIZE iopt, FP_SIZE m, const FP_REAL *x,
                          const FP_REAL *y, const FP_REAL *w, FP_SIZE shared_x, const FP_REAL *xb,
                          const FP_REAL *xe, FP_SIZE k, FP_REAL s, FP_SIZE nest, FP_SIZE *n,
                          FP_REAL *t, FP_REAL *c, FP_REAL *fp, FP_FLAG *ier, FP_SIZE ngpus)
{
    if (nbatch <= 0) return FPB_STATUS_OK;
    if (!x || !y || !n || !t || !c || !fp || !ier || (xb && !xe) || (!xb && xe)) {
        set_last_error("fp_curfit_batch_c: null required pointer");
        return FPB_ERR_ARG;
    }
    if (iopt == 1) {
        set_last_error("fp_curfit_batch_c: iopt=1 needs per-curve state; use fp_curfit_c or the "
                       "device entry point with fpint/nrdata");
        return FPB_ERR_ARG;
    }
    if (m < 1 || nest < 1) {  // nothing addressable: mirror curfit's ier=10
        for (FP_SIZE b = 0; b < nbatch; ++b) ier[b] = FPB_INPUT_ERROR;
        return FPB_STATUS_OK;
    }
    return shard_over_gpus(nbatch, ngpus, 128, [&](fpb_engine *e, long long b0, long long b1) {
        return fit_slice_host(e->device, b0, b1, iopt, m, x, y, w, shared_x, xb, xe, k, s, nest, n,
                              t, c, fp, ier, nullptr, nullptr);
    });
}

int32_t fp_splev_batch_c(FP_SIZE nspl, const FP_REAL *t, const FP_SIZE *n, const FP_REAL *c,
                         FP_SIZE ldt, FP_SIZE k, const FP_REAL *x, FP_REAL *y, FP_SIZE m, FP_FLAG e,
                         FP_SIZE mode, FP_FLAG *ier, FP_SIZE ngpus)
{
    if (nspl <= 0) return FPB_STATUS_OK;
    if (!t || !n || !c || !x || !y) {
        set_last_error("fp_splev_batch_c: null required pointer");
        return FPB_ERR_ARG;
    }
    if (m < 1) {
        if (ier)
            for (FP_SIZE j = 0; j < nspl; ++j) ier[j] = FPB_INPUT_ERROR;
        return FPB_STATUS_OK;
    }
    return shard_over_gpus(nspl, ngpus, 8, [&](fpb_engine *eng, long long j0, long long j1) {
        return splev_slice_host(eng, j0, j1, t, n, c, ldt, k, x, y, m, e, mode, ier);
    });
}

// ---- the reference's own flat entry points: a batch of one curve ----

void fp_curfit_c(FP_SIZE iopt, FP_SIZE m, const FP_REAL *x, const FP_REAL *y, const FP_REAL *w,
                 FP_REAL xb, FP_REAL xe, FP_SIZE k, FP_REAL s, FP_SIZE nest, FP_SIZE *n, FP_REAL *t,
                 FP_REAL *c, FP_REAL *fp, FP_REAL *wrk, FP_SIZE lwrk, FP_SIZE *iwrk, FP_FLAG *ier)
{
    // data check that does not need the data, core:1259-1270
    *ier = FPB_INPUT_ERROR;
    if (k <= 0 || k > 5) return;
    if (iopt < -1 || iopt > 1) return;
    if (m < k + 1 || nest < 2 * (k + 1)) return;
    if (lwrk < m * (k + 1) + nest * (7 + 3 * k)) return;
    fpb_engine *e = default_engine(-1);
    if (!e) {
        *ier = FPB_ERR_CUDA;  // no device: fail loudly, there is no CPU path
        return;
    }
    // wrk(1:nest) is fpint and iwrk(1:nest) is nrdata in the reference (core:1289-1297):
    // that is the state an iopt=1 continuation reads back.
    int32_t rc = fit_slice_host(e->device, 0, 1, iopt, m, x, y, w, 0, &xb, &xe, k, s, nest, n, t, c,
                                fp, ier, wrk, iwrk);
    if (rc != FPB_STATUS_OK) *ier = rc;
}

void fp_splev_c(const FP_REAL *t, FP_SIZE n, const FP_REAL *c, FP_SIZE k, const FP_REAL *x,
                FP_REAL *y, FP_SIZE m, FP_FLAG e, FP_FLAG *ier)
{
    *ier = FPB_INPUT_ERROR;
    if (m < 1) return;  // core:17844-17845
    HostApiLock lk(-1);
    fpb_engine *eng = default_engine(-1);
    if (!eng) {
        *ier = FPB_ERR_CUDA;
        return;
    }
    // one spline, its m points spread over the whole device (nshared mode)
    DeviceGuard g(eng->device);
    cudaStream_t st = eng->stream;
    DevBuf *B = eng->scratch;
    // up to 4096 points: one chunk of exactly m points (no padding, no extra copy or sync)
    const int chunk = m <= 4096 ? m : 4096;
    const long long nchunks = ((long long)m + chunk - 1) / chunk;
    const size_t padded = (size_t)nchunks * chunk;
    const int ldt = std::max((n + 1) & ~1, 2 * (k + 1));
    if (k < 1 || k > 5 || n < 2 * (k + 1)) return;  // not a valid spline: nothing safe to evaluate
    if (!B[0].reserve((size_t)ldt * 8) || !B[1].reserve((size_t)ldt * 8) || !B[2].reserve(64) ||
        !B[3].reserve(padded * 8) || !B[4].reserve(padded * 8) ||
        !B[5].reserve((size_t)nchunks * 4)) {
        *ier = FPB_ERR_ALLOC;
        return;
    }
    bool ok = cuda_ok(cudaMemcpyAsync(B[0].ptr, t, (size_t)n * 8, cudaMemcpyHostToDevice, st), "H2D");
    ok = ok && cuda_ok(cudaMemcpyAsync(B[1].ptr, c, (size_t)n * 8, cudaMemcpyHostToDevice, st), "H2D");
    ok = ok && cuda_ok(cudaMemcpyAsync(B[2].ptr, &n, 4, cudaMemcpyHostToDevice, st), "H2D");
    ok = ok && cuda_ok(cudaMemcpyAsync(B[3].ptr, x, (size_t)m * 8, cudaMemcpyHostToDevice, st), "H2D");
    if (e == FPB_OUTSIDE_NOT_ALLOWED)
        ok = ok && cuda_ok(cudaMemcpyAsync(B[4].ptr, y, (size_t)m * 8, cudaMemcpyHostToDevice, st),
                           "H2D");
    // pad the tail chunk with an in-support abscissa so that it cannot raise ier=6
    if (ok && padded > (size_t)m) {
        std::vector<double> pad(padded - m, x[0]);
        // x[0] itself may be outside the support; t(k+1) never is
        std::fill(pad.begin(), pad.end(), t[k]);
        ok = cuda_ok(cudaMemcpyAsync(B[3].as<double>() + m, pad.data(), pad.size() * 8,
                                     cudaMemcpyHostToDevice, st), "H2D pad");
        ok = ok && cuda_ok(cudaStreamSynchronize(st), "sync");
    }
    if (!ok) {
        *ier = FPB_ERR_CUDA;
        return;
    }
    int32_t rc = fp_splev_batch_dev_c(eng, (FP_SIZE)nchunks, B[0].as<double>(), B[2].as<int>(),
                                      B[1].as<double>(), ldt, k, B[3].as<double>(),
                                      B[4].as<double>(), chunk, e, FPB_SPLEV_EXACT, 1,
                                      B[5].as<int>(), nullptr);
    if (rc != FPB_STATUS_OK) {
        *ier = rc;
        return;
    }
    std::vector<int> hier((size_t)nchunks);
    ok = cuda_ok(cudaMemcpyAsync(hier.data(), B[5].ptr, (size_t)nchunks * 4, cudaMemcpyDeviceToHost,
                                 st), "D2H");
    ok = ok && cuda_ok(cudaStreamSynchronize(st), "sync");
    if (!ok) {
        *ier = FPB_ERR_CUDA;
        return;
    }
    int result = FPB_OK;
    for (int v : hier)
        if (v != FPB_OK) result = v;
    if (result == FPB_INVALID_RANGE) {
        // reference semantics (core:17871-17873): y is filled up to the first offender only
        const double tb = t[k], te = t[n - k - 1];
        long long first = m;
        for (long long i = 0; i < m; ++i)
            if (x[i] < tb || x[i] > te) {
                first = i;
                break;
            }
        if (first > 0)
            ok = cuda_ok(cudaMemcpy(y, B[4].ptr, (size_t)first * 8, cudaMemcpyDeviceToHost), "D2H y");
    } else {
        ok = cuda_ok(cudaMemcpy(y, B[4].ptr, (size_t)m * 8, cudaMemcpyDeviceToHost), "D2H y");
    }
    *ier = ok ? result : FPB_ERR_CUDA;
}

}  // extern "C"
