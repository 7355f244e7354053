#!/usr/bin/env python
"""bench.py -- headline benchmark of the hot path (BASELINE.json metric):
curve fits/sec (256-pt, k=3) and splev GB/s, on N B200s vs the host CPU.

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K ...  # reference algorithm on host cores

A "step" is one pass of the batched fit over one batch of synthetic curves
(BASELINE config 2: 2^20 curves x 256 points per GPU, k=3, s=m*sigma^2=2.56, nest=359,
point-major SoA resident in HBM).  N>1: one process per GPU (torchrun), every rank fits its
own batch -- independent curves, no collective in the data path ("scaling": "weak").
`value`   : whole-job fits/s with inputs resident in HBM (CUDA events, max over ranks).
`e2e`     : the same metric through the C-ABI host-pointer call fp_curfit_batch_c: pinned host
            buffers, H2D + tile transpose + fit + D2H all inside the timed region.
`splev`   : second half of the metric (BASELINE config 5 shape: 2^20 k=5 splines x 954 points per
            GPU): GB/s at 16.5 algorithmic bytes/point, exact and fast modes, own roofline.
`regrid`  : BASELINE config 4 (one 4096 x 4096 gridded smoothing fit, z resident in HBM) and
            config 5b (bispev on that grid), with the serial CPU port timed on a 1024^2 sample.
`roofline`: fit kernel against measured HBM peak (it is FP64 issue/latency bound, so this
            fraction is tiny by nature) -- the meaningful bound is in `fp64`.
The reference arm times the CPU restatement of the reference algorithm (oracle/, OpenMP over
curves, all host threads): the reference itself is Fortran and cannot be built in this image.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

M, K, SIGMA, NEST = 256, 3, 0.1, 359
S = M * SIGMA * SIGMA  # 2.56
B_PER_GPU = 1 << 20
SPLEV_NSPL, SPLEV_M, SPLEV_K, SPLEV_N = 1 << 20, 954, 5, 32
SHARED_B, SHARED_M, SHARED_K, SHARED_NINT = 1 << 24, 512, 3, 28   # config 3 AT ITS STATED SIZE: 2^24 curves x 512 points on one GPU (68.7 GB of y)
GRID_M, GRID_K, GRID_SIGMA, GRID_NEST = 4096, 3, 0.05, 128          # config 4: 4096 x 4096 grid, kx=ky=3
PARCUR_B, PARCUR_M, PARCUR_IDIM = 1 << 19, 256, 2                   # parametric curves
PERCUR_B, PERCUR_M = 1 << 20, 256                                   # periodic curves
# DRAM traffic per unit measured with `ncu --set full` (dram__bytes_read.sum + dram__bytes_write.sum); the bench
# itself never runs under ncu.  The fit pipeline's figure is read from profiles/fit_traffic.json, which is
# rewritten whenever a new capture of the pass / part-2 kernels is committed (it names its source round).
def fit_traffic_record():
    try:
        with open(os.path.join(ROOT, "profiles", "fit_traffic.json")) as f:
            d = json.load(f)
        return float(d["dram_bytes_per_fit"]), str(d["source"])
    except Exception:
        return None, "no committed ncu capture (profiles/fit_traffic.json missing)"


NCU_SPLEV_DRAM_BYTES_PER_POINT = (8.561e9 + 7.965e9) / (1048576 * 954)   # profiles/r01b_kernels.md
NCU_SHARED_DRAM_BYTES_PER_CURVE = (17.85e9 + 1.839e9) / 4194304   # shared_apply2 (r01b)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--curves", type=int, default=B_PER_GPU, help="curves per GPU per step")
    ap.add_argument("--shared-curves", type=int, default=SHARED_B, help="config-3 curves per GPU")
    ap.add_argument("--no-splev", action="store_true")
    ap.add_argument("--no-shared", action="store_true")
    ap.add_argument("--no-regrid", action="store_true")
    ap.add_argument("--no-parcur", action="store_true")
    ap.add_argument("--no-percur", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    return ap.parse_args()


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region."""
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.samples = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.samples:
            parts = [q.strip() for q in ln.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def host_threads():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


SEED = 1234
PARITY_SAMPLE = 8192     # curves of the SAME batch re-fitted by the oracle (n, fp, ier bit-identical)


def synth_device(torch, dev, b0, B, m):
    """Config-2 batch generated in HBM, point-major [m][B]: curves b0 .. b0+B-1 of the counter-based
    generator tests/synthdata.py -- the CPU arm regenerates the same curves bit for bit."""
    import synthdata
    x = torch.empty((m, B), device=dev, dtype=torch.float64)
    y = torch.empty((m, B), device=dev, dtype=torch.float64)
    step = 1 << 18
    for c0 in range(0, B, step):
        nb = min(step, B - c0)
        xc, yc = synthdata.c2_curves(torch, b0 + c0, nb, m, SIGMA, SEED, device=dev)
        x[:, c0:c0 + nb] = xc
        y[:, c0:c0 + nb] = yc
    return x, y


def synth_host(np, b0, B, m):
    import synthdata
    return synthdata.c2_curves(np, b0, B, m, SIGMA, SEED)


def fit_flops(n, passes, piters, m=M, k=K):
    """algorithmic flop count of SURVEY.md 8(d) (div and sqrt = 1 flop), from the kernel's own
    per-curve counters P (LSQ passes) and I (p-iterations) and the final knot count n."""
    k1 = k + 1
    nk1 = n - k1
    n8 = n - 2 * k1
    per_pass = m * (3.5 * k * (k + 1) + k1 + 1 + 13 * k1 + 3 * k1 * (k1 - 1) + 2) \
        + nk1 * (2 * k + 1) + m * (2 * k1 + 3)
    per_it = (13 + 6 * k1) * (n8 * nk1 - n8 * (n8 - 1) / 2) + nk1 * (2 * k1 + 1) + m * (2 * k1 + 3)
    return passes * per_pass + piters * per_it


def fit_divsqrt(n, passes, piters, m=M, k=K):
    """IEEE divisions and square roots per fit (same counters): fpbspl k(k+1)/2 divisions per point,
    3 divisions + 1 square root per Givens rotation (fpgivs), one division per back-substituted row."""
    k1 = k + 1
    nk1 = n - k1
    n8 = n - 2 * k1
    steps = n8 * nk1 - n8 * (n8 - 1) / 2          # rotations of one p-iteration
    ndiv = passes * (m * (k * (k + 1) / 2 + 3 * k1) + nk1) + piters * (3 * steps + nk1 + 1)
    nsqrt = passes * (m * k1) + piters * steps
    return ndiv, nsqrt


def run_reference(args, rank, world):
    """CPU arm: the reference algorithm (oracle port; the Fortran cannot be built here) with all
    host threads, each step a bounded sample of the same workload."""
    if rank != 0:
        return
    import numpy as np
    o3 = os.path.join(ROOT, "oracle", "libfitpack_oracle_O3.so")
    build_kind = "-O2 -ffp-contract=off"
    if os.path.exists(o3):   # the CPU arm is timed at -O3 (same source, still no FMA)
        os.environ["FITPACK_ORACLE_LIB"] = o3
        build_kind = "-O3 -ffp-contract=off"
    import oracle_lib as O
    # torchrun exports OMP_NUM_THREADS=1: ask for every core this process may run on explicitly
    threads = host_threads()
    x, y = synth_host(np, 0, 65536, M)   # the first 65536 curves of rank 0's GPU batch, bit for bit
    t0 = time.perf_counter()
    O.curfit_batch(0, x[:256], y[:256], None, K, S, NEST, nthreads=threads)
    rate = 256 / (time.perf_counter() - t0)
    budget = 120.0 / max(1, args.steps + args.warmup)       # seconds per step
    sample = int(min(65536, max(64, rate * min(budget, 8.0))))
    xs, ys = x[:sample], y[:sample]
    for _ in range(args.warmup):
        O.curfit_batch(0, xs, ys, None, K, S, NEST, nthreads=threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        out = O.curfit_batch(0, xs, ys, None, K, S, NEST, nthreads=threads)
    dt = time.perf_counter() - t0
    value = sample * args.steps / dt
    assert (out["ier"] <= 0).all()
    desc = ("first %d curves of the GPU arm's rank-0 batch (counter-based generator, bit-identical inputs) "
            "per step; oracle port built %s (no FMA: the reference's arithmetic)" % (sample, build_kind))
    print(json.dumps({
        "impl": "reference", "metric": "curve fits/sec (256-pt, k=3)", "value": value,
        "unit": "fits/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "batched smoothing fit: %d-pt curves, k=3, s=2.56, nest=359, "
                               "per-curve knot insertion (config 2)" % M,
                   "curves_per_step": sample},
        "cpu_baseline": {"value": value, "unit": "fits/s", "cores": threads, "kind": "port",
                         "sample": desc},
        "e2e": {"value": value, "unit": "fits/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import numpy as np
    import torch
    import torch.distributed as dist

    import fitpack_b200
    from fitpack_b200.batch import Engine

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU: fitpack_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    eng = Engine(dev)
    B = args.curves
    hbm_peak, peak_src = measured_peaks()
    x, y = synth_device(torch, dev, rank * B, B, M)
    out = {}
    # ---------------------------------------------------------------- fit: device-resident
    for _ in range(args.warmup):
        res = eng.curfit_batch(x, y, None, k=K, s=S, nest=NEST, want_stats=True, out=out)
        out = {kk: res[kk] for kk in ("n", "t", "c", "fp", "ier", "stats")}
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = eng.launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kernel_ms = []
    barrier()
    ev0.record()
    for _ in range(args.steps):
        res = eng.curfit_batch(x, y, None, k=K, s=S, nest=NEST, want_stats=True, out=out)
    ev1.record()
    barrier()
    clocks = sampler.stop()
    step_ms = ev0.elapsed_time(ev1) / args.steps
    kernel_ms.append(eng.last_kernel_ms())  # CUDA events around the last fit-kernel launch
    launches = eng.launches - launches0
    step_ms_max = max_over_ranks(step_ms)
    value = world * B / (step_ms_max * 1e-3)

    n = res["n"]
    ier = res["ier"]
    stats = res["stats"].to(torch.float64)
    n_f = n.to(torch.float64)
    flops = float(fit_flops(n_f, stats[:, 0], stats[:, 1]).sum().item())
    mean_n = float(n_f.mean().item())
    ok_frac = float((ier <= 0).to(torch.float64).mean().item())
    gathered = None
    if world > 1:
        # optional final gather of the per-curve summary over NVLink (outside the timed region);
        # the fit itself has no collective
        from fitpack_b200.dist import gather_fit_summary
        n_all, fp_all, ier_all = gather_fit_summary(n, res["fp"], ier)
        ok_frac = float((ier_all <= 0).to(torch.float64).mean().item())
        gathered = int(n_all.numel())
    alg_bytes = B * (2 * 8 * M + 16 * mean_n + 16)
    fit_ms = kernel_ms[-1]
    traffic_per_fit, traffic_src = fit_traffic_record()
    roof = {"bound": "hbm", "kernel": "fit pipeline: curfit_pass_kernel<3> x R + curfit_part2_kernel<3>",
            "achieved": alg_bytes / (fit_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
            "traffic": None if traffic_per_fit is None else traffic_per_fit * B,
            "traffic_source": traffic_src,
            "peak_source": peak_src,
            "algorithmic_bytes_per_fit": alg_bytes / B,
            "note": "fit kernel is FP64 issue/latency bound (~%.0f kflop per 4.4 KB): HBM "
                    "fraction is tiny by construction; see fp64" % (flops / B / 1e3)}
    roof["frac"] = roof["achieved"] / roof["peak"]
    # the same kernels measured by the bytes they actually move (ncu): workspace traffic next to the
    # FP64-pipe share (see fp64)
    if traffic_per_fit is not None:
        roof["traffic_gbs"] = roof["traffic"] / (fit_ms * 1e-3) / 1e9
        roof["traffic_frac_of_peak"] = roof["traffic_gbs"] / hbm_peak
    dmul_dadd = eng.probe_fp64_gops(False)
    dfma = eng.probe_fp64_gops(True)
    ndiv_t, nsqrt_t = fit_divsqrt(n_f, stats[:, 0], stats[:, 1])
    ndiv, nsqrt = float(ndiv_t.sum().item()), float(nsqrt_t.sum().item())
    ddiv = eng.probe_fp64_gops(2)
    dsqrt = eng.probe_fp64_gops(3)
    # FP64 issue slots: one per mul/add, (mul rate / div rate) per IEEE division, likewise sqrt
    slots = None
    if dmul_dadd > 0 and ddiv > 0 and dsqrt > 0:
        slots = (flops - ndiv - nsqrt) + ndiv * (dmul_dadd / ddiv) + nsqrt * (dmul_dadd / dsqrt)
    fp64 = {"achieved_gflops": flops / (fit_ms * 1e-3) / 1e9,
            "peak_gops_dmul_dadd": dmul_dadd, "peak_gops_dfma": dfma,
            "frac_of_unfused_peak": flops / (fit_ms * 1e-3) / 1e9 / dmul_dadd if dmul_dadd > 0 else None,
            "flops_per_fit": flops / B, "mean_lsq_passes": float(stats[:, 0].mean().item()),
            "mean_p_iters": float(stats[:, 1].mean().item()), "mean_n": mean_n,
            "max_n": int(n.max().item()),
            # distribution of the per-curve work: percentiles 50 / 90 / 99 / 100
            "n_percentiles": [int(v) for v in torch.quantile(n_f[:1 << 20], torch.tensor([.5, .9, .99, 1.], device=dev, dtype=torch.float64)).tolist()],
            "lsq_passes_percentiles": [int(v) for v in torch.quantile(stats[:1 << 20, 0], torch.tensor([.5, .9, .99, 1.], device=dev, dtype=torch.float64)).tolist()],
            "p_iters_percentiles": [int(v) for v in torch.quantile(stats[:1 << 20, 1], torch.tensor([.5, .9, .99, 1.], device=dev, dtype=torch.float64)).tolist()],
            "peak_gops_ddiv": ddiv, "peak_gops_dsqrt": dsqrt,
            "div_per_fit": ndiv / B, "sqrt_per_fit": nsqrt / B,
            "issue_slots_per_fit": None if slots is None else slots / B,
            "fits_per_s_at_fp64_issue_roofline": None if slots is None else dmul_dadd * 1e9 / (slots / B),
            "frac_of_fp64_issue_roofline": None if slots is None else (B / (fit_ms * 1e-3)) / (dmul_dadd * 1e9 / (slots / B)),
            "note": "the bound of this pipeline: measured rates of unfused DMUL/DADD, IEEE division and square "
                    "root on this GPU; a correctly rounded division costs ~13.6 FP64 issue slots"}

    roof["fp64_bound"] = {
        "probe_gops": {"dmul_dadd": dmul_dadd, "dfma": dfma, "ddiv": ddiv, "dsqrt": dsqrt},
        "probe_source": "fpb_probe_fp64_gops on this GPU, this run (MEASURED_PEAKS.json has no FP64 figure)",
        "achieved_gflops": fp64["achieved_gflops"], "frac_of_unfused_peak": fp64["frac_of_unfused_peak"],
        "issue_slots_per_fit": fp64["issue_slots_per_fit"],
        "frac_of_fp64_issue_roofline": fp64["frac_of_fp64_issue_roofline"],
        "mean_lsq_passes": fp64["mean_lsq_passes"], "mean_p_iters": fp64["mean_p_iters"], "mean_n": mean_n}
    # dependent-issue latency bound: every warp of the fit kernels is one chain of dependent FP64 instructions
    # (ILP ~ 1), so a scheduler with W resident warps issues at most W / latency instructions per cycle
    lat_ma = eng.probe_fp64_gops(4)
    lat_fma = eng.probe_fp64_gops(5)
    lat_mufu = eng.probe_fp64_gops(6)
    roof["fp64_bound"]["dependent_latency_cycles"] = {"dmul_dadd": lat_ma, "dfma": lat_fma, "mufu_rcp64h_plus_dfma_pair_avg": lat_mufu}
    if lat_fma and lat_fma > 0:
        roof["fp64_bound"]["issue_rate_bound_per_scheduler"] = {
            "warps_per_scheduler": 5, "ipc_bound": min(1.0, 5.0 / lat_fma),
            "note": "20 warps per SM (96 registers) = 5 per scheduler; ncu: issue slots busy 56 % in the pass "
                    "kernel, 49 % in part 2 (profiles/r02_kernels.md)"}
    roof["secondary"] = []   # the other kernels of the metric, filled in below (driver keeps `roofline`)

    # ---------------------------------------------------------------- one curve per call (the reference's call pattern)
    if rank == 0 and not args.no_e2e:
      try:
        L1 = fitpack_b200.lib()
        xs1, ys1 = synth_host(np, 0, 1, M)
        x1, y1 = np.ascontiguousarray(xs1[0]), np.ascontiguousarray(ys1[0])

        def one_call():
            return fitpack_b200.curfit(0, x1, y1, None, float(x1[0]), float(x1[-1]), K, S, NEST)

        def lat(fn, reps):
            fn()
            fn()
            t0 = time.perf_counter()
            for _ in range(reps):
                fn()
            return 1e3 * (time.perf_counter() - t0) / reps
        solo_ms = lat(one_call, 100)
        prev = L1.fpb_set_single_curve_kernel(0)
        batch1_ms = lat(one_call, 20)
        L1.fpb_set_single_curve_kernel(prev)
        roof["single_call"] = {
            "api": "fp_curfit_c, curve 0 of the bench batch (m = %d, k = %d, s = %g), wall clock per call" % (M, K, S),
            "warp_per_curve_kernel_ms": solo_ms, "batch_kernels_one_thread_per_curve_ms": batch1_ms,
            "kernel": "curfit_solo_kernel<3,1,false>: one warp per curve, band-QR sweeps as a wavefront over lane groups; "
                      "bound by the dependent Givens chain (29 FP64 + 3 MUFU per tick), not by HBM"}
        if not args.no_cpu:
            import oracle_lib as O1
            roof["single_call"]["cpu_port_ms"] = lat(
                lambda: O1.curfit(0, x1, y1, None, float(x1[0]), float(x1[-1]), K, S, NEST), 50)
      except Exception as exc:   # a secondary record must not cost the headline line
        roof["single_call"] = {"error": repr(exc)}

    # ---------------------------------------------------------------- parity sample on the SAME batch
    parity = None
    if rank == 0 and not args.no_cpu:
        import oracle_lib as O
        ns_ = min(PARITY_SAMPLE, B)
        xs_h, ys_h = synth_host(np, 0, ns_, M)
        same_in = bool(np.array_equal(xs_h, x[:, :ns_].t().cpu().numpy())
                       and np.array_equal(ys_h, y[:, :ns_].t().cpu().numpy()))
        op = O.curfit_batch(0, xs_h, ys_h, None, K, S, NEST, nthreads=host_threads())
        n_h, fp_h, ier_h = (res[kk][:ns_].cpu().numpy() for kk in ("n", "fp", "ier"))
        t_h, c_h = res["t"][:ns_].cpu().numpy(), res["c"][:ns_].cpu().numpy()
        same_tc = all(np.array_equal(t_h[b, :op["n"][b]], op["t"][b, :op["n"][b]])
                      and np.array_equal(c_h[b, :op["n"][b] - K - 1], op["c"][b, :op["n"][b] - K - 1])
                      for b in range(ns_))
        parity = {"curves": ns_, "inputs_bit_identical_host_vs_device": same_in,
                  "n_ier_identical": bool(np.array_equal(n_h, op["n"]) and np.array_equal(ier_h, op["ier"])),
                  "fp_bit_identical": bool(np.array_equal(fp_h, op["fp"])),
                  "t_c_bit_identical": bool(same_tc),
                  "what": "first %d curves of this batch re-fitted by the CPU oracle (reference algorithm)" % ns_}

    # ---------------------------------------------------------------- fit: end to end (host buffers)
    e2e = None
    if not args.no_e2e:
        L = fitpack_b200.lib()
        Be = B
        xh = torch.empty((Be, M), dtype=torch.float64, pin_memory=True)
        yh = torch.empty((Be, M), dtype=torch.float64, pin_memory=True)
        xh.copy_(eng.transpose(x))   # curve-major, as a caller of curfit holds the data
        yh.copy_(eng.transpose(y))
        nh = torch.zeros(Be, dtype=torch.int32, pin_memory=True)
        th = torch.zeros((Be, NEST), dtype=torch.float64, pin_memory=True)
        ch = torch.zeros((Be, NEST), dtype=torch.float64, pin_memory=True)
        fph = torch.zeros(Be, dtype=torch.float64, pin_memory=True)
        ierh = torch.zeros(Be, dtype=torch.int32, pin_memory=True)
        torch.cuda.synchronize()

        def e2e_step():
            rc = L.fp_curfit_batch_c(Be, 0, M, xh.data_ptr(), yh.data_ptr(), None, 0, None, None,
                                     K, S, NEST, nh.data_ptr(), th.data_ptr(), ch.data_ptr(),
                                     fph.data_ptr(), ierh.data_ptr(), 1)
            if rc != 0:
                raise RuntimeError("fp_curfit_batch_c failed: " + L.fpb_last_error().decode())

        for _ in range(3):   # warm-up: staging buffers of both host forms (the library measures the copy rate on the
            e2e_step()       # first call and may switch from the wavefront to the chunk-per-engine form on the next)
        esteps = max(1, args.steps)     # timed over the same number of steps as `value`
        barrier()
        step_ms = []
        t0 = time.perf_counter()
        for _ in range(esteps):
            ts = time.perf_counter()
            e2e_step()           # returns with the results in the caller's host arrays
            step_ms.append(1e3 * (time.perf_counter() - ts))
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        dt = max_over_ranks(dt)
        nsum = int(nh.to(torch.int64).sum().item())
        assert bool((nh == n.cpu()).all()) and bool((ierh == ier.cpu()).all()), "e2e != device path"
        # spot check of the scattered knots/coefficients against the device-resident result
        tdev, cdev = res["t"][:4096].cpu(), res["c"][:4096].cpu()
        for b in range(0, 4096, 97):
            nb_ = int(nh[b])
            assert bool((th[b, :nb_] == tdev[b, :nb_]).all()) and bool((ch[b, :nb_ - K - 1] == cdev[b, :nb_ - K - 1]).all()), "e2e t/c"
        e2e = {"value": world * Be * esteps / dt, "unit": "fits/s",
               "h2d_bytes_per_step": int(Be * M * 8 * 2 + Be * 8 + 24),
               # n, fp, ier, dense offsets per curve + the dense t(1:n), c(1:n) of every curve
               "d2h_bytes_per_step": int(Be * 24 + 2 * nsum * 8),
               "steps": esteps, "ms_per_step": 1e3 * dt / esteps,
               "ms_per_step_min_median_max_rank0": [round(min(step_ms), 2), round(sorted(step_ms)[len(step_ms) // 2], 2),
                                                    round(max(step_ms), 2)],
               "api": "fp_curfit_batch_c (host pointers, pinned)"}
        del xh, yh, th, ch

    # ---------------------------------------------------------------- splev (config 5 shape)
    splev = None
    if not args.no_splev:
        del x, y, out, res
        torch.cuda.empty_cache()
        g = torch.Generator(device=dev)
        g.manual_seed(99 + rank)
        f64 = torch.float64
        ns, ms, ks, nn = SPLEV_NSPL, SPLEV_M, SPLEV_K, SPLEV_N
        t = torch.zeros((ns, nn), device=dev, dtype=f64)
        nint = nn - 2 * (ks + 1)
        t[:, ks + 1:ks + 1 + nint] = torch.linspace(0, 1, nint + 2, device=dev, dtype=f64)[1:-1]
        t[:, ks + 1 + nint:] = 1.0
        c = torch.randn((ns, nn), generator=g, device=dev, dtype=f64)
        nvec = torch.full((ns,), nn, device=dev, dtype=torch.int32)
        xe = torch.rand((ns, ms), generator=g, device=dev, dtype=f64)
        xe = torch.sort(xe, dim=1).values
        so = {}
        splev = {"config": {"workload": "splev bulk evaluation: %d k=%d splines (n=%d) x %d sorted "
                                        "points per GPU (config 5 shape)" % (ns, ks, nn, ms)},
                 "algorithmic_bytes_per_point": 16.0 + 8.0 * (nn + nn - ks - 1) / ms}
        for mode, name in ((1, "fast"), (0, "exact")):
            for _ in range(max(1, args.warmup)):
                r = eng.splev_batch(t, nvec, c, ks, xe, e=0, mode=mode, out=so)
                so = {"y": r["y"], "ier": r["ier"]}
            barrier()
            ev0.record()
            for _ in range(args.steps):
                r = eng.splev_batch(t, nvec, c, ks, xe, e=0, mode=mode, out=so)
            ev1.record()
            barrier()
            ms_step = max_over_ranks(ev0.elapsed_time(ev1) / args.steps)
            kms = eng.last_kernel_ms()
            bytes_alg = ns * ms * splev["algorithmic_bytes_per_point"]
            splev[name] = {
                "value": world * bytes_alg / (ms_step * 1e-3) / 1e9, "unit": "GB/s",
                "gpts_per_s": world * ns * ms / (ms_step * 1e-3) / 1e9, "ms_per_step": ms_step,
                "roofline": {"bound": "hbm", "kernel": ("splev_fast_kernel<5>" if mode == 1 else "splev_kernel<5> (exact)"),
                             "achieved": bytes_alg / (kms * 1e-3) / 1e9, "peak": hbm_peak,
                             "unit": "GB/s", "frac": bytes_alg / (kms * 1e-3) / 1e9 / hbm_peak,
                             "traffic": (NCU_SPLEV_DRAM_BYTES_PER_POINT * ns * ms) if mode == 1 else None,
                             "peak_source": peak_src}}
        for name in ("fast", "exact"):
            rr = dict(splev[name]["roofline"])
            rr.update({"value_gbs": splev[name]["value"], "gpts_per_s": splev[name]["gpts_per_s"],
                       "ms_per_step": splev[name]["ms_per_step"], "workload": splev["config"]["workload"]})
            roof["secondary"].append(rr)
        splev["gpu_launches"] = 2 * args.steps
        del t, c, nvec, xe, so, r
        torch.cuda.empty_cache()

    # ---------------------------------------------------------------- shared-abscissa LSQ (config 3 shape)
    shared = None
    if not args.no_shared:
        torch.cuda.empty_cache()
        f64 = torch.float64
        Bs, ms, ks, nint = args.shared_curves, SHARED_M, SHARED_K, SHARED_NINT
        gsh = torch.Generator(device=dev)
        gsh.manual_seed(7 + rank)
        xs_ = torch.linspace(0, 1, ms, device=dev, dtype=f64)
        nn = 2 * (ks + 1) + nint
        tsh = torch.zeros(nn, device=dev, dtype=f64)
        tsh[ks + 1:ks + 1 + nint] = torch.linspace(0, 1, nint + 2, device=dev, dtype=f64)[1:-1]
        ysh = torch.empty((ms, Bs), device=dev, dtype=f64)   # point-major; 68.7 GB at 2^24 curves
        for r0 in range(0, ms, 64):                           # filled in row blocks (no second 68 GB temporary)
            ysh[r0:r0 + 64].normal_(generator=gsh)
        sho = {}
        nk1 = nn - ks - 1
        bytes_curve = 8 * ms + 8 * nk1 + 12
        flops_curve = ms * (1 + 6 * (ks + 1) + 2) + nk1 * (2 * ks + 1)
        shared = {"config": {"workload": "shared-abscissa least squares (iopt=-1): %d curves x %d points per "
                                         "GPU, k=3, %d interior knots (config 3%s)"
                                         % (Bs, ms, nint, " at its stated size" if Bs == 1 << 24 else " shape")}}
        l0 = eng.launches
        for smode, sname in ((0, "exact"), (1, "fast")):
            for _ in range(max(1, args.warmup)):
                r = eng.curfit_shared(xs_, ysh, tsh, k=ks, out=sho, mode=smode)
                sho = {"c": r["c"], "fp": r["fp"], "ier": r["ier"]}
            barrier()
            ev0.record()
            for _ in range(args.steps):
                r = eng.curfit_shared(xs_, ysh, tsh, k=ks, out=sho, mode=smode)
            ev1.record()
            barrier()
            ms_step = max_over_ranks(ev0.elapsed_time(ev1) / args.steps)
            kms = eng.last_kernel_ms()
            assert int(r["ier"].cpu()[0]) == 0
            blk = {"value": world * Bs / (ms_step * 1e-3), "unit": "fits/s", "ms_per_step": ms_step,
                   "arithmetic": "reference rotations, 4 mul + 2 add, bit-identical to curfit(iopt=-1)"
                                 if smode == 0 else "FMA rotations (2 mul + 2 fma), within 1e-12 of exact",
                   "roofline": {"bound": "hbm", "kernel": "shared_apply2_kernel<3,%s>" % ("FMA" if smode else "exact"),
                                "achieved": Bs * bytes_curve / (kms * 1e-3) / 1e9, "peak": hbm_peak,
                                "unit": "GB/s", "frac": Bs * bytes_curve / (kms * 1e-3) / 1e9 / hbm_peak,
                                "traffic": NCU_SHARED_DRAM_BYTES_PER_CURVE * Bs, "peak_source": peak_src,
                                "algorithmic_bytes_per_curve": bytes_curve},
                   "fp64_gflops": Bs * (flops_curve if smode == 0 else flops_curve * 17.0 / 27.0) / (kms * 1e-3) / 1e9}
            if smode == 0:
                shared.update(blk)   # exact mode is the headline of this block
            else:
                shared["fast"] = blk
        shared["gpu_launches"] = int(eng.launches - l0)
        for blk in (shared, shared["fast"]):
            rr = dict(blk["roofline"])
            rr.update({"value_fits_per_s": blk["value"], "ms_per_step": blk["ms_per_step"],
                       "workload": shared["config"]["workload"]})
            roof["secondary"].append(rr)
        if rank == 0 and world == 1 and not args.no_cpu:
            # CPU baseline of this block + parity on the SAME curves: the first and the last 4096 curves of the
            # batch (the last ones sit 2^33 elements into y: an index-width error would show there) re-fitted by
            # the oracle's per-curve curfit(iopt=-1); exact mode must agree bit for bit
            import oracle_lib as O
            nthr_ = host_threads()
            eng.curfit_shared(xs_, ysh, tsh, k=ks, out=sho, mode=0)
            torch.cuda.synchronize()
            ns_ = min(4096, Bs // 2)
            sel = torch.cat([torch.arange(ns_, device=dev), torch.arange(Bs - ns_, Bs, device=dev)])
            y_h = np.ascontiguousarray(ysh[:, sel].t().cpu().numpy())
            t_h = np.zeros((2 * ns_, nn))
            t_h[:] = tsh.cpu().numpy()
            n_h = np.full(2 * ns_, nn, np.int32)
            t0 = time.perf_counter()
            oc = O.curfit_batch(-1, xs_.cpu().numpy(), y_h, None, ks, 0.0, nn, xb=0.0, xe=1.0, shared_x=True,
                                n=n_h, t=t_h, nthreads=nthr_)
            dts_ = time.perf_counter() - t0
            c_g = sho["c"][sel].cpu().numpy()
            fp_g = sho["fp"][sel].cpu().numpy()
            shared["cpu_baseline"] = {
                "value": 2 * ns_ / dts_, "unit": "fits/s", "cores": nthr_, "kind": "port",
                "sample": "first and last %d curves of the same batch, oracle curfit(iopt=-1) per curve, OpenMP over "
                          "curves; c and fp bit-identical to the GPU result (exact mode): %s"
                          % (ns_, bool(np.array_equal(c_g[:, :nk1], oc["c"][:, :nk1]) and np.array_equal(fp_g, oc["fp"])))}
        del ysh, sho, r
        torch.cuda.empty_cache()

    # ---------------------------------------------------------------- parametric curves
    parc = None
    if not args.no_parcur:
        torch.cuda.empty_cache()
        f64 = torch.float64
        Bp, mp, dp, kp, sig = PARCUR_B, PARCUR_M, PARCUR_IDIM, 3, 0.05
        gp = torch.Generator(device=dev)
        gp.manual_seed(21 + rank)
        th = torch.sort(torch.rand((mp, Bp), generator=gp, device=dev, dtype=f64) * 6.283185307179586, dim=0).values
        th = th + torch.arange(mp, device=dev, dtype=f64)[:, None] * 1e-9
        fr = 0.5 + 2.5 * torch.rand((1, dp, Bp), generator=gp, device=dev, dtype=f64)
        ph = 6.0 * torch.rand((1, dp, Bp), generator=gp, device=dev, dtype=f64)
        am = 0.5 + 1.5 * torch.rand((1, dp, Bp), generator=gp, device=dev, dtype=f64)
        xp_ = (am * torch.cos(fr * th[:, None, :] + ph)
               + sig * torch.randn((mp, dp, Bp), generator=gp, device=dev, dtype=f64)).contiguous()
        del th, fr, ph, am
        sp_ = mp * dp * sig * sig
        nestp_ = mp + kp + 1
        for _ in range(max(1, min(args.warmup, 2))):
            rp = eng.parcur_batch(xp_, None, k=kp, s=sp_, nest=nestp_, ipar=0, want_stats=True)
        barrier()
        l0 = eng.launches
        ev0.record()
        for _ in range(args.steps):
            rp = eng.parcur_batch(xp_, None, k=kp, s=sp_, nest=nestp_, ipar=0, want_stats=True)
        ev1.record()
        barrier()
        p_ms = max_over_ranks(ev0.elapsed_time(ev1) / args.steps)
        p_launch = int(eng.launches - l0)
        stp = rp["stats"].to(f64)
        n_p = rp["n"].to(f64)
        okp = float((rp["ier"] <= 0).to(f64).mean().item())
        # curev of every fitted curve at its own parameters: 8 B read + 8*idim B written per point
        ud = rp["u"].t().contiguous()
        for _ in range(2):
            xev, iev = eng.curev_batch(dp, rp["t"], rp["n"], rp["c"], kp, ud)
        barrier()
        ev0.record()
        for _ in range(args.steps):
            xev, iev = eng.curev_batch(dp, rp["t"], rp["n"], rp["c"], kp, ud, out=(xev, iev))
        ev1.record()
        barrier()
        c_ms = max_over_ranks(ev0.elapsed_time(ev1) / args.steps)
        # FAST mode: one fast-splev pass per coordinate (local polynomials + Horner)
        for _ in range(2):
            xfv, ifv = eng.curev_batch(dp, rp["t"], rp["n"], rp["c"], kp, ud, mode=1)
        barrier()
        ev0.record()
        for _ in range(args.steps):
            xfv, ifv = eng.curev_batch(dp, rp["t"], rp["n"], rp["c"], kp, ud, out=(xfv, ifv), mode=1)
        ev1.record()
        barrier()
        cf_ms = max_over_ranks(ev0.elapsed_time(ev1) / args.steps)
        cscale = torch.clamp(rp["c"].abs().amax(dim=1), min=1.0)[:, None, None]
        cf_err = float(((xfv - xev).abs() / cscale).max().item())
        del xfv
        cur_bytes = Bp * mp * 8.0 * (1 + dp)
        pbytes = Bp * (8.0 * mp * (dp + 1) + 8.0 * float(n_p.mean().item()) * (dp + 1) + 16)
        parc = {
            "config": {"workload": "batched parametric smoothing fit parcur: %d curves x %d points x idim=%d "
                                   "per GPU, k=3, ipar=0 (chord-length parameters on the device), "
                                   "s=m*idim*sigma^2=%.2f, nest=%d; then curev of every curve at its "
                                   "data parameters" % (Bp, mp, dp, sp_, nestp_)},
            "value": world * Bp / (p_ms * 1e-3), "unit": "fits/s", "ms_per_step": p_ms,
            "success_fraction": okp, "mean_n": float(n_p.mean().item()),
            "mean_lsq_passes": float(stp[:, 0].mean().item()), "mean_p_iters": float(stp[:, 1].mean().item()),
            "gpu_launches": p_launch,
            "roofline": {"bound": "hbm", "kernel": "fit pipeline instantiated for idim=%d (curfit_pass_kernel<3,.,%d> "
                         "x R + curfit_part2_kernel): FP64 issue/latency bound like config 2, HBM fraction "
                         "small by construction" % (dp, dp),
                         "achieved": pbytes / (p_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                         "frac": pbytes / (p_ms * 1e-3) / 1e9 / hbm_peak, "traffic": None,
                         "peak_source": peak_src, "algorithmic_bytes_per_fit": pbytes / Bp},
            "curev": {"ms": c_ms, "value": world * cur_bytes / (c_ms * 1e-3) / 1e9, "unit": "GB/s",
                      "gpts_per_s": world * Bp * mp / (c_ms * 1e-3) / 1e9,
                      "roofline": {"bound": "hbm", "kernel": "curev_kernel<3> (exact arithmetic)",
                                   "achieved": cur_bytes / (c_ms * 1e-3) / 1e9, "peak": hbm_peak,
                                   "unit": "GB/s", "frac": cur_bytes / (c_ms * 1e-3) / 1e9 / hbm_peak,
                                   "traffic": None, "peak_source": peak_src,
                                   "algorithmic_bytes_per_point": 8.0 * (1 + dp)}},
            "curev_fast": {"ms": cf_ms, "value": world * cur_bytes / (cf_ms * 1e-3) / 1e9, "unit": "GB/s",
                           "gpts_per_s": world * Bp * mp / (cf_ms * 1e-3) / 1e9,
                           "max_err_vs_exact_rel_to_max_c": cf_err,
                           "roofline": {"bound": "hbm", "kernel": "splev_fast_kernel<3> x idim passes (curev FAST)",
                                        "achieved": cur_bytes / (cf_ms * 1e-3) / 1e9, "peak": hbm_peak,
                                        "unit": "GB/s", "frac": cur_bytes / (cf_ms * 1e-3) / 1e9 / hbm_peak,
                                        "traffic": None, "peak_source": peak_src,
                                        "algorithmic_bytes_per_point": 8.0 * (1 + dp)}}}
        # dense evaluation (many points per knot interval), compact tables: where FAST is meant to be used
        Bd, md = min(Bp, 1 << 16), 2048
        nmx = int(rp["n"][:Bd].max().item())
        nmx += nmx & 1
        td_ = rp["t"][:Bd, :nmx].contiguous()
        cd_ = rp["c"][:Bd, :dp * nmx].contiguous()
        nd_ = rp["n"][:Bd].contiguous()
        udn = torch.linspace(0, 1, md, device=dev, dtype=f64)[None, :].expand(Bd, md).contiguous()
        dense = {}
        for mode_, nm_ in ((0, "exact"), (1, "fast")):
            xo_, io_ = eng.curev_batch(dp, td_, nd_, cd_, kp, udn, mode=mode_)
            barrier()
            ev0.record()
            for _ in range(args.steps):
                xo_, io_ = eng.curev_batch(dp, td_, nd_, cd_, kp, udn, out=(xo_, io_), mode=mode_)
            ev1.record()
            barrier()
            d_ms = max_over_ranks(ev0.elapsed_time(ev1) / args.steps)
            dense[nm_] = {"ms": d_ms, "gpts_per_s": world * Bd * md / (d_ms * 1e-3) / 1e9,
                          "value": world * Bd * md * 8.0 * (1 + dp) / (d_ms * 1e-3) / 1e9, "unit": "GB/s",
                          "frac_of_hbm_peak": Bd * md * 8.0 * (1 + dp) / (d_ms * 1e-3) / 1e9 / hbm_peak}
            if mode_ == 0:
                xex_ = xo_
            else:
                dense["max_err_vs_exact_rel_to_max_c"] = float(((xo_ - xex_).abs() / cscale[:Bd]).max().item())
        dense["workload"] = "%d fitted idim=%d curves x %d points each, compact tables (ldt=%d)" % (Bd, dp, md, nmx)
        parc["curev_dense"] = dense
        del xo_, xex_, udn
        for key_ in ("curev", "curev_fast"):
            rr = dict(parc[key_]["roofline"])
            rr.update({"gpts_per_s": parc[key_]["gpts_per_s"], "ms_per_step": parc[key_]["ms"],
                       "workload": "curev of %d fitted idim=%d curves x %d points" % (Bp, dp, mp)})
            roof["secondary"].append(rr)
        roof["secondary"].append({"bound": "fp64-chain", "kernel": parc["roofline"]["kernel"], "unit": "GB/s",
                                  "achieved": parc["roofline"]["achieved"], "peak": hbm_peak,
                                  "frac": parc["roofline"]["frac"], "traffic": None, "peak_source": peak_src,
                                  "value_fits_per_s": parc["value"], "ms_per_step": p_ms,
                                  "workload": parc["config"]["workload"]})
        roof["secondary"].append({"bound": "hbm", "kernel": "curev FAST, dense evaluation", "unit": "GB/s",
                                  "achieved": dense["fast"]["value"] / world, "peak": hbm_peak,
                                  "frac": dense["fast"]["frac_of_hbm_peak"], "traffic": None,
                                  "peak_source": peak_src, "exact_gbs": dense["exact"]["value"] / world,
                                  "workload": dense["workload"]})
        if rank == 0 and world == 1 and not args.no_cpu:
            import oracle_lib as O
            nthr_ = host_threads()
            ns_ = 8192
            xs_h = np.ascontiguousarray(xp_[:, :, :ns_].permute(2, 0, 1).cpu().numpy())
            t0 = time.perf_counter()
            oc = O.parcur_batch(0, 0, dp, xs_h, None, kp, sp_, nestp_, nthreads=nthr_)
            dtp = time.perf_counter() - t0
            same = bool(np.array_equal(oc["n"], rp["n"][:ns_].cpu().numpy())
                        and np.array_equal(oc["fp"], rp["fp"][:ns_].cpu().numpy()))
            t0 = time.perf_counter()
            xo_, io_ = O.curev_batch(dp, oc["t"], oc["n"], oc["c"], kp, oc["u"], nthreads=nthr_)
            dtc_ = time.perf_counter() - t0
            same_ev = bool(np.array_equal(xo_, xev[:ns_].cpu().numpy()))
            parc["curev"]["cpu_baseline"] = {
                "value": ns_ * mp * 8.0 * (1 + dp) / dtc_ / 1e9, "unit": "GB/s", "gpts_per_s": ns_ * mp / dtc_ / 1e9,
                "cores": nthr_, "kind": "port",
                "sample": "%d curves x %d points, oracle curev, OpenMP over curves; values bit-identical to the "
                          "GPU result: %s" % (ns_, mp, same_ev)}
            parc["cpu_baseline"] = {"value": ns_ / dtp, "unit": "fits/s", "cores": nthr_, "kind": "port",
                                    "sample": "first %d curves of the same batch, oracle port of parcur/fppara, "
                                              "OpenMP over curves; n and fp bit-identical to the GPU result: %s"
                                              % (ns_, same)}
        del xp_, rp, xev, ud
        torch.cuda.empty_cache()

    # ---------------------------------------------------------------- periodic curves
    perc = None
    if not args.no_percur:
        torch.cuda.empty_cache()
        f64 = torch.float64
        Bq, mq, kq, sigq = PERCUR_B, PERCUR_M, 3, 0.1
        gq = torch.Generator(device=dev)
        gq.manual_seed(31 + rank)
        xq = torch.sort(torch.rand((mq, Bq), generator=gq, device=dev, dtype=f64) * 6.283185307179586, dim=0).values
        xq = (xq + torch.arange(mq, device=dev, dtype=f64)[:, None] * 1e-9).contiguous()
        frq = torch.randint(1, 4, (1, Bq), generator=gq, device=dev).to(f64)
        yq = ((0.5 + 1.5 * torch.rand((1, Bq), generator=gq, device=dev, dtype=f64))
              * torch.sin(frq * xq + 6.0 * torch.rand((1, Bq), generator=gq, device=dev, dtype=f64))
              + sigq * torch.randn((mq, Bq), generator=gq, device=dev, dtype=f64)).contiguous()
        sq_ = mq * sigq * sigq
        nestq = mq + 2 * kq
        for _ in range(max(1, min(args.warmup, 2))):
            rq = eng.percur_batch(xq, yq, None, k=kq, s=sq_, nest=nestq)
        barrier()
        lq0 = eng.launches
        ev0.record()
        for _ in range(args.steps):
            rq = eng.percur_batch(xq, yq, None, k=kq, s=sq_, nest=nestq)
        ev1.record()
        barrier()
        q_ms = max_over_ranks(ev0.elapsed_time(ev1) / args.steps)
        q_launch = int(eng.launches - lq0)
        nq = rq["n"].to(f64)
        qbytes = Bq * (2 * 8.0 * mq + 16.0 * float(nq.mean().item()) + 16)
        perc = {
            "config": {"workload": "batched periodic smoothing fit percur: %d curves x %d points per GPU, k=3, "
                                   "s=m*sigma^2=%.2f, nest=%d" % (Bq, mq, sq_, nestq)},
            "value": world * Bq / (q_ms * 1e-3), "unit": "fits/s", "ms_per_step": q_ms,
            "success_fraction": float((rq["ier"] <= 0).to(f64).mean().item()),
            "mean_n": float(nq.mean().item()), "gpu_launches": q_launch,
            "roofline": {"bound": "hbm", "kernel": "percur pipeline: percur_pass_kernel<3> x R + percur_part2_kernel<3> "
                         "(one thread per curve, re-bucketed per knot set): FP64 issue/latency bound (two-block "
                         "Givens chains), HBM fraction small by construction", "achieved": qbytes / (q_ms * 1e-3) / 1e9, "peak": hbm_peak,
                         "unit": "GB/s", "frac": qbytes / (q_ms * 1e-3) / 1e9 / hbm_peak, "traffic": None,
                         "peak_source": peak_src, "algorithmic_bytes_per_fit": qbytes / Bq}}
        roof["secondary"].append({"bound": "fp64-chain", "kernel": perc["roofline"]["kernel"], "unit": "GB/s",
                                  "achieved": perc["roofline"]["achieved"], "peak": hbm_peak,
                                  "frac": perc["roofline"]["frac"], "traffic": None, "peak_source": peak_src,
                                  "value_fits_per_s": perc["value"], "ms_per_step": q_ms,
                                  "workload": perc["config"]["workload"]})
        if rank == 0 and world == 1 and not args.no_cpu:
            import oracle_lib as O
            nthr_ = host_threads()
            ns_ = 8192
            xh_ = np.ascontiguousarray(xq[:, :ns_].t().cpu().numpy())
            yh_ = np.ascontiguousarray(yq[:, :ns_].t().cpu().numpy())
            t0 = time.perf_counter()
            oq = O.percur_batch(0, xh_, yh_, None, kq, sq_, nestq, nthreads=nthr_)
            dtq = time.perf_counter() - t0
            same = bool(np.array_equal(oq["n"], rq["n"][:ns_].cpu().numpy())
                        and np.array_equal(oq["fp"], rq["fp"][:ns_].cpu().numpy()))
            perc["cpu_baseline"] = {"value": ns_ / dtq, "unit": "fits/s", "cores": nthr_, "kind": "port",
                                    "sample": "first %d curves of the same batch, oracle port of percur/fpperi, "
                                              "OpenMP over curves; n and fp bit-identical to the GPU result: %s"
                                              % (ns_, same)}
        del xq, yq, rq
        torch.cuda.empty_cache()

    # ---------------------------------------------------------------- gridded surface (configs 4, 5b)
    grid = None
    if not args.no_regrid:
        torch.cuda.empty_cache()
        gm, gk = GRID_M, GRID_K
        gx = np.linspace(-1.5, 1.5, gm)
        gg = torch.Generator(device=dev)
        gg.manual_seed(4 + rank)
        GX = torch.from_numpy(gx).to(dev)[:, None]
        GY = torch.from_numpy(gx).to(dev)[None, :]
        gz = (torch.exp(-(GX ** 2 + GY ** 2)) + 0.3 * torch.sin(2 * GX) * torch.cos(1.5 * GY)
              + GRID_SIGMA * torch.randn((gm, gm), generator=gg, device=dev, dtype=torch.float64))
        gs = gm * gm * GRID_SIGMA ** 2

        def grid_fit():
            return fitpack_b200.regrid(0, gx, gx, gz, -1.5, 1.5, -1.5, 1.5, gk, gk, gs, GRID_NEST,
                                       GRID_NEST, engine=eng)
        for _ in range(max(1, min(args.warmup, 2))):
            rg = grid_fit()
        gsteps = max(1, min(args.steps, 3))
        barrier()
        l0 = eng.launches
        ev0.record()
        for _ in range(gsteps):
            rg = grid_fit()
        ev1.record()
        barrier()
        g_ms = max_over_ranks(ev0.elapsed_time(ev1) / gsteps)
        ncalls = int(eng.grid_fpgrre_calls)
        g_launch = int(eng.launches - l0)
        gout = torch.empty((gm, gm), device=dev, dtype=torch.float64)
        for _ in range(3):
            fitpack_b200.bispev(rg["tx"], rg["nx"], rg["ty"], rg["ny"], rg["c"], gk, gk, gx, gx,
                                out=gout, engine=eng)
        torch.cuda.synchronize()
        ev_ms = eng.last_kernel_ms()
        # FAST mode of bispev (FMA sums)
        gfast = torch.empty((gm, gm), device=dev, dtype=torch.float64)
        for _ in range(3):
            fitpack_b200.bispev(rg["tx"], rg["nx"], rg["ty"], rg["ny"], rg["c"], gk, gk, gx, gx,
                                out=gfast, engine=eng, mode=1)
        torch.cuda.synchronize()
        evf_ms = eng.last_kernel_ms()
        bispev_fast_err = float(((gfast - gout).abs().max() / max(1.0, float(np.abs(rg["c"]).max()))).item())
        del gfast
        resid = float(((gz - gout) ** 2).sum().item())
        assert rg["ier"] == 0 and abs(resid - rg["fp"]) <= 1e-9 * rg["fp"], "regrid fp != residual of bispev"
        # algorithmic bytes of one fpgrre solve: z is read twice (x reduction, residual pass)
        g_bytes = 2 * 8 * gm * gm
        grid = {
            "config": {"workload": "gridded smoothing fit regrid: %d x %d grid in HBM, kx=ky=%d, "
                                   "s=mx*my*sigma^2=%.0f, nest=%d (config 4); then bispev on the "
                                   "same grid (config 5b)" % (gm, gm, gk, gs, GRID_NEST)},
            "value": world / (g_ms * 1e-3), "unit": "grid fits/s", "ms_per_fit": g_ms,
            "fpgrre_solves_per_fit": ncalls, "ms_per_fpgrre_solve": g_ms / max(ncalls, 1),
            "nx": rg["nx"], "ny": rg["ny"], "ier": rg["ier"], "gpu_launches": g_launch,
            "roofline": {"bound": "hbm", "kernel": "fpgrre solve (grid_plan + 2 x grid_apply + "
                         "grid_resid): latency bound by the sequential Givens chain of the band QR "
                         "(one rotation depends on the previous one), HBM fraction small by nature",
                         "achieved": ncalls * g_bytes / (g_ms * 1e-3) / 1e9, "peak": hbm_peak,
                         "unit": "GB/s", "frac": ncalls * g_bytes / (g_ms * 1e-3) / 1e9 / hbm_peak,
                         "traffic": None, "peak_source": peak_src,
                         "algorithmic_bytes_per_solve": g_bytes},
            "bispev": {"ms": ev_ms, "value": gm * gm * 8 / (ev_ms * 1e-3) / 1e9, "unit": "GB/s",
                       "roofline": {"bound": "hbm", "kernel": "grid_bispev_kernel",
                                    "achieved": gm * gm * 8 / (ev_ms * 1e-3) / 1e9, "peak": hbm_peak,
                                    "unit": "GB/s", "frac": gm * gm * 8 / (ev_ms * 1e-3) / 1e9 / hbm_peak,
                                    "traffic": None, "peak_source": peak_src,
                                    "algorithmic_bytes_per_point": 8}},
            "bispev_fast": {"ms": evf_ms, "value": gm * gm * 8 / (evf_ms * 1e-3) / 1e9, "unit": "GB/s",
                            "max_err_vs_exact_rel_to_max_c": bispev_fast_err,
                            "frac_of_hbm_peak": gm * gm * 8 / (evf_ms * 1e-3) / 1e9 / hbm_peak}}
        for key_, nm_ in (("bispev", "grid_bispev_kernel (exact)"), ("bispev_fast", "grid_bispev_kernel (FMA sums)")):
            roof["secondary"].append({"bound": "hbm", "kernel": nm_, "unit": "GB/s", "peak": hbm_peak,
                                      "achieved": grid[key_]["value"], "frac": grid[key_]["value"] / hbm_peak,
                                      "traffic": None, "peak_source": peak_src,
                                      "workload": "bispev on a %d x %d grid, kx=ky=%d" % (gm, gm, gk)})
        roof["secondary"].append({"bound": "latency", "kernel": "regrid: fpgrre solve (plan + 2 x apply + resid)",
                                  "ms_per_fit": g_ms, "fpgrre_solves_per_fit": ncalls,
                                  "achieved": grid["roofline"]["achieved"], "unit": "GB/s", "peak": hbm_peak,
                                  "frac": grid["roofline"]["frac"], "traffic": None, "peak_source": peak_src,
                                  "workload": grid["config"]["workload"]})
        if rank == 0 and world == 1 and not args.no_cpu:
            import oracle_lib as O
            cm = 1024   # bounded CPU sample: the same surface on a 1024 x 1024 grid
            cx = np.linspace(-1.5, 1.5, cm)
            czz = gz[::gm // cm, ::gm // cm].contiguous().cpu().numpy()
            t0 = time.perf_counter()
            oc = O.regrid(0, cx, cx, czz, -1.5, 1.5, -1.5, 1.5, gk, gk, cm * cm * GRID_SIGMA ** 2,
                          GRID_NEST, GRID_NEST)
            dtc = time.perf_counter() - t0
            czd = torch.from_numpy(czz).to(dev)
            t0 = time.perf_counter()
            og = fitpack_b200.regrid(0, cx, cx, czd, -1.5, 1.5, -1.5, 1.5, gk, gk,
                                     cm * cm * GRID_SIGMA ** 2, GRID_NEST, GRID_NEST, engine=eng)
            torch.cuda.synchronize()
            dtg = time.perf_counter() - t0
            grid["cpu_baseline"] = {
                "value": 1.0 / dtc, "unit": "grid fits/s", "cores": 1, "kind": "port",
                "sample": "one %d x %d regrid of the same surface (oracle port, serial like the "
                          "reference); this GPU path on the same sample: %.1f ms vs %.1f ms CPU; "
                          "same knots: %s" % (cm, cm, dtg * 1e3, dtc * 1e3,
                                              bool(oc["nx"] == og["nx"] and oc["ny"] == og["ny"]
                                                   and np.array_equal(oc["tx"], og["tx"])
                                                   and np.array_equal(oc["ty"], og["ty"])))}
        del gz, gout, GX, GY
        torch.cuda.empty_cache()

    # ---------------------------------------------------------------- CPU baseline (rank 0, N=1)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        import oracle_lib as O
        rng = np.random.default_rng(1234)
        nthr = host_threads()
        xs, ys = synth_host(np, 0, 65536, M)    # the first 65536 curves of the batch the GPU fitted
        t0 = time.perf_counter()
        O.curfit_batch(0, xs[:512], ys[:512], None, K, S, NEST, nthreads=nthr)
        rate = 512 / (time.perf_counter() - t0)
        sample = int(min(65536, max(512, rate * 12)))
        t0 = time.perf_counter()
        O.curfit_batch(0, xs[:sample], ys[:sample], None, K, S, NEST, nthreads=nthr)
        dt = time.perf_counter() - t0
        if splev is not None:
            # second half of the metric: splev of the CPU port on the config-5 shape (k=5, n=32,
            # 954 points per spline), OpenMP over splines, bounded sample
            ks_, nn_, ms_ = SPLEV_K, SPLEV_N, SPLEV_M
            nsp = 16384
            tt = np.zeros((nsp, nn_))
            ni_ = nn_ - 2 * (ks_ + 1)
            tt[:, ks_ + 1:ks_ + 1 + ni_] = np.linspace(0, 1, ni_ + 2)[1:-1]
            tt[:, ks_ + 1 + ni_:] = 1.0
            cc = rng.standard_normal((nsp, nn_))
            xx = np.sort(rng.uniform(0, 1, (nsp, ms_)), axis=1)
            nv = np.full(nsp, nn_, np.int32)
            O.splev_batch(tt[:256], nv[:256], cc[:256], ks_, xx[:256], nthreads=nthr)
            t0s = time.perf_counter()
            O.splev_batch(tt, nv, cc, ks_, xx, nthreads=nthr)
            dts = time.perf_counter() - t0s
            splev["cpu_baseline"] = {
                "value": nsp * ms_ * splev["algorithmic_bytes_per_point"] / dts / 1e9, "unit": "GB/s",
                "gpts_per_s": nsp * ms_ / dts / 1e9, "cores": nthr, "kind": "port",
                "sample": "%d k=5 splines x %d points, oracle splev (reference arithmetic), OpenMP over "
                          "splines" % (nsp, ms_)}
        cpu = {"value": sample / dt, "unit": "fits/s", "cores": nthr, "kind": "port",
               "sample": "first %d curves of the SAME batch (counter-based generator, inputs bit-identical to the "
                         "device batch), oracle port of the reference algorithm built -O2 -ffp-contract=off, OpenMP "
                         "over curves (reference is Fortran: no compiler in this image)" % sample}

    if rank == 0:
        line = {
            "metric": "curve fits/sec (256-pt, k=3)", "value": value, "unit": "fits/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": step_ms_max, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "batched smoothing fit: %d independent curves x %d points per "
                                   "GPU, k=3, fixed s=2.56, nest=359, per-curve knot insertion "
                                   "(config 2)" % (B, M),
                       "curves_per_gpu": B, "layout": "point-major SoA in HBM",
                       "l2": "inputs (%.1f GB/step) exceed the 126 MB L2; no flush needed"
                             % (2 * B * M * 8 / 1e9),
                       "success_fraction": ok_frac, "gathered_summary_curves": gathered,
                       "generator": "counter-based (tests/synthdata.py): curve b of rank r = global curve r*B+b, "
                                    "bit-identical on host and device",
                       "parity_sample": parity},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
            "roofline": roof, "fp64": fp64, "splev": splev, "shared_lsq": shared, "parcur": parc, "percur": perc, "regrid": grid,
            "cpu_baseline": cpu,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
