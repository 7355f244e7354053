"""One fit path alone, bench shapes.  usage: python profiles/scripts/curves_only.py KIND [log2 curves] [reps]
KIND: fit (config 2 curfit pipeline) | parcur (idim=2) | percur | curev"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import bench
from fitpack_b200.batch import Engine

kind = sys.argv[1]
B = 1 << (int(sys.argv[2]) if len(sys.argv) > 2 else 18)
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
dev = torch.device("cuda", 0)
eng = Engine(dev)
f64 = torch.float64
g = torch.Generator(device=dev); g.manual_seed(5)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
m = 256
if kind == "fit":
    x, y = bench.synth_device(torch, dev, 0, B, bench.M)
    run = lambda: eng.curfit_batch(x, y, None, k=3, s=bench.S, nest=bench.NEST, want_stats=True)
elif kind in ("parcur", "curev"):
    d = 2
    th = torch.sort(torch.rand((m, B), generator=g, device=dev, dtype=f64) * 6.283185307179586, dim=0).values
    th = th + torch.arange(m, device=dev, dtype=f64)[:, None] * 1e-9
    fr = 0.5 + 2.5 * torch.rand((1, d, B), generator=g, device=dev, dtype=f64)
    xx = (torch.cos(fr * th[:, None, :]) + 0.05 * torch.randn((m, d, B), generator=g, device=dev, dtype=f64)).contiguous()
    run = lambda: eng.parcur_batch(xx, None, k=3, s=m * d * 0.0025, nest=m + 4, ipar=0, want_stats=True)
    if kind == "curev":
        r = run()
        ud = r["u"].t().contiguous()
        run = lambda: eng.curev_batch(d, r["t"], r["n"], r["c"], 3, ud)
elif kind == "percur":
    xq = torch.sort(torch.rand((m, B), generator=g, device=dev, dtype=f64) * 6.283185307179586, dim=0).values
    xq = (xq + torch.arange(m, device=dev, dtype=f64)[:, None] * 1e-9).contiguous()
    yq = (torch.sin(2 * xq) + 0.1 * torch.randn((m, B), generator=g, device=dev, dtype=f64)).contiguous()
    run = lambda: eng.percur_batch(xq, yq, None, k=3, s=m * 0.01, nest=m + 6)
else:
    raise SystemExit("unknown kind")
run()
torch.cuda.synchronize()
for i in range(reps):
    e0.record(); run(); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print("%s: %d curves, %.2f ms -> %.3f M/s" % (kind, B, ms, B / ms / 1e3))
