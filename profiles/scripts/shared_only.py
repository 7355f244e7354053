"""Config-3-shaped shared-abscissa least-squares run (2^22 curves x 512 points, k=3, 28 interior
knots), exact and fast modes.  usage: python profiles/scripts/shared_only.py [reps] [mode|-1 both]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from fitpack_b200.batch import Engine

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 4
which = int(sys.argv[2]) if len(sys.argv) > 2 else -1
Bs, ms, ks, nint = 1 << 22, 512, 3, 28
dev = torch.device("cuda", 0)
eng = Engine(dev)
f64 = torch.float64
g = torch.Generator(device=dev); g.manual_seed(7)
xs = torch.linspace(0, 1, ms, device=dev, dtype=f64)
nn = 2 * (ks + 1) + nint
t = torch.zeros(nn, device=dev, dtype=f64)
t[ks + 1:ks + 1 + nint] = torch.linspace(0, 1, nint + 2, device=dev, dtype=f64)[1:-1]
y = torch.randn((ms, Bs), generator=g, device=dev, dtype=f64)
nk1 = nn - ks - 1
bytes_curve = 8 * ms + 8 * nk1 + 12
for mode in ((0, 1) if which < 0 else (which,)):
    out, best = {}, 1e9
    for i in range(reps):
        r = eng.curfit_shared(xs, y, t, k=ks, out=out, mode=mode)
        out = {"c": r["c"], "fp": r["fp"], "ier": r["ier"]}
        torch.cuda.synchronize()
        best = min(best, eng.last_kernel_ms())
    print("shared LSQ mode=%d: best kernel ms %.3f -> %.1f GB/s (%.1f%% of 6561.6), %.0f M fits/s" % (
        mode, best, Bs * bytes_curve / best / 1e6, 100 * Bs * bytes_curve / best / 1e6 / 6561.6, Bs / best / 1e3))
