"""Config-5-shaped splev run (2^20 k=5 splines x 954 sorted points), FAST mode only: the command the
splev ncu captures attach to.  usage: python profiles/scripts/splev_only.py [reps] [mode] [nspl]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from fitpack_b200.batch import Engine

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
mode = int(sys.argv[2]) if len(sys.argv) > 2 else 1
ns = int(sys.argv[3]) if len(sys.argv) > 3 else 1 << 20
ms, ks, nn = 954, 5, 32
dev = torch.device("cuda", 0)
eng = Engine(dev)
eng.use_torch_stream()
g = torch.Generator(device=dev); g.manual_seed(99)
f64 = torch.float64
t = torch.zeros((ns, nn), device=dev, dtype=f64)
nint = nn - 2 * (ks + 1)
t[:, ks + 1:ks + 1 + nint] = torch.linspace(0, 1, nint + 2, device=dev, dtype=f64)[1:-1]
t[:, ks + 1 + nint:] = 1.0
c = torch.randn((ns, nn), generator=g, device=dev, dtype=f64)
nvec = torch.full((ns,), nn, device=dev, dtype=torch.int32)
xe = torch.sort(torch.rand((ns, ms), generator=g, device=dev, dtype=f64), dim=1).values
so = {}
best = 1e9
for i in range(reps):
    r = eng.splev_batch(t, nvec, c, ks, xe, e=0, mode=mode, out=so)
    so = {"y": r["y"], "ier": r["ier"]}
    torch.cuda.synchronize()
    best = min(best, eng.last_kernel_ms())
bytes_alg = ns * ms * (16.0 + 8.0 * (nn + nn - ks - 1) / ms)
print("splev mode=%d best kernel ms %.4f  -> %.1f GB/s (%.1f%% of 6561.6)" % (
    mode, best, bytes_alg / best / 1e6, 100 * bytes_alg / best / 1e6 / 6561.6))
