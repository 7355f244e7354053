"""Config-4-shaped regrid run (mx x my grid in HBM, kx=ky=3, s = mx*my*sigma^2) + bispev on the
same grid.  usage: python profiles/scripts/regrid_only.py [m] [reps] [check_oracle]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import fitpack_b200 as F
from fitpack_b200.batch import Engine

m = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
check = int(sys.argv[3]) if len(sys.argv) > 3 else 0
dev = torch.device("cuda", 0)
eng = Engine(dev)
sigma, k, nest = 0.05, 3, 128
x = np.linspace(-1.5, 1.5, m); y = np.linspace(-1.5, 1.5, m)
g = torch.Generator(device=dev); g.manual_seed(4)
X = torch.from_numpy(x).to(dev)[:, None]; Y = torch.from_numpy(y).to(dev)[None, :]
z = torch.exp(-(X ** 2 + Y ** 2)) + 0.3 * torch.sin(2 * X) * torch.cos(1.5 * Y) \
    + sigma * torch.randn((m, m), generator=g, device=dev, dtype=torch.float64)
s = m * m * sigma * sigma
for i in range(reps):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    r = F.regrid(0, x, y, z, -1.5, 1.5, -1.5, 1.5, k, k, s, nest, nest, engine=eng)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    print("regrid %dx%d: wall %.2f ms, device %.2f ms, fpgrre calls %d, nx=%d ny=%d ier=%d fp/s=%.6f" % (
        m, m, (t1 - t0) * 1e3, eng.last_kernel_ms(), eng.grid_fpgrre_calls, r["nx"], r["ny"], r["ier"], r["fp"] / s))
out = torch.empty((m, m), device=dev, dtype=torch.float64)
for i in range(reps):
    zz, ier = F.bispev(r["tx"], r["nx"], r["ty"], r["ny"], r["c"], k, k, x, y, out=out, engine=eng)
    torch.cuda.synchronize()
    ms = eng.last_kernel_ms()
    print("bispev %dx%d: kernel %.3f ms -> %.1f GB/s written" % (m, m, ms, m * m * 8 / ms / 1e6))
print("residual check: sum (z - s(x,y))^2 / fp = %.12f" % (float(((z - out) ** 2).sum()) / r["fp"]))
if check:
    import oracle_lib as O
    zh = z.cpu().numpy()
    t0 = time.perf_counter()
    o = O.regrid(0, x, y, zh, -1.5, 1.5, -1.5, 1.5, k, k, s, nest, nest)
    t1 = time.perf_counter()
    nc = (o["nx"] - k - 1) * (o["ny"] - k - 1)
    print("oracle: %.1f ms; nx,ny,ier = %d %d %d; same knots: %s; max |dc|/max|c| = %.3e; dfp/fp = %.3e" % (
        (t1 - t0) * 1e3, o["nx"], o["ny"], o["ier"],
        bool(o["nx"] == r["nx"] and o["ny"] == r["ny"] and np.array_equal(o["tx"][:o["nx"]], r["tx"][:o["nx"]])
             and np.array_equal(o["ty"][:o["ny"]], r["ty"][:o["ny"]])),
        np.abs(o["c"][:nc] - r["c"][:nc]).max() / np.abs(o["c"][:nc]).max(), abs(o["fp"] - r["fp"]) / o["fp"]))
