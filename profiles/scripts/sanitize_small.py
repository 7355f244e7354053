"""small end-to-end exercise of every kernel, meant to run under compute-sanitizer memcheck"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests"))
import numpy as np, torch
import fitpack_b200 as F
from fitpack_b200.batch import Engine
from refdata import synth_curves
dev = torch.device("cuda", 0)
eng = Engine(dev)
rng = np.random.default_rng(3)
for k in (1, 2, 3, 4, 5):
    B, m = 97, 61
    x, y = synth_curves(rng, B, m)
    w = rng.uniform(0.5, 2, (B, m))
    for (s, nest, ww) in ((0.3, m + k + 1, None), (0.0, m + k + 1, w), (1e-6, 20, w)):
        xd = torch.from_numpy(np.ascontiguousarray(x.T)).to(dev)
        yd = torch.from_numpy(np.ascontiguousarray(y.T)).to(dev)
        wd = None if ww is None else torch.from_numpy(np.ascontiguousarray(ww.T)).to(dev)
        r = eng.curfit_batch(xd, yd, wd, k=k, s=s, nest=nest, want_stats=True)
        xe = torch.from_numpy(rng.uniform(-0.1, 1.1, (B, 131))).to(dev)
        for mode in (0, 1):
            for e in (0, 1, 2, 3):
                eng.splev_batch(r["t"], r["n"], r["c"], k, xe, e=e, mode=mode, want_idx=(e == 1))
        xe2 = torch.from_numpy(rng.uniform(0, 1, (B, 600))).to(dev)
        eng.splev_batch(r["t"], r["n"], r["c"], k, xe2, e=0, mode=1)
    xs = np.linspace(0, 1, m)
    t = np.zeros(2 * (k + 1) + 5); t[k + 1:k + 6] = np.linspace(0, 1, 7)[1:-1]
    eng.curfit_shared(torch.from_numpy(xs).to(dev), yd, torch.from_numpy(t).to(dev), k=k)
torch.cuda.synchronize()
r = F.curfit(0, x[0], y[0], None, 0.0, 1.0, 3, 0.3, 80)
print(F.splev(r["t"], r["n"], r["c"], 3, np.linspace(0, 1, 5000))[1], "done")
