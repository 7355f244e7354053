"""small exercise of the parametric / periodic curve kernels (all degrees, idim 1..5, weights, s = 0,
tight nest, iopt = -1), meant to run under `compute-sanitizer --tool memcheck`"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import fitpack_b200 as F
from fitpack_b200.batch import Engine
from refdata import synth_param_curves
dev = torch.device("cuda", 0)
eng = Engine(dev)
rng = np.random.default_rng(9)
for k in (1, 2, 3, 4, 5):
    for idim in (1, 2, 3, 5):
        B, m = 67, 37
        th, x = synth_param_curves(rng, B, m, idim)
        w = rng.uniform(0.5, 2, (B, m))
        for (s, nest, ww, ipar) in ((m * idim * 0.003, m + k + 1, None, 0), (0.0, m + k + 1, w, 1), (1e-7, 2 * k + 4, w, 0)):
            xd = torch.from_numpy(np.ascontiguousarray(x.transpose(1, 2, 0))).to(dev)
            wd = None if ww is None else torch.from_numpy(np.ascontiguousarray(ww.T)).to(dev)
            kw = {}
            if ipar:
                kw = dict(u=torch.from_numpy(np.ascontiguousarray(th.T)).to(dev),
                          ub=torch.from_numpy(th[:, 0].copy()).to(dev), ue=torch.from_numpy(th[:, -1].copy()).to(dev))
            r = eng.parcur_batch(xd, wd, k=k, s=s, nest=nest, ipar=ipar, want_stats=True, **kw)
            ud = r["u"].t().contiguous()
            eng.curev_batch(idim, r["t"], r["n"], r["c"], k, ud)
    B, m = 67, 41
    xs = np.sort(rng.uniform(0, 6, (B, m)), axis=1) + np.arange(m)[None, :] * 1e-9
    ys = np.sin(xs) + 0.1 * rng.standard_normal((B, m))
    w = rng.uniform(0.5, 2, (B, m))
    for (s, nest, ww) in ((m * 0.01, m + 2 * k, None), (0.0, m + 2 * k, w), (1e-7, 2 * k + 4, w)):
        xd = torch.from_numpy(np.ascontiguousarray(xs.T)).to(dev)
        yd = torch.from_numpy(np.ascontiguousarray(ys.T)).to(dev)
        wd = None if ww is None else torch.from_numpy(np.ascontiguousarray(ww.T)).to(dev)
        eng.percur_batch(xd, yd, wd, k=k, s=s, nest=nest)
    # iopt = -1 with given interior knots through the flat entry points
    t = np.zeros(2 * k + 2 + 4); t[k + 1:k + 5] = np.quantile(xs[0], [0.2, 0.4, 0.6, 0.8])
    F.percur(-1, xs[0], ys[0], w[0], k, 0.0, t.size, n=t.size, t=t)
    th1, x1 = synth_param_curves(rng, 1, m, 2)
    t = np.zeros(2 * k + 2 + 3); t[k + 1:k + 4] = [0.25, 0.5, 0.75]
    F.parcur(-1, 0, 2, x1[0], None, k, 0.0, t.size, n=t.size, t=t)
torch.cuda.synchronize()
print("done")
