"""latency of the flat (batch-of-one) entry points: fp_curfit_c, fp_splev_c, fp_percur_c, fp_parcur_c"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import fitpack_b200 as F
import oracle_lib as O
rng = np.random.default_rng(0)
m = 256
x = np.sort(rng.uniform(0, 1, m)); y = np.sin(6 * x) + 0.1 * rng.standard_normal(m)
xy = np.stack([np.cos(6 * x), np.sin(6 * x)], 1) + 0.05 * rng.standard_normal((m, 2))
def bench(fn, reps=50):
    fn(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    return (time.perf_counter() - t0) / reps * 1e3
r = F.curfit(0, x, y, None, 0, 1, 3, 2.56, 359)
print("fp_curfit_c  m=256 k=3: %.3f ms (GPU, batch of one)   oracle CPU: %.3f ms" % (
    bench(lambda: F.curfit(0, x, y, None, 0, 1, 3, 2.56, 359)), bench(lambda: O.curfit(0, x, y, None, 0, 1, 3, 2.56, 359))))
print("fp_splev_c   m=256     : %.3f ms   oracle CPU: %.3f ms" % (
    bench(lambda: F.splev(r["t"], r["n"], r["c"], 3, x)), bench(lambda: O.splev(r["t"], r["n"], r["c"], 3, x))))
print("fp_percur_c  m=256 k=3: %.3f ms   oracle CPU: %.3f ms" % (
    bench(lambda: F.percur(0, x * 6, y, None, 3, 2.56, 262)), bench(lambda: O.percur(0, x * 6, y, None, 3, 2.56, 262))))
print("fp_parcur_c  m=256 k=3 idim=2: %.3f ms   oracle CPU: %.3f ms" % (
    bench(lambda: F.parcur(0, 0, 2, xy, None, 3, 1.28, 260)), bench(lambda: O.parcur(0, 0, 2, xy, None, 3, 1.28, 260))))
