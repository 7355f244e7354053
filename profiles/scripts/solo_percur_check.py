"""fp_percur_c / small fp_percur_batch_c on the curve-per-CTA kernel (fpb_percur_solo.cu): oracle parity and timings"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import fitpack_b200 as F
import oracle_lib as O

rng = np.random.default_rng(31)
ncase = int(sys.argv[1]) if len(sys.argv) > 1 else 200
bad = 0
for trial in range(ncase):
    k = int(rng.integers(1, 6)); m = int(rng.integers(2 * k + 4, 260))
    x = np.sort(rng.uniform(0, 2 * np.pi, m)); x[0] = 0.0; x[-1] = 2 * np.pi
    y = np.sin(x * int(rng.integers(1, 5))) + 0.05 * rng.standard_normal(m); y[-1] = y[0]
    w = rng.uniform(0.5, 2.0, m) if trial % 3 else None
    s = float(m * 0.0025 * rng.choice([0.0, 0.1, 1.0, 10.0, 1000.0]))
    nest = int(rng.choice([m + 2 * k, max(2 * k + 4, m // 2)]))
    o = O.percur(0, x, y, w, k, s, nest)
    g = F.percur(0, x, y, w, k, s, nest)
    n = int(o["n"])
    ok = (o["n"], o["ier"]) == (g["n"], g["ier"]) and (o["ier"] == 10 or (o["fp"] == g["fp"] and np.array_equal(o["t"][:n], g["t"][:n])
          and np.array_equal(o["c"][:n - k - 1], g["c"][:n - k - 1], equal_nan=True)))
    if ok and o["ier"] <= 0 and s > 0 and trial % 2 == 0:
        o1 = O.percur(1, x, y, w, k, s * 0.4, nest, n=o["n"], t=o["t"], wrk=o["wrk"], iwrk=o["iwrk"], c=o["c"], fp=o["fp"])
        g1 = F.percur(1, x, y, w, k, s * 0.4, nest, n=g["n"], t=g["t"], wrk=g["wrk"], iwrk=g["iwrk"], c=g["c"], fp=g["fp"])
        n1 = int(o1["n"])
        ok = (o1["n"], o1["ier"], o1["fp"]) == (g1["n"], g1["ier"], g1["fp"]) and np.array_equal(o1["t"][:n1], g1["t"][:n1]) \
            and np.array_equal(o1["c"][:n1 - k - 1], g1["c"][:n1 - k - 1], equal_nan=True)
    if not ok:
        bad += 1
        if bad <= 6:
            print("MISMATCH trial %d m=%d k=%d s=%g nest=%d: oracle n=%d ier=%d fp=%r | gpu n=%d ier=%d fp=%r" % (
                trial, m, k, s, nest, o["n"], o["ier"], o["fp"], g["n"], g["ier"], g["fp"]))
print("cases %d, mismatches %d" % (ncase, bad))

def bench(fn, reps=50):
    fn(); fn()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    return (time.perf_counter() - t0) / reps * 1e3
rng = np.random.default_rng(0)
m = 256
x = np.sort(rng.uniform(0, 1, m)); y = np.sin(6 * x) + 0.1 * rng.standard_normal(m)
on = bench(lambda: F.percur(0, x * 6, y, None, 3, 2.56, 262))
F.lib().fpb_set_single_curve_kernel(0)
off = bench(lambda: F.percur(0, x * 6, y, None, 3, 2.56, 262), 20)
F.lib().fpb_set_single_curve_kernel(1)
cpu = bench(lambda: O.percur(0, x * 6, y, None, 3, 2.56, 262))
print("fp_percur_c m=256 k=3: curve-per-CTA kernel %.3f ms | batch kernels %.3f ms | CPU oracle %.3f ms" % (on, off, cpu))
for B in (16, 128, 512, 1024, 2048):
    th = np.sort(rng.uniform(0, 2 * np.pi, (B, 256)), axis=1); th[:, 0] = 0; th[:, -1] = 2 * np.pi
    yy = np.sin(2 * th) + 0.1 * rng.standard_normal(th.shape); yy[:, -1] = yy[:, 0]
    o = O.percur_batch(0, th, yy, None, 3, 2.56, 262)
    res = {}
    for sel in (1, 0):
        F.lib().fpb_set_single_curve_kernel(sel)
        g = F.percur_batch(0, th, yy, None, 3, 2.56, 262)
        assert np.array_equal(g["n"], o["n"]) and np.array_equal(g["fp"], o["fp"]) and np.array_equal(g["ier"], o["ier"])
        res[sel] = bench(lambda: F.percur_batch(0, th, yy, None, 3, 2.56, 262), 10)
    F.lib().fpb_set_single_curve_kernel(1)
    print("fp_percur_batch_c B=%5d: default %.3f ms | batch kernels %.3f ms" % (B, res[1], res[0]))
