"""warp-per-curve kernel (fp_curfit_c): random cases against the CPU oracle, bit for bit, then the latency of one call"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import fitpack_b200 as F
import oracle_lib as O

def same(a, b, n, k):
    return (a["n"] == b["n"] and a["ier"] == b["ier"] and a["fp"] == b["fp"]
            and np.array_equal(a["t"][:n], b["t"][:n]) and np.array_equal(a["c"][:n - k - 1], b["c"][:n - k - 1], equal_nan=True))

rng = np.random.default_rng(11)
bad = 0
ncase = int(sys.argv[1]) if len(sys.argv) > 1 else 400
for trial in range(ncase):
    m = int(rng.integers(8, 400)); k = int(rng.integers(1, 6))
    if m < 2 * k + 2: m = 2 * k + 2
    x = np.sort(rng.uniform(0, 1, m)); x[0], x[-1] = 0.0, 1.0
    y = np.sin(2 * np.pi * rng.uniform(0.5, 5) * x) + 0.05 * rng.standard_normal(m)
    w = rng.uniform(0.5, 2.0, m) if trial % 3 else np.ones(m)
    if trial % 11 == 0: w[rng.integers(0, m)] = 0.0
    s = float(m * 0.0025 * rng.choice([0.0, 0.03, 0.3, 1.0, 3.0, 30.0]))
    nest = int(rng.choice([m + k + 1, max(2 * k + 2, m // 2), max(102, int(np.ceil(1.4 * m)))]))
    o = O.curfit(0, x, y, w, 0.0, 1.0, k, s, nest)
    g = F.curfit(0, x, y, w, 0.0, 1.0, k, s, nest)
    n = int(o["n"])
    ok = same(o, g, n, k) if o["ier"] != 10 and n <= nest else (o["ier"] == g["ier"] and o["n"] == g["n"])
    if ok and o["ier"] <= 0 and trial % 2 == 0:
        # continuation with a smaller s (iopt = 1), then least squares on the knots found (iopt = -1)
        s2 = s * 0.5
        o1 = O.curfit(1, x, y, w, 0.0, 1.0, k, s2, nest, n=o["n"], t=o["t"], c=o["c"], fp=o["fp"], wrk=o["wrk"], iwrk=o["iwrk"])
        g1 = F.curfit(1, x, y, w, 0.0, 1.0, k, s2, nest, n=g["n"], t=g["t"], c=g["c"], fp=g["fp"], wrk=g["wrk"], iwrk=g["iwrk"])
        ok = ok and same(o1, g1, int(o1["n"]), k)
        o2 = O.curfit(-1, x, y, w, 0.0, 1.0, k, s2, nest, n=o1["n"], t=o1["t"].copy())
        g2 = F.curfit(-1, x, y, w, 0.0, 1.0, k, s2, nest, n=g1["n"], t=g1["t"].copy())
        ok = ok and same(o2, g2, int(o2["n"]), k)
    if not ok:
        bad += 1
        if bad <= 8:
            print("MISMATCH trial %d m=%d k=%d s=%g nest=%d: oracle n=%d ier=%d fp=%r | gpu n=%d ier=%d fp=%r" % (
                trial, m, k, s, nest, o["n"], o["ier"], o["fp"], g["n"], g["ier"], g["fp"]))
            nn = min(int(o["n"]), int(g["n"]))
            dt = np.nonzero(o["t"][:nn] != g["t"][:nn])[0]
            dc = np.nonzero(o["c"][:nn - k - 1] != g["c"][:nn - k - 1])[0]
            print("   first differing t index %s, c index %s" % (dt[:3], dc[:3]))
print("cases %d, mismatches %d" % (ncase, bad))

def bench(fn, reps=200):
    fn(); fn()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    return (time.perf_counter() - t0) / reps * 1e3
rng = np.random.default_rng(0)
m = 256
x = np.sort(rng.uniform(0, 1, m)); y = np.sin(6 * x) + 0.1 * rng.standard_normal(m)
for nest in (359, 260):
    r = F.curfit(0, x, y, None, 0, 1, 3, 2.56, nest)
    on = bench(lambda: F.curfit(0, x, y, None, 0, 1, 3, 2.56, nest))
    F.lib().fpb_set_single_curve_kernel(0)
    off = bench(lambda: F.curfit(0, x, y, None, 0, 1, 3, 2.56, nest), 50)
    F.lib().fpb_set_single_curve_kernel(1)
    cpu = bench(lambda: O.curfit(0, x, y, None, 0, 1, 3, 2.56, nest))
    print("fp_curfit_c m=256 k=3 s=2.56 nest=%d (n=%d): warp-per-curve %.3f ms | batch kernels %.3f ms | CPU oracle %.3f ms" % (
        nest, r["n"], on, off, cpu))
for (mm, kk, ss) in ((1000, 3, 10.0), (64, 5, 0.05), (256, 3, 0.0)):
    x = np.sort(rng.uniform(0, 1, mm)); y = np.sin(6 * x) + 0.1 * rng.standard_normal(mm)
    nest = mm + kk + 1
    r = F.curfit(0, x, y, None, 0, 1, kk, ss, nest)
    print("m=%d k=%d s=%g (n=%d ier=%d): warp-per-curve %.3f ms | CPU oracle %.3f ms" % (
        mm, kk, ss, r["n"], r["ier"], bench(lambda: F.curfit(0, x, y, None, 0, 1, kk, ss, nest), 50),
        bench(lambda: O.curfit(0, x, y, None, 0, 1, kk, ss, nest), 50)))

# ---- small batches through fp_curfit_batch_c: warp-per-curve kernel (one CTA per curve) vs the batch kernels ----
import synthdata  # noqa: E402  (the bench workload: 256-point curves, k = 3, s = 2.56)
for B in (1, 16, 64, 256, 512, 1024, 2048, 4096, 8192):
    xs, ys = synthdata.c2_curves(np, 0, B, 256)
    o = O.curfit_batch(0, xs, ys, None, 3, 2.56, 260)
    res = {}
    for on in (1, 0):
        F.lib().fpb_set_single_curve_kernel(on)
        g = F.curfit_batch(0, xs, ys, None, 3, 2.56, 260)
        assert np.array_equal(g["n"], o["n"]) and np.array_equal(g["ier"], o["ier"]) and np.array_equal(g["fp"], o["fp"])
        assert all(np.array_equal(g["t"][b, :o["n"][b]], o["t"][b, :o["n"][b]]) and
                   np.array_equal(g["c"][b, :o["n"][b] - 4], o["c"][b, :o["n"][b] - 4]) for b in range(B))
        res[on] = bench(lambda: F.curfit_batch(0, xs, ys, None, 3, 2.56, 260), 20)
    F.lib().fpb_set_single_curve_kernel(1)
    print("fp_curfit_batch_c B=%5d: warp-per-curve %.3f ms | batch kernels %.3f ms | bit-identical to the oracle" % (B, res[1], res[0]))
