"""fp_parcur_c on the warp-per-curve kernel (idim 2, 3): random cases against the CPU oracle, bit for bit; latency"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import fitpack_b200 as F
import oracle_lib as O


def same(o, g, idim, k):
    if (o["n"], o["ier"]) != (g["n"], g["ier"]):
        return False
    if not (np.array_equal(o["u"], g["u"]) and o["ub"] == g["ub"] and o["ue"] == g["ue"]):
        return False
    if o["ier"] == 10:
        return True
    n = int(o["n"])
    if o["fp"] != g["fp"] or not np.array_equal(o["t"][:n], g["t"][:n]):
        return False
    return all(np.array_equal(o["c"][d * n:d * n + n - k - 1], g["c"][d * n:d * n + n - k - 1], equal_nan=True)
               for d in range(idim))


rng = np.random.default_rng(21)
ncase = int(sys.argv[1]) if len(sys.argv) > 1 else 300
bad = 0
for trial in range(ncase):
    idim = int(rng.choice([2, 3])); k = int(rng.integers(1, 6))
    m = int(rng.integers(2 * k + 2, 300))
    ph = np.sort(rng.uniform(0, 1, m)) * rng.uniform(1, 6)
    x = np.stack([np.cos(ph), np.sin(ph), ph * 0.3][:idim], 1) + 0.03 * rng.standard_normal((m, idim))
    w = rng.uniform(0.5, 2.0, m) if trial % 3 else None
    s = float(m * 0.001 * rng.choice([0.0, 0.1, 1.0, 10.0, 1000.0]))
    nest = int(rng.choice([m + k + 1, max(2 * k + 2, m // 2)]))
    ipar = int(trial % 4 == 1)
    kw = {}
    if ipar:
        u = np.cumsum(rng.uniform(0.1, 1.0, m)); kw = dict(u=u, ub=float(u[0]) - (trial % 2) * 0.1, ue=float(u[-1]))
    o = O.parcur(0, ipar, idim, x, w, k, s, nest, **kw)
    g = F.parcur(0, ipar, idim, x, w, k, s, nest, **kw)
    ok = same(o, g, idim, k)
    if ok and o["ier"] <= 0 and s > 0 and trial % 2 == 0:
        s2 = s * 0.4
        o1 = O.parcur(1, ipar, idim, x, w, k, s2, nest, u=o["u"], ub=o["ub"], ue=o["ue"], n=o["n"], t=o["t"], wrk=o["wrk"], iwrk=o["iwrk"], c=o["c"])
        g1 = F.parcur(1, ipar, idim, x, w, k, s2, nest, u=g["u"], ub=g["ub"], ue=g["ue"], n=g["n"], t=g["t"], wrk=g["wrk"], iwrk=g["iwrk"], c=g["c"])
        ok = ok and same(o1, g1, idim, k)
        o2 = O.parcur(-1, ipar, idim, x, w, k, s2, nest, u=o1["u"], ub=o1["ub"], ue=o1["ue"], n=o1["n"], t=o1["t"])
        g2 = F.parcur(-1, ipar, idim, x, w, k, s2, nest, u=g1["u"], ub=g1["ub"], ue=g1["ue"], n=g1["n"], t=g1["t"])
        ok = ok and same(o2, g2, idim, k)
    if not ok:
        bad += 1
        if bad <= 8:
            print("MISMATCH trial %d idim=%d m=%d k=%d s=%g nest=%d ipar=%d: oracle n=%d ier=%d fp=%r | gpu n=%d ier=%d fp=%r" % (
                trial, idim, m, k, s, nest, ipar, o["n"], o["ier"], o["fp"], g["n"], g["ier"], g["fp"]))
print("cases %d, mismatches %d" % (ncase, bad))
# rejected inputs: unsorted parameters, non-positive weight, bad knots for iopt = -1
m, k, idim, nest = 60, 3, 2, 70
ph = np.linspace(0, 3, m); x = np.stack([np.cos(ph), np.sin(ph)], 1)
u = np.linspace(0, 1, m); ub_ = u.copy(); ub_[10] = ub_[9]
wb = np.ones(m); wb[5] = 0.0
tb = np.zeros(nest); tb[4:8] = [0.2, 0.2, 0.5, 0.7]
for name, a in (("u not increasing", dict(ipar=1, u=ub_, ub=0.0, ue=1.0)), ("w = 0", dict(ipar=0, w=wb)),
                ("coincident knots", dict(ipar=1, iopt=-1, u=u, ub=0.0, ue=1.0, n=12, t=tb))):
    iopt = a.pop("iopt", 0); ipar = a.pop("ipar"); w = a.pop("w", None)
    o = O.parcur(iopt, ipar, idim, x, w, k, 0.1, nest, **a)
    g = F.parcur(iopt, ipar, idim, x, w, k, 0.1, nest, **a)
    print("%-18s oracle ier %d gpu ier %d same outputs %s t ends %s" % (name, o["ier"], g["ier"], same(o, g, idim, k),
                                                                     np.array_equal(o["t"], g["t"])))

def bench(fn, reps=100):
    fn(); fn()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    return (time.perf_counter() - t0) / reps * 1e3
rng = np.random.default_rng(0)
m = 256
xx = np.sort(rng.uniform(0, 1, m))
xy = np.stack([np.cos(6 * xx), np.sin(6 * xx)], 1) + 0.05 * rng.standard_normal((m, 2))
on = bench(lambda: F.parcur(0, 0, 2, xy, None, 3, 1.28, 260))
F.lib().fpb_set_single_curve_kernel(0)
off = bench(lambda: F.parcur(0, 0, 2, xy, None, 3, 1.28, 260), 30)
F.lib().fpb_set_single_curve_kernel(1)
cpu = bench(lambda: O.parcur(0, 0, 2, xy, None, 3, 1.28, 260))
print("fp_parcur_c m=256 k=3 idim=2: warp-per-curve %.3f ms | batch kernels %.3f ms | CPU oracle %.3f ms" % (on, off, cpu))
