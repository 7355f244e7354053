#!/usr/bin/env python
"""Golden vectors from the reference ITSELF: replays a fixed case list through the flat C binding of a
gfortran build of /root/reference (oracle/_ref/libfitpack_ref.so, built by build_ref.sh) and writes
tests/golden/ref_*.npz.  tests/test_ref_golden.py (oracle, CPU) and tests/test_gpu_ref_golden.py (CUDA
path) consume those files when present and require BIT-identical n, t, c, fp, ier -- the pin to the Fortran
binary that the scipy goldens (netlib translation) cannot give.

    sh oracle/ref_recipe/build_ref.sh && python oracle/ref_recipe/make_ref_golden.py
    python oracle/ref_recipe/make_ref_golden.py --backend oracle --out /tmp/x   # plumbing self-test (no gfortran)

Case list (all inputs are regenerated deterministically by the tests, only outputs are stored):
  * ref_curfit.npz  every input of tests/golden/curfit_golden.npz (224 cases, k=1..5, all ier classes), the
                    mncurf sweep (test/fitpack_tests.f90:1138-1269) incl. its iopt=1 / iopt=-1 steps, and the
                    directed deviation cases of tests/deviation_cases.py (fpknot seeding, subnormal pivot,
                    m+1 knot sets, fpchec walk)
  * ref_c2.npz      65 536 config-2 curves (tests/synthdata.py, seed 1234, curves 0..65535, 256 points, k=3,
                    s=2.56, nest=359): n, ier, fp and a 64-bit FNV-1a hash of t(1:n) and c(1:n-k-1) per curve
"""
import argparse
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLD = os.path.join(ROOT, "tests", "golden")


def fnv1a64(buf):
    """FNV-1a over the bytes of a float64 array (vectorised over 8-byte words would change the hash: plain loop)"""
    h = 0xcbf29ce484222325
    for byte in np.ascontiguousarray(buf).view(np.uint8).tolist():
        h = ((h ^ byte) * 0x100000001b3) & 0xFFFFFFFFFFFFFFFF
    return h


# ----------------------------------------------------------------------------------------------------------
class RefLib:
    """the flat reference API (include/fitpack_core_c.h:69-132) of a shared library: same call shapes as
    fitpack_b200.core / tests/oracle_lib.py"""

    def __init__(self, path):
        from fitpack_b200._lib import SYMBOLS
        self.L = C.CDLL(path)
        for name in ("fp_curfit_c", "fp_splev_c", "fp_percur_c"):
            fn = getattr(self.L, name)
            fn.restype, fn.argtypes = SYMBOLS[name]

    @staticmethod
    def _p(a):
        return None if a is None else a.ctypes.data_as(C.c_void_p)

    def curfit(self, iopt, x, y, w, xb, xe, k, s, nest, n=0, t=None, wrk=None, iwrk=None, lwrk=None, c=None):
        x = np.ascontiguousarray(x, np.float64)
        y = np.ascontiguousarray(y, np.float64)
        m = x.size
        w = np.ones(m) if w is None else np.ascontiguousarray(w, np.float64)
        if lwrk is None:
            lwrk = max(m * (k + 1) + nest * (7 + 3 * k), 1) if wrk is None else wrk.size
        t = np.zeros(max(nest, 1)) if t is None else np.ascontiguousarray(t, np.float64)
        c = np.zeros(max(nest, 1)) if c is None else c
        wrk = np.zeros(max(lwrk, 1)) if wrk is None else wrk
        iwrk = np.zeros(max(nest, 1), np.int32) if iwrk is None else iwrk
        n_, fp_, ier_ = C.c_int32(int(n)), C.c_double(0.0), C.c_int32(0)
        self.L.fp_curfit_c(int(iopt), m, self._p(x), self._p(y), self._p(w), float(xb), float(xe), int(k), float(s),
                           int(nest), C.addressof(n_), self._p(t), self._p(c), C.addressof(fp_), self._p(wrk),
                           int(lwrk), self._p(iwrk), C.addressof(ier_))
        return dict(n=n_.value, t=t, c=c, fp=fp_.value, ier=ier_.value, wrk=wrk, iwrk=iwrk)

    def percur(self, iopt, x, y, w, k, s, nest, n=0, t=None, wrk=None, iwrk=None):
        x = np.ascontiguousarray(x, np.float64)
        y = np.ascontiguousarray(y, np.float64)
        m = x.size
        w = np.ones(m) if w is None else np.ascontiguousarray(w, np.float64)
        lwrk = m * (k + 1) + nest * (8 + 5 * k)
        t = np.zeros(max(nest, 1)) if t is None else np.ascontiguousarray(t, np.float64)
        c = np.zeros(max(nest, 1))
        wrk = np.zeros(max(lwrk, 1)) if wrk is None else wrk
        iwrk = np.zeros(max(nest, 1), np.int32) if iwrk is None else iwrk
        n_, fp_, ier_ = C.c_int32(int(n)), C.c_double(0.0), C.c_int32(0)
        self.L.fp_percur_c(int(iopt), m, self._p(x), self._p(y), self._p(w), int(k), float(s), int(nest),
                           C.addressof(n_), self._p(t), self._p(c), C.addressof(fp_), self._p(wrk), wrk.size,
                           self._p(iwrk), C.addressof(ier_))
        return dict(n=n_.value, t=t, c=c, fp=fp_.value, ier=ier_.value, wrk=wrk, iwrk=iwrk)

    def curfit_batch(self, iopt, x, y, w, k, s, nest):
        B, m = y.shape
        n = np.zeros(B, np.int32)
        t = np.zeros((B, nest))
        c = np.zeros((B, nest))
        fp = np.zeros(B)
        ier = np.zeros(B, np.int32)
        for b in range(B):
            r = self.curfit(iopt, x[b], y[b], None if w is None else w[b], x[b, 0], x[b, -1], k, s, nest)
            n[b], fp[b], ier[b] = r["n"], r["fp"], r["ier"]
            t[b], c[b] = r["t"], r["c"]
        return dict(n=n, t=t, c=c, fp=fp, ier=ier)


class OracleBackend:
    """plumbing self-test: the repo's own CPU oracle behind the same interface"""

    def __init__(self):
        import oracle_lib
        self.O = oracle_lib

    def curfit(self, *a, **kw):
        return self.O.curfit(*a, **kw)

    def percur(self, *a, **kw):
        return self.O.percur(*a, **kw)

    def curfit_batch(self, iopt, x, y, w, k, s, nest):
        return self.O.curfit_batch(iopt, x, y, w, k, s, nest)


# ----------------------------------------------------------------------------------------------------------
def curfit_cases():
    """(name, kwargs) list shared with the consuming tests: every call is reproducible from this list"""
    import deviation_cases as D
    from refdata import MNCURF_STEPS, MNCURF_X, MNCURF_Y
    cases = []
    g = np.load(os.path.join(GOLD, "curfit_golden.npz"))
    for i in range(g["k"].size):
        o, m = int(g["off"][i]), int(g["m"][i])
        cases.append(("golden%03d" % i, dict(iopt=0, x=g["x"][o:o + m], y=g["y"][o:o + m], w=g["w"][o:o + m],
                                             xb=float(g["xb"][i]), xe=float(g["xe"][i]), k=int(g["k"][i]),
                                             s=float(g["s"][i]), nest=int(g["nest"][i]))))
    for k in (3, 5):
        for step, (iopt, s) in enumerate(MNCURF_STEPS):
            cases.append(("mncurf_k%d_step%d" % (k, step), dict(iopt=iopt, x=MNCURF_X, y=MNCURF_Y, w=np.ones(25),
                                                                 xb=0.0, xe=24.0, k=k, s=s, nest=35,
                                                                 chain="mncurf_k%d" % k)))
    c = D.nan_case()
    cases.append(("dev_nan", dict(iopt=0, x=c["x"], y=c["y"], w=None, xb=c["xb"], xe=c["xe"], k=c["k"], s=c["s"],
                                  nest=c["nest"])))
    for wf in (0.0, 1e-310):
        c = D.subnormal_case(wf)
        cases.append(("dev_subnormal_%g" % wf, dict(iopt=-1, x=c["x"], y=c["y"], w=c["w"], xb=c["xb"], xe=c["xe"],
                                                    k=c["k"], s=0.0, nest=c["nest"], n=c["n"], t=c["t"])))
    c = D.spin_case()
    cases.append(("dev_spin_first", dict(iopt=0, x=c["x"], y=c["y"], w=None, xb=c["xb"], xe=c["xe"], k=c["k"],
                                         s=c["s_first"], nest=c["nest"], chain="spin")))
    cases.append(("dev_spin_second", dict(iopt=1, x=c["x"], y=c["y"], w=None, xb=c["xb"], xe=c["xe"], k=c["k"],
                                          s=c["s_second"], nest=c["nest"], chain="spin", zero_nrdata=True)))
    c = D.fpchec_case()
    t = np.zeros(8)
    t[:6] = c["t"]
    cases.append(("dev_fpchec", dict(iopt=-1, x=c["x"], y=np.array([1.0, 2.0, 3.0, 4.0]), w=None, xb=0.0, xe=10.0,
                                     k=c["k"], s=0.0, nest=8, n=c["n"], t=t)))
    return cases


def run_curfit_cases(backend):
    """replays curfit_cases(); `chain` cases carry t/n/wrk/iwrk from one step to the next (iopt=1)"""
    out, chains = {}, {}
    for name, kw in curfit_cases():
        kw = dict(kw)
        chain = kw.pop("chain", None)
        zero = kw.pop("zero_nrdata", False)
        st = chains.get(chain) if chain else None
        k = kw["k"]
        if st is not None and kw["iopt"] != 0:
            kw.update(n=st["n"], t=st["t"].copy(), wrk=st["wrk"].copy(), iwrk=st["iwrk"].copy())
            if kw["iopt"] == -1:   # mncurf's last step: 7 interior knots 3, 6, ..., 21
                kw["t"][k + 1:k + 8] = 3.0 * np.arange(1, 8)
                kw["n"] = 9 + 2 * k
            if zero:
                kw["iwrk"][:st["n"] - 2 * (k + 1) + 1] = 0
        elif chain and kw["iopt"] == -1:
            t = np.zeros(kw["nest"])
            t[k + 1:k + 8] = 3.0 * np.arange(1, 8)
            kw.update(n=9 + 2 * k, t=t)
        if chain:
            kw.setdefault("lwrk", 1000)
        x, y, w = kw.pop("x"), kw.pop("y"), kw.pop("w")
        r = backend.curfit(kw.pop("iopt"), x, y, w, kw.pop("xb"), kw.pop("xe"), k, kw.pop("s"), kw.pop("nest"),
                           **{a: kw[a] for a in ("n", "t", "wrk", "iwrk", "lwrk") if a in kw})
        if chain:
            chains[chain] = r
        n = r["n"]
        out[name + "_n"], out[name + "_ier"], out[name + "_fp"] = np.int32(n), np.int32(r["ier"]), np.float64(r["fp"])
        out[name + "_t"] = np.array(r["t"][:max(n, 0)])
        out[name + "_c"] = np.array(r["c"][:max(n - k - 1, 0)])
    return out


def run_c2(backend, ncurves):
    import synthdata
    x, y = synthdata.c2_curves(np, 0, ncurves, 256, 0.1, 1234)
    r = backend.curfit_batch(0, x, y, None, 3, 2.56, 359)
    ht = np.array([fnv1a64(r["t"][b, :r["n"][b]]) for b in range(ncurves)], np.uint64)
    hc = np.array([fnv1a64(r["c"][b, :r["n"][b] - 4]) for b in range(ncurves)], np.uint64)
    return dict(n=r["n"].astype(np.int16), ier=r["ier"].astype(np.int8), fp=r["fp"], hash_t=ht, hash_c=hc)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lib", default=os.path.join(ROOT, "oracle", "_ref", "libfitpack_ref.so"))
    ap.add_argument("--backend", default="ref", choices=["ref", "oracle"])
    ap.add_argument("--out", default=GOLD)
    ap.add_argument("--c2-curves", type=int, default=65536)
    a = ap.parse_args()
    if a.backend == "ref":
        if not os.path.exists(a.lib):
            raise SystemExit("no reference build at %s: run oracle/ref_recipe/build_ref.sh on a box with gfortran" % a.lib)
        backend = RefLib(a.lib)
    else:
        backend = OracleBackend()
    os.makedirs(a.out, exist_ok=True)
    np.savez_compressed(os.path.join(a.out, "ref_curfit.npz"), **run_curfit_cases(backend))
    np.savez_compressed(os.path.join(a.out, "ref_c2.npz"), **run_c2(backend, a.c2_curves))
    print("wrote ref_curfit.npz, ref_c2.npz to", a.out, "backend:", a.backend)


if __name__ == "__main__":
    main()
