"""CPU: the oracle against the committed golden vectors (second source: scipy's C
translation of netlib FITPACK, tests/golden/make_golden.py) and against every
assertion the reference's own tests make for this path."""
import os

import numpy as np
import pytest

import oracle_lib as O
from refdata import FPKNOT_X, FPKNOT_Y, MNCURF_STEPS, MNCURF_X, MNCURF_Y

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_equal_semantics():
    # core:18644-18647 : exact equality for normal numbers, |x|<tiny against zero
    tiny = np.finfo(float).tiny
    assert O.equal(1.0, 1.0)
    assert not O.equal(1.0, np.nextafter(1.0, 2.0))
    assert not O.equal(1.0, np.nextafter(1.0, 0.0))
    assert O.equal(0.0, 0.0)
    assert O.equal(tiny / 2, 0.0)
    assert not O.equal(tiny, 0.0)
    assert O.equal(-tiny / 4, 0.0)
    rng = np.random.default_rng(1)
    for _ in range(2000):
        a = rng.standard_normal() * 10.0 ** rng.integers(-300, 300)
        b = np.nextafter(a, np.inf) if rng.random() < 0.5 else a
        assert O.equal(a, b) == (abs(a - b) < tiny)


def test_knot_interval_equals_linear_walk():
    # core:2432-2473 : hybrid search == plain linear walk for non-decreasing knots
    rng = np.random.default_rng(2)
    for _ in range(300):
        n = int(rng.integers(8, 60))
        t = np.sort(rng.uniform(0, 1, n))
        if rng.random() < 0.3:
            t[3] = t[4]
        l0 = int(rng.integers(1, n // 2))
        lmax = int(rng.integers(l0, n))
        x = rng.uniform(-0.1, 1.1)
        l = l0
        while l < lmax and x >= t[l]:
            l += 1
        assert O.knot_interval(t, x, l0, lmax) == l


def test_curfit_golden_bit_exact():
    g = np.load(os.path.join(G, "curfit_golden.npz"))
    ncase = g["k"].size
    assert ncase > 200
    for i in range(ncase):
        k, m, nest = int(g["k"][i]), int(g["m"][i]), int(g["nest"][i])
        o = int(g["off"][i])
        x, y, w = g["x"][o:o + m], g["y"][o:o + m], g["w"][o:o + m]
        r = O.curfit(0, x, y, w, float(g["xb"][i]), float(g["xe"][i]), k, float(g["s"][i]), nest)
        assert r["ier"] == int(g["ier"][i]), i
        if r["ier"] == 10:
            continue
        n = int(g["n"][i])
        assert r["n"] == n, i
        to = int(g["toff"][i])
        assert np.array_equal(r["t"][:n], g["t"][to:to + n]), i
        assert np.array_equal(r["c"][:n - k - 1], g["c"][to:to + n - k - 1]), i
        assert r["fp"] == float(g["fp"][i]), i


def test_mncurf_golden_and_reference_assertions():
    g = np.load(os.path.join(G, "mncurf_golden.npz"))
    x, y, w = MNCURF_X, MNCURF_Y, np.ones(25)
    for k in (3, 5):
        # (a) the reference's own sequence (test/fitpack_tests.f90:1173-1251): all ier<=0,
        #     state carried through wrk/iwrk for iopt=1
        nest, lwrk = 35, 1000
        t = np.zeros(nest); c = np.zeros(nest); wrk = np.zeros(lwrk); iwrk = np.zeros(nest, np.int32)
        n = 0
        seq = {}
        for step, (iopt, s) in enumerate(MNCURF_STEPS, 1):
            if iopt == -1:
                t[k + 1:k + 8] = 3.0 * np.arange(1, 8)
                n = 9 + 2 * k
            r = O.curfit(iopt, x, y, w, 0.0, 24.0, k, s, nest, n=n, t=t, c=c, wrk=wrk, iwrk=iwrk,
                         lwrk=lwrk)
            n = r["n"]
            assert O.lib().orc_success(r["ier"]), (k, step)
            sp, ier = O.splev(t, n, c, k, x, e=0)
            assert ier == 0
            seq[step] = (n, t[:n].copy(), c[:n].copy(), r["fp"], r["ier"], sp.copy())
        # (b) steps scipy can replay exactly
        for step, key in ((1, "s1000"), (5, "s30"), (6, "s0"), (7, "lsq")):
            n, tt, cc, fp, ier, sp = seq[step]
            assert n == int(g[f"k{k}_{key}_n"]) and ier == int(g[f"k{k}_{key}_ier"])
            assert np.array_equal(tt, g[f"k{k}_{key}_t"])
            assert np.array_equal(cc[:n - k - 1], g[f"k{k}_{key}_c"][:n - k - 1])
            assert fp == float(g[f"k{k}_{key}_fp"])
            if key != "lsq":
                assert np.array_equal(sp, g[f"k{k}_{key}_sp"])
        # step 2 (iopt=1 right after the polynomial) equals a fresh s=60 fit
        n, tt, cc, fp, ier, sp = seq[2]
        assert n == int(g[f"k{k}_s60_n"]) and fp == float(g[f"k{k}_s60_fp"])
        # (c) iopt=1 continuation on unchanged knots
        r = O.curfit(0, x, y, w, 0.0, 24.0, k, 10.0, nest)
        assert r["n"] == int(g[f"k{k}_s10_n"]) and r["fp"] == float(g[f"k{k}_s10_fp"])
        r2 = O.curfit(1, x, y, w, 0.0, 24.0, k, 30.0, nest, n=r["n"], t=r["t"], wrk=r["wrk"],
                      iwrk=r["iwrk"])
        n = r2["n"]
        assert n == int(g[f"k{k}_cont30_n"]) and r2["ier"] == int(g[f"k{k}_cont30_ier"])
        assert r2["fp"] == float(g[f"k{k}_cont30_fp"])
        assert np.array_equal(r2["c"][:n - k - 1], g[f"k{k}_cont30_c"][:n - k - 1])
        # interpolating spline reproduces the data (ier=-1, fp=0)
        n, tt, cc, fp, ier, sp = seq[6]
        assert ier == -1 and fp == 0.0 and n == 25 + k + 1
        assert np.allclose(sp, y, rtol=0, atol=1e-10)


def test_splev_golden_bit_exact():
    g = np.load(os.path.join(G, "splev_golden.npz"))
    xs = g["mnspev_x"]
    for k in range(1, 6):
        t, c = g[f"mnspev_k{k}_t"], g[f"mnspev_k{k}_c"]
        y, ier = O.splev(t, t.size, c, k, xs, e=0)
        assert ier == 0
        assert np.array_equal(y, g[f"mnspev_k{k}_y"])
    for i in range(int(g["nrandom"])):
        t, c, x, k = g[f"r{i}_t"], g[f"r{i}_c"], g[f"r{i}_x"], int(g[f"r{i}_k"])
        for e in (0, 1, 3):
            y, ier = O.splev(t, t.size, c, k, x, e=e)
            assert ier == 0
            assert np.array_equal(y, g[f"r{i}_y_e{e}"]), (i, e)
        # e=2: reference returns FITPACK_INVALID_RANGE (6), core:17872 (netlib returns 1)
        y, ier = O.splev(t, t.size, c, k, x, e=2)
        assert ier == 6
        y, ier = O.splev(t, t.size, c, k, np.clip(x, 0, 1), e=2)
        assert ier == 0


def test_splev_interval_is_history_independent():
    # core:17880-17887: result == rightmost l in [k1,nk1] with t(l)<=arg
    rng = np.random.default_rng(5)
    for k in (1, 3, 5):
        t = np.concatenate([np.zeros(k + 1), np.sort(rng.uniform(0, 1, 17)), np.ones(k + 1)])
        n = t.size
        c = rng.standard_normal(n)
        x = rng.uniform(-0.1, 1.1, 400)
        y1, _, l1 = O.splev(t, n, c, k, x, want_idx=True)
        p = rng.permutation(400)
        y2, _, l2 = O.splev(t, n, c, k, x[p], want_idx=True)
        assert np.array_equal(y1[p], y2) and np.array_equal(l1[p], l2)
        k1, nk1 = k + 1, n - k - 1
        for xi, li in zip(x, l1):
            cand = [l for l in range(k1, nk1 + 1) if t[l - 1] <= xi]
            assert li == (max(cand) if cand else k1)


def test_m_lt_1_and_input_errors():
    t = np.array([0., 0, 0, 0, 1, 1, 1, 1]); c = np.zeros(8)
    _, ier = O.splev(t, 8, c, 3, np.zeros(0))
    assert ier == 10
    x = np.linspace(0, 1, 10); y = x ** 2
    assert O.curfit(0, x, y, None, 0, 1, 0, 1.0, 20)["ier"] == 10      # k out of range
    assert O.curfit(0, x, y, None, 0, 1, 6, 1.0, 20)["ier"] == 10
    assert O.curfit(2, x, y, None, 0, 1, 3, 1.0, 20)["ier"] == 10      # iopt
    assert O.curfit(0, x[:3], y[:3], None, 0, 1, 3, 1.0, 20)["ier"] == 10  # m<k1
    assert O.curfit(0, x, y, None, 0, 1, 3, 1.0, 7)["ier"] == 10       # nest<2k+2
    assert O.curfit(0, x, y, None, 0.1, 1, 3, 1.0, 20)["ier"] == 10    # xb>x(1)
    assert O.curfit(0, x, y, None, 0, 0.9, 3, 1.0, 20)["ier"] == 10    # xe<x(m)
    assert O.curfit(0, x[::-1], y, None, 0, 1, 3, 1.0, 20)["ier"] == 10  # decreasing x
    assert O.curfit(0, x, y, None, 0, 1, 3, -1.0, 20)["ier"] == 10     # s<0
    assert O.curfit(0, x, y, None, 0, 1, 3, 0.0, 13)["ier"] == 10      # s=0 needs nest>=m+k1
    assert O.curfit(0, x, y, None, 0, 1, 3, 1.0, 20, lwrk=10)["ier"] == 10  # lwrk too small
    r = O.curfit(0, x, y, None, 0, 1, 3, 0.0, 14)
    assert r["ier"] == -1 and r["n"] == 14


def test_fpknot_crash_regression():
    # test/fitpack_curve_tests.f90:61-85: new_fit(x,y,order=1) with the default smoothing
    # trajectory 1000 -> 60 -> 30 (core:351-370), iopt 0 -> 1 -> 1, nest = max(102, ceil(1.4 m))
    m, k = 109, 1
    nest = max(102, int(np.ceil(np.float32(1.4) * m)))
    lwrk = m * 6 + nest * 33
    t = np.zeros(nest); c = np.zeros(nest); wrk = np.zeros(lwrk); iwrk = np.zeros(nest, np.int32)
    n, iopt = 0, 0
    for s in (1000.0, 60.0, 30.0):
        r = O.curfit(iopt, FPKNOT_X, FPKNOT_Y, None, 0.0, 108.0, k, s, nest, n=n, t=t, c=c,
                     wrk=wrk, iwrk=iwrk, lwrk=lwrk)
        n = r["n"]
        assert r["ier"] <= 0
        iopt = 1
    assert abs(r["fp"] - 30.0) < 30.0 * 1e-3


def test_sine_interpolation_analytic():
    # test/fitpack_curve_tests.f90:418-505 (value part) and test/test_curve.cpp:31-93
    N = 201
    x = np.linspace(0, 2 * np.pi, N)
    y = np.sin(x)
    nest = max(102, int(np.ceil(np.float32(1.4) * N)))
    r = O.curfit(0, x, y, None, x[0], x[-1], 3, 0.0, nest)
    assert r["ier"] == -1 and r["n"] == N + 4
    xm = 0.5 * (x[:-1] + x[1:])
    sp, ier = O.splev(r["t"], r["n"], r["c"], 3, xm, e=3)
    assert ier == 0
    assert np.max(np.abs(sp - np.sin(xm))) < 1e-7


def test_self_consistency_properties():
    # style of test/fitpack_grid_nd_tests.f90 gates: fp == recomputed SSR, knots valid,
    # ier=0 => |fp-s| < 1e-3 s
    from refdata import synth_curves
    rng = np.random.default_rng(11)
    x, y = synth_curves(rng, 40, 256)
    for b in range(40):
        k = 3 if b % 2 == 0 else 5
        r = O.curfit(0, x[b], y[b], None, 0.0, 1.0, k, 2.56, 359)
        n = r["n"]
        assert r["ier"] in (0, -2)
        if r["ier"] == 0:
            assert abs(r["fp"] - 2.56) < 1e-3 * 2.56
        else:
            assert n == 2 * k + 2 and r["fp"] < 2.56 * (1 + 1e-3)
        sp, _ = O.splev(r["t"], n, r["c"], k, x[b])
        ssr = float(np.sum((y[b] - sp) ** 2))
        assert abs(ssr - r["fp"]) <= 1e-11 * max(1.0, ssr)
        assert O.fpchec(x[b], r["t"][:n], n, k) == 0


def test_argsort_matches_numpy_on_distinct_values():
    rng = np.random.default_rng(3)
    for n in (1, 2, 5, 8, 9, 33, 257):
        v = rng.permutation(n).astype(float) * 0.37
        assert np.array_equal(O.argsort(v) - 1, np.argsort(v))
    v = np.arange(50.0)
    assert np.array_equal(O.argsort(v), np.arange(1, 51))


def test_batch_driver_equals_single_calls():
    from refdata import synth_curves
    rng = np.random.default_rng(12)
    x, y = synth_curves(rng, 64, 128)
    out = O.curfit_batch(0, x, y, None, 3, 1.28, 190, want_stats=True, nthreads=4)
    for b in range(64):
        r = O.curfit(0, x[b], y[b], None, 0.0, 1.0, 3, 1.28, 190)
        assert out["n"][b] == r["n"] and out["ier"][b] == r["ier"] and out["fp"][b] == r["fp"]
        assert np.array_equal(out["t"][b, :r["n"]], r["t"][:r["n"]])
        assert np.array_equal(out["c"][b, :r["n"] - 4], r["c"][:r["n"] - 4])
        assert tuple(out["stats"][b]) == r["stats"]


# ------------------------------------------------------------------ gridded surfaces (dims = 2)
from refdata import DAREGR_X, DAREGR_Y, DAREGR_Z, MNREGR_STEPS  # noqa: E402


def _regrid_cases():
    g = np.load(os.path.join(G, "regrid_golden.npz"))
    for i in range(int(g["ncase"])):
        p = f"g{i}_"
        kx, ky, nxest, nyest, nx, ny, ier = (int(v) for v in g[p + "par"])
        yield dict(i=i, x=g[p + "x"], y=g[p + "y"], z=g[p + "z"], kx=kx, ky=ky, nxest=nxest,
                   nyest=nyest, nx=nx, ny=ny, ier=ier, s=float(g[p + "s"]), fp=float(g[p + "fp"]),
                   tx=g[p + "tx"], ty=g[p + "ty"], c=g[p + "c"],
                   ex=g[p + "ex"] if p + "ex" in g.files else None,
                   ey=g[p + "ey"] if p + "ey" in g.files else None,
                   ez=g[p + "ez"] if p + "ez" in g.files else None)


def test_regrid_and_bispev_golden_bit_exact_in_netlib_mode():
    """scipy's C translation of netlib regrid/bispev is reproduced bit for bit (n, knots, c, fp,
    ier, values) with the documented reference-vs-netlib deviations switched off."""
    O.set_netlib_compat(True)
    try:
        ncase = ncapped = 0
        for q in _regrid_cases():
            x, y = q["x"], q["y"]
            r = O.regrid(0, x, y, q["z"], x[0], x[-1], y[0], y[-1], q["kx"], q["ky"], q["s"],
                         q["nxest"], q["nyest"])
            if q["ier"] != -1 and (q["nx"] == x.size + q["kx"] + 1 or q["ny"] == y.size + q["ky"] + 1):
                # An axis was driven to the interpolation cap m+k+1 by knot insertion: scipy's C
                # translation then rewrites the OTHER axis' knots (observed: one knot inserted
                # without a count update), which neither netlib's fpregr nor the reference
                # (core:12254-12264, plain `exit add_knots`) does.  Not comparable; 6 of 49 cases.
                ncapped += 1
                continue
            assert r["ier"] == q["ier"], q["i"]
            if q["ier"] == 10:
                continue
            assert (r["nx"], r["ny"]) == (q["nx"], q["ny"]), q["i"]
            assert np.array_equal(r["tx"][:q["nx"]], q["tx"]) and np.array_equal(r["ty"][:q["ny"]], q["ty"])
            assert np.array_equal(r["c"][:q["c"].size], q["c"]), q["i"]
            assert r["fp"] == q["fp"], q["i"]
            if q["ez"] is not None:
                zz, ie = O.bispev(q["tx"], q["nx"], q["ty"], q["ny"], q["c"], q["kx"], q["ky"],
                                  q["ex"], q["ey"])
                assert ie == 0 and np.array_equal(zz.ravel(), q["ez"]), q["i"]
            ncase += 1
        assert ncase >= 40 and ncapped <= 6
    finally:
        O.set_netlib_compat(False)


def _grid_residual(r, kx, ky, x, y, z):
    zz, ie = O.bispev(r["tx"], r["nx"], r["ty"], r["ny"], r["c"], kx, ky, x, y)
    assert ie == 0
    return float(((z.reshape(x.size, y.size) - zz) ** 2).sum())


def test_regrid_reference_mode_properties():
    """Default (reference) mode: same code with the reference's knot-direction arbiter and sum
    association.  Self-consistency gates in the style of test/fitpack_grid_nd_tests.f90:1-21."""
    for q in _regrid_cases():
        x, y = q["x"], q["y"]
        r = O.regrid(0, x, y, q["z"], x[0], x[-1], y[0], y[-1], q["kx"], q["ky"], q["s"],
                     q["nxest"], q["nyest"])
        if q["ier"] == 10:
            assert r["ier"] == 10
            continue
        assert r["ier"] in (-2, -1, 0, 1, 2, 3)
        res = _grid_residual(r, q["kx"], q["ky"], x, y, q["z"])
        assert abs(res - r["fp"]) <= 1e-9 * max(1.0, res), (q["i"], res, r["fp"])
        if r["ier"] == 0:
            assert abs(r["fp"] - q["s"]) < 1e-3 * q["s"]
        if r["ier"] == -1:
            assert (r["nx"], r["ny"]) == (x.size + q["kx"] + 1, y.size + q["ky"] + 1) and res < 1e-18
        if r["ier"] == -2:
            assert (r["nx"], r["ny"]) == (2 * q["kx"] + 2, 2 * q["ky"] + 2)
        for (t, n, k, xx) in ((r["tx"], r["nx"], q["kx"], x), (r["ty"], r["ny"], q["ky"], y)):
            assert O.fpchec(xx, t[:n], n, k) == 0   # Schoenberg-Whitney holds for the chosen knots


def test_mnregr_sequence_reference_assertions():
    """reference test mnregr (test/fitpack_tests.f90:3170-3362): six regrid calls incl. iopt=1
    continuations and prescribed knots, each followed by bispev; all must succeed."""
    x, y, z = DAREGR_X, DAREGR_Y, DAREGR_Z
    nxest = nyest = 17
    state = dict(tx=np.zeros(nxest), ty=np.zeros(nyest), wrk=np.zeros(2000),
                 iwrk=np.zeros(200, np.int32), nx=0, ny=0)
    fps = []
    for (iopt, kx, ky, s) in MNREGR_STEPS:
        if iopt == -1:
            state["nx"] = state["ny"] = 11
            state["tx"][:] = 0.0
            state["ty"][:] = 0.0
            state["tx"][kx + 1:kx + 4] = [-0.5, 0.0, 0.5]
            state["ty"][ky + 1:ky + 4] = [-0.5, 0.0, 0.5]
        r = O.regrid(iopt, x, y, z, x[0], x[-1], y[0], y[-1], kx, ky, s, nxest, nyest,
                     nx=state["nx"], ny=state["ny"], tx=state["tx"], ty=state["ty"],
                     wrk=state["wrk"], iwrk=state["iwrk"], lwrk=2000, kwrk=200)
        assert r["ier"] <= 0, (iopt, kx, ky, s, r["ier"])
        state.update(nx=r["nx"], ny=r["ny"], tx=r["tx"], ty=r["ty"])
        zz, ie = O.bispev(r["tx"], r["nx"], r["ty"], r["ny"], r["c"], kx, ky, x, y)
        assert ie == 0
        res = float(((z.reshape(11, 11) - zz) ** 2).sum())
        assert abs(res - r["fp"]) < 1e-10
        fps.append(r["fp"])
    assert fps[3] == 0.0 and abs(fps[1] - 0.22) < 0.22e-3     # interpolation; smoothing hit s
    assert (state["nx"], state["ny"]) == (11, 11)


def test_regrid_and_bispev_input_errors():
    x, y, z = DAREGR_X, DAREGR_Y, DAREGR_Z
    ok = dict(iopt=0, x=x, y=y, z=z, xb=x[0], xe=x[-1], yb=y[0], ye=y[-1], kx=3, ky=3, s=0.1,
              nxest=15, nyest=15)
    assert O.regrid(**ok)["ier"] <= 0
    for bad in (dict(kx=0), dict(ky=6), dict(iopt=2), dict(nxest=7), dict(s=-1.0), dict(xb=-1.0),
                dict(ye=1.0), dict(s=0.0, nyest=14), dict(x=x[::-1].copy())):
        assert O.regrid(**{**ok, **bad})["ier"] == 10, bad
    assert O.regrid(**ok, lwrk=10)["ier"] == 10
    r = O.regrid(**ok)
    zz, ie = O.bispev(r["tx"], r["nx"], r["ty"], r["ny"], r["c"], 3, 3, x[::-1].copy(), y)
    assert ie == 10


# ------------------------------------------------------------------ spline calculus
def _calculus_cases():
    g = np.load(os.path.join(G, "calculus_golden.npz"))
    for i in range(int(g["nspline"])):
        yield g, f"s{i}_"


def test_splder_spalde_splint_golden_bit_exact():
    nchk = 0
    for g, p in _calculus_cases():
        t, c, x, k = g[p + "t"], g[p + "c"], g[p + "x"], int(g[p + "k"])
        n = t.size
        for nu in range(k + 1):
            for e in (0, 1):
                y, ier, wrk = O.splder(t, n, c, k, nu, x, e)
                assert ier == 0 and np.array_equal(y, g[p + f"d{nu}_e{e}"]), (p, nu, e)
                nchk += 1
        for xx, dref in zip(g[p + "ax"], g[p + "ad"]):
            d, ier = O.spalde(t, n, c, k + 1, float(xx))
            assert ier == 0 and np.array_equal(d, dref)
        for (a, b), vref in zip(g[p + "ab"], g[p + "iv"]):
            v, _ = O.splint(t, n, c, k, float(a), float(b))
            assert v == vref
    assert nchk >= 150


def test_calculus_semantics():
    t = np.array([0., 0, 0, 0, 0.5, 1, 1, 1, 1])
    c = np.array([1., 2, -1, 3, 0.5, 0, 0, 0, 0])
    x = np.array([0.25, 1.5, 0.75])
    # e=2: ier=1 at the first outside point, y filled up to it (core:17771-17773)
    y, ier, _ = O.splder(t, 9, c, 3, 1, x, 2)
    assert ier == 1
    assert y[0] == O.splder(t, 9, c, 3, 1, x[:1], 0)[0][0]
    # e=3 is not a splder policy: extrapolates like 0
    assert np.array_equal(O.splder(t, 9, c, 3, 2, x, 3)[0], O.splder(t, 9, c, 3, 2, x, 0)[0])
    assert O.splder(t, 9, c, 3, 4, x, 0)[1] == 10 and O.splder(t, 9, c, 3, -1, x, 0)[1] == 10
    # nu = 0 is splev
    assert np.array_equal(O.splder(t, 9, c, 3, 0, x, 0)[0], O.splev(t, 9, c, 3, x, 0)[0])
    # spalde outside the support -> 10; d[0] is the spline value
    assert O.spalde(t, 9, c, 4, 1.5)[1] == 10
    d, ier = O.spalde(t, 9, c, 4, 0.25)
    assert ier == 0 and abs(d[0] - O.splev(t, 9, c, 3, np.array([0.25]))[0][0]) < 1e-15
    # integral is antisymmetric in its limits and additive
    a1, _ = O.splint(t, 9, c, 3, 0.1, 0.9)
    a2, _ = O.splint(t, 9, c, 3, 0.9, 0.1)
    a3 = O.splint(t, 9, c, 3, 0.1, 0.4)[0] + O.splint(t, 9, c, 3, 0.4, 0.9)[0]
    assert a1 == -a2 and abs(a1 - a3) < 1e-14


# ------------------------------------------------------------------ parametric curves
def _parcur_cases():
    g = np.load(os.path.join(G, "parcur_golden.npz"))
    for i in range(g["k"].size):
        k, m, idim = int(g["k"][i]), int(g["m"][i]), int(g["idim"][i])
        xo, uo = int(g["xoff"][i]), int(g["uoff"][i])
        yield i, g, dict(k=k, m=m, idim=idim, ipar=int(g["ipar"][i]), iopt=int(g["iopt"][i]),
                         s=float(g["s"][i]), nest=int(g["nest"][i]), ub=float(g["ub"][i]),
                         ue=float(g["ue"][i]), s0=float(g["s0"][i]),
                         x=g["x"][xo:xo + m * idim].reshape(m, idim), w=g["w"][uo:uo + m],
                         u_in=g["u_in"][uo:uo + m], u=g["u"][uo:uo + m])


def run_parcur_case(fn, cs, g, i):
    """Replay golden case i through a parcur implementation `fn` with oracle_lib.parcur's signature."""
    k, idim, nest = cs["k"], cs["idim"], cs["nest"]
    if cs["iopt"] == 0:
        return fn(0, cs["ipar"], idim, cs["x"], cs["w"], k, cs["s"], nest, u=cs["u_in"], ub=cs["ub"],
                  ue=cs["ue"])
    if cs["iopt"] == -1:
        n = int(g["n"][i]) if int(g["ier"][i]) != 10 else nest
        to = int(g["toff"][i])
        t = np.zeros(nest)
        if int(g["ier"][i]) != 10:
            t[:n] = g["t"][to:to + n]
        return fn(-1, cs["ipar"], idim, cs["x"], cs["w"], k, 0.0, nest, u=cs["u_in"], ub=cs["ub"],
                  ue=cs["ue"], n=n, t=t)
    r0 = fn(0, cs["ipar"], idim, cs["x"], cs["w"], k, cs["s0"], nest, u=cs["u_in"], ub=cs["ub"], ue=cs["ue"])
    return fn(1, cs["ipar"], idim, cs["x"], cs["w"], k, cs["s"], nest, u=r0["u"], ub=r0["ub"], ue=r0["ue"],
              n=r0["n"], t=r0["t"], wrk=r0["wrk"], iwrk=r0["iwrk"])


def check_parcur_case(r, cs, g, i, exact=True, rtol=0.0):
    k, idim = cs["k"], cs["idim"]
    assert r["ier"] == int(g["ier"][i]), i
    if r["ier"] == 10:
        return
    n = int(g["n"][i])
    assert r["n"] == n, i
    nk1 = n - k - 1
    to, co = int(g["toff"][i]), int(g["coff"][i])
    assert np.array_equal(r["t"][:n], g["t"][to:to + n]), i
    assert np.array_equal(r["u"], cs["u"]), i
    cc = np.concatenate([r["c"][d * n:d * n + nk1] for d in range(idim)])
    gc = g["c"][co:co + nk1 * idim]
    if exact:
        assert np.array_equal(cc, gc), i
        assert r["fp"] == float(g["fp"][i]), i
    else:
        assert np.allclose(cc, gc, rtol=rtol, atol=rtol * np.abs(gc).max()), i
        assert abs(r["fp"] - float(g["fp"][i])) <= rtol * max(abs(float(g["fp"][i])), 1e-300), i


def test_parcur_and_curev_golden_bit_exact_in_netlib_mode():
    """scipy's C translation of netlib parcur/fppara: the oracle with the documented reference
    deviations switched to the netlib form (initial p = nk1/sum(a) instead of the reference's
    sum(a)/nk1 core:9119, fp accumulated coordinate by coordinate instead of fp+sum(xi**2)
    core:8972, chord lengths by sqrt(sum d^2) instead of NORM2 core:15243, fpknot seeding, loop
    bound) must reproduce n, knots, u, c, fp, ier bit for bit; curev == splev(e=3) per coordinate."""
    O.set_netlib_compat(1)
    try:
        ncase = 0
        for i, g, cs in _parcur_cases():
            r = run_parcur_case(O.parcur, cs, g, i)
            check_parcur_case(r, cs, g, i)
            ncase += 1
            if r["ier"] != 10:
                eo = int(g["evoff"][i])
                pts = g["ev"][eo:eo + 23]
                want = g["ev"][eo + 23:eo + 23 * (1 + cs["idim"])].reshape(23, cs["idim"])
                got, ier = O.curev(cs["idim"], r["t"], r["n"], r["c"], cs["k"], pts)
                assert ier == 0 and np.array_equal(got, want), i
        assert ncase > 140
    finally:
        O.set_netlib_compat(0)


def test_parcur_reference_mode_properties():
    """Default (reference) mode: same discrete outcome as netlib on regular data, fp equals the
    residual recomputed through curev, ier=0 => |fp-s| < 1e-3 s, knots pass fpchec; curev rejects
    decreasing parameters (core:1145) and clamps outside [tb,te]."""
    nchk = 0
    for i, g, cs in _parcur_cases():
        if cs["iopt"] != 0 or int(g["ier"][i]) == 10:
            continue
        r = run_parcur_case(O.parcur, cs, g, i)
        assert r["ier"] == int(g["ier"][i]) and r["n"] == int(g["n"][i]), i
        ev, ier = O.curev(cs["idim"], r["t"], r["n"], r["c"], cs["k"], r["u"])
        assert ier == 0
        res = ((cs["w"][:, None] * (ev - cs["x"])) ** 2).sum()
        assert abs(res - r["fp"]) <= 1e-9 * max(1.0, res, cs["s"]), i
        if r["ier"] == 0:
            assert abs(r["fp"] - cs["s"]) < 1e-3 * cs["s"], i
        assert O.fpchec(r["u"], r["t"], r["n"], cs["k"]) == 0, i
        nchk += 1
    assert nchk > 60
    t = np.array([0., 0, 0, 0, 1, 1, 1, 1])
    c = np.arange(16.0)
    _, ier = O.curev(2, t, 8, c, 3, np.array([0.5, 0.4]))
    assert ier == 10
    v, ier = O.curev(2, t, 8, c, 3, np.array([-1.0, 2.0]))
    assert ier == 0 and np.array_equal(v, [[0., 8.], [3., 11.]])


# ------------------------------------------------------------------ periodic curves
def _percur_cases():
    g = np.load(os.path.join(G, "percur_golden.npz"))
    for i in range(g["k"].size):
        m, o = int(g["m"][i]), int(g["off"][i])
        yield i, g, dict(k=int(g["k"][i]), m=m, iopt=int(g["iopt"][i]), s=float(g["s"][i]),
                         s0=float(g["s0"][i]), nest=int(g["nest"][i]), x=g["x"][o:o + m],
                         y=g["y"][o:o + m], w=g["w"][o:o + m])


def run_percur_case(fn, cs, g, i):
    """Replay golden case i through a percur implementation `fn` with oracle_lib.percur's signature."""
    k, nest = cs["k"], cs["nest"]
    if cs["iopt"] == 0:
        return fn(0, cs["x"], cs["y"], cs["w"], k, cs["s"], nest)
    if cs["iopt"] == -1:
        n = int(g["n"][i]) if int(g["ier"][i]) != 10 else nest
        t = np.zeros(nest)
        if int(g["ier"][i]) != 10:
            to = int(g["toff"][i])
            t[:n] = g["t"][to:to + n]
        return fn(-1, cs["x"], cs["y"], cs["w"], k, 0.0, nest, n=n, t=t)
    r0 = fn(0, cs["x"], cs["y"], cs["w"], k, cs["s0"], nest)
    n0 = r0["n"]
    return fn(1, cs["x"], cs["y"], cs["w"], k, cs["s"], n0, n=n0, t=r0["t"][:n0].copy(), wrk=r0["wrk"],
              iwrk=r0["iwrk"][:n0].copy())


def check_percur_case(r, cs, g, i):
    assert r["ier"] == int(g["ier"][i]), i
    if r["ier"] == 10:
        return
    n, k = int(g["n"][i]), cs["k"]
    assert r["n"] == n, i
    to = int(g["toff"][i])
    assert np.array_equal(r["t"][:n], g["t"][to:to + n]), i
    assert np.array_equal(r["c"][:n - k - 1], g["c"][to:to + n - k - 1]), i
    assert r["fp"] == float(g["fp"][i]), i


def test_percur_golden_bit_exact_in_netlib_mode():
    """scipy's C translation of netlib percur/fpperi/fpchep: n, knots, c, fp, ier bit for bit (the
    oracle's netlib switch only changes fpknot seeding, the zero-pivot tests and the association of
    the initial-p sum core:10081)."""
    O.set_netlib_compat(1)
    try:
        ncase = 0
        for i, g, cs in _percur_cases():
            check_percur_case(run_percur_case(O.percur, cs, g, i), cs, g, i)
            ncase += 1
        assert ncase > 190
    finally:
        O.set_netlib_compat(0)


def test_percur_reference_mode_properties():
    """Default (reference) mode: same discrete outcome, periodicity c(n7+j) = c(j), value and k-1
    derivatives continuous across the period, fp equals the recomputed residual, ier=0 => |fp-s|<1e-3 s,
    chosen knots pass fpchep."""
    nchk = 0
    for i, g, cs in _percur_cases():
        if cs["iopt"] != 0 or int(g["ier"][i]) == 10:
            continue
        r = run_percur_case(O.percur, cs, g, i)
        assert r["ier"] == int(g["ier"][i]) and r["n"] == int(g["n"][i]), i
        n, k = r["n"], cs["k"]
        n7 = n - 2 * k - 1
        if n7 >= 1 and r["ier"] != -2:
            assert np.array_equal(r["c"][n7:n7 + k], r["c"][:k]), i
        m1 = cs["m"] - 1
        ev, ier = O.splev(r["t"], n, r["c"], k, cs["x"][:m1])
        res = ((cs["w"][:m1] * (ev - cs["y"][:m1])) ** 2).sum()
        assert abs(res - r["fp"]) <= 1e-9 * max(1.0, res, cs["s"]), i
        e0, _ = O.splev(r["t"], n, r["c"], k, cs["x"][[0, -1]])
        assert abs(e0[0] - e0[1]) <= 1e-9 * max(1.0, abs(e0[0])), i
        if r["ier"] == 0:
            assert abs(r["fp"] - cs["s"]) < 1e-3 * cs["s"], i
        if n > 2 * k + 2:
            assert O.fpchep(cs["x"], r["t"], n, k) == 0, i
        nchk += 1
    assert nchk > 100
