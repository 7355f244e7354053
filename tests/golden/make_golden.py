"""Generate the golden vectors under tests/golden/ from the SECOND SOURCE.

The reference (perazz/fitpack) is Fortran and cannot be compiled in the build
container (no Fortran compiler), so its own binary cannot produce fixtures.
scipy.interpolate._fitpack is a C translation of the same Dierckx/netlib
algorithms; on inputs where the reference does not deliberately deviate from
netlib (fpknot seeding on all-zero/NaN residuals, zero-pivot
skip, subnormal pivots, main-loop trip count) the two must agree bit for bit.
This script records scipy's answers; tests/test_oracle_golden.py checks the
oracle against them, and the GPU tests check the CUDA path against both.

Run here (needs scipy):  python tests/golden/make_golden.py
Outputs: curfit_golden.npz, splev_golden.npz, mncurf_golden.npz
"""
import os

import numpy as np
from scipy.interpolate import _fitpack

HERE = os.path.dirname(os.path.abspath(__file__))

MNCURF_Y = np.array([1.0, 1.0, 1.4, 1.1, 1.0, 1.0, 4.0, 9.0, 13.0, 13.4, 12.8, 13.1, 13.0, 14.0,
                     13.0, 13.5, 10.0, 2.0, 3.0, 2.5, 2.5, 2.5, 3.0, 4.0, 3.5])


def lwrk_min(m, k, nest):
    return m * (k + 1) + nest * (7 + 3 * k)


def scipy_curfit(iopt, x, y, w, xb, xe, k, s, nest, t_in=None):
    m = x.size
    t = np.zeros(nest) if t_in is None else t_in.copy()
    n, t, c, fp, ier = _fitpack.curfit(iopt, x, y, w, xb, xe, k, s, t,
                                       np.zeros(lwrk_min(m, k, t.size)),
                                       np.zeros(t.size, np.int32))
    return int(n), np.asarray(t), np.asarray(c), float(fp), int(ier)


def synth_curves(rng, B, m, sigma=0.1):
    """C2-style synthetic batch: jittered abscissae, noisy sinusoid."""
    u = rng.uniform(-1, 1, (B, m))
    x = (np.arange(m)[None, :] + 0.3 * u) / (m - 1)
    x[:, 0] = 0.0
    x[:, -1] = 1.0
    A = rng.uniform(0.5, 2, (B, 1))
    f = rng.uniform(1, 4, (B, 1))
    ph = rng.uniform(0, 2 * np.pi, (B, 1))
    y = A * np.sin(2 * np.pi * f * x + ph) + sigma * rng.standard_normal((B, m))
    return x, y


def make_mncurf():
    """reference test mncurf (test/fitpack_tests.f90:1138-1269), the steps scipy can replay:
    iopt=0 and iopt=-1 calls, plus iopt=1 continuations that add no knots."""
    x = np.arange(25.0)
    w = np.ones(25)
    out = {}
    for k in (3, 5):
        for s in (1000.0, 60.0, 10.0, 30.0, 0.0):
            n, t, c, fp, ier = scipy_curfit(0, x, MNCURF_Y, w, 0.0, 24.0, k, s, 35)
            key = f"k{k}_s{int(s)}"
            out[key + "_n"] = n
            out[key + "_t"] = t[:n]
            out[key + "_c"] = c[:n]
            out[key + "_fp"] = fp
            out[key + "_ier"] = ier
            y, ier2 = _fitpack.splev(t[:n], c[:n], k, x, 0)
            out[key + "_sp"] = np.asarray(y)
        tin = np.zeros(9 + 2 * k)
        tin[k + 1:k + 8] = 3.0 * np.arange(1, 8)
        n, t, c, fp, ier = scipy_curfit(-1, x, MNCURF_Y, w, 0.0, 24.0, k, 0.0, tin.size, tin)
        key = f"k{k}_lsq"
        out[key + "_n"] = n
        out[key + "_t"] = t[:n]
        out[key + "_c"] = c[:n]
        out[key + "_fp"] = fp
        out[key + "_ier"] = ier
        # iopt=1 continuation without knot growth: s=10 fresh, then s=30 on the same knots
        t0 = np.zeros(35)
        wrk = np.zeros(lwrk_min(25, k, 35))
        iwrk = np.zeros(35, np.int32)
        n, t, c, fp, ier = _fitpack.curfit(0, x, MNCURF_Y, w, 0.0, 24.0, k, 10.0, t0, wrk, iwrk)
        tt = np.asarray(t)[:n].copy()
        n3, t3, c3, fp3, ier3 = _fitpack.curfit(1, x, MNCURF_Y, w, 0.0, 24.0, k, 30.0, tt, wrk, iwrk)
        key = f"k{k}_cont30"
        out[key + "_n"] = int(n3)
        out[key + "_t"] = np.asarray(t3)[:n3]
        out[key + "_c"] = np.asarray(c3)[:n3]
        out[key + "_fp"] = float(fp3)
        out[key + "_ier"] = int(ier3)
    out["x"] = x
    out["y"] = MNCURF_Y
    np.savez_compressed(os.path.join(HERE, "mncurf_golden.npz"), **out)


def make_curfit():
    rng = np.random.default_rng(20261019)
    cases = []
    # (a) random mixed cases, all degrees, several s regimes, weights, wider [xb,xe]
    for trial in range(160):
        k = int(rng.integers(1, 6))
        m = int(rng.integers(k + 2, 200))
        x = np.sort(rng.uniform(0, 1, m))
        if rng.random() < 0.3:
            x = np.linspace(0, 1, m)
        x = np.unique(x)
        m = x.size
        if m < k + 2:
            continue
        f = rng.uniform(1, 4)
        y = rng.uniform(.5, 2) * np.sin(2 * np.pi * f * x + rng.uniform(0, 6)) \
            + 0.1 * rng.standard_normal(m)
        w = np.ones(m) if rng.random() < 0.5 else rng.uniform(0.5, 2, m)
        s = float(rng.choice([0.0, m * 0.01, m * 0.001, m * 0.1, 1e-6, 1e-12, 1000.0]))
        nest = int(rng.choice([m + k + 1, max(2 * k + 2, m // 4), max(2 * k + 3, m // 2)]))
        xb, xe = x[0], x[-1]
        if rng.random() < 0.2:
            xb -= 0.1
            xe += 0.05
        cases.append((k, x, y, w, xb, xe, s, nest))
    # (b) C2-shaped: m=256, k=3, s=m*sigma^2, nest=359
    xs, ys = synth_curves(rng, 48, 256)
    for b in range(48):
        cases.append((3, xs[b], ys[b], np.ones(256), 0.0, 1.0, 2.56, 359))
    # (c) C5-shaped source fits: k=5
    xs, ys = synth_curves(rng, 16, 256)
    for b in range(16):
        cases.append((5, xs[b], ys[b], np.ones(256), 0.0, 1.0, 2.56, 359))

    rec = dict(k=[], m=[], s=[], nest=[], xb=[], xe=[], n=[], fp=[], ier=[], off=[], toff=[])
    xcat, ycat, wcat, tcat, ccat = [], [], [], [], []
    off = toff = 0
    for (k, x, y, w, xb, xe, s, nest) in cases:
        n, t, c, fp, ier = scipy_curfit(0, x, y, w, xb, xe, k, s, nest)
        if not np.isfinite(fp):
            continue
        if ier == 10:
            n = 0
        rec["k"].append(k); rec["m"].append(x.size); rec["s"].append(s); rec["nest"].append(nest)
        rec["xb"].append(xb); rec["xe"].append(xe); rec["n"].append(n); rec["fp"].append(fp)
        rec["ier"].append(ier); rec["off"].append(off); rec["toff"].append(toff)
        xcat.append(x); ycat.append(y); wcat.append(w)
        tcat.append(t[:n]); ccat.append(c[:n])
        off += x.size
        toff += n
    np.savez_compressed(
        os.path.join(HERE, "curfit_golden.npz"),
        k=np.array(rec["k"], np.int32), m=np.array(rec["m"], np.int32), s=np.array(rec["s"]),
        nest=np.array(rec["nest"], np.int32), xb=np.array(rec["xb"]), xe=np.array(rec["xe"]),
        n=np.array(rec["n"], np.int32), fp=np.array(rec["fp"]), ier=np.array(rec["ier"], np.int32),
        off=np.array(rec["off"], np.int64), toff=np.array(rec["toff"], np.int64),
        x=np.concatenate(xcat), y=np.concatenate(ycat), w=np.concatenate(wcat),
        t=np.concatenate(tcat), c=np.concatenate(ccat))
    print("curfit cases:", len(rec["k"]), "ier histogram:",
          {int(v): int((np.array(rec["ier"]) == v).sum()) for v in np.unique(rec["ier"])})


def make_splev():
    rng = np.random.default_rng(777)
    out = {}
    # mnspev (test/fitpack_tests.f90:3592-3667) with the netlib knot placement
    # (interior knots at t(k+2..k+5)); every entry of t defined.
    xs = 0.05 * np.arange(21)
    for k in range(1, 6):
        k1 = k + 1
        n = 2 * k1 + 4
        t = np.zeros(n)
        t[n - k1:] = 1.0
        t[k1:k1 + 4] = [0.1, 0.3, 0.4, 0.8]
        nk1 = n - k1
        c = np.zeros(n)
        c[:nk1] = [0.01 * i * (i - 5) for i in range(1, nk1 + 1)]
        y, ier = _fitpack.splev(t, c, k, xs, 0)
        out[f"mnspev_k{k}_t"] = t
        out[f"mnspev_k{k}_c"] = c
        out[f"mnspev_k{k}_y"] = np.asarray(y)
    out["mnspev_x"] = xs
    # random splines, all degrees, all out-of-support policies, unsorted x incl. outside points
    idx = 0
    for k in range(1, 6):
        for trial in range(6):
            nint = int(rng.integers(0, 30))
            interior = np.sort(rng.uniform(0, 1, nint))
            if trial == 5 and nint > 3:
                interior[2] = interior[1]  # a double interior knot
            t = np.concatenate([np.zeros(k + 1), interior, np.ones(k + 1)])
            n = t.size
            c = np.zeros(n)
            c[:n - k - 1] = rng.standard_normal(n - k - 1)
            x = rng.uniform(-0.2, 1.2, 257)
            x[:4] = [0.0, 1.0, -0.2, 1.2]
            if nint:
                x[4] = interior[0]
            for e in (0, 1, 3):
                y, ier = _fitpack.splev(t, c, k, x, e)
                out[f"r{idx}_y_e{e}"] = np.asarray(y)
            out[f"r{idx}_t"] = t
            out[f"r{idx}_c"] = c
            out[f"r{idx}_x"] = x
            out[f"r{idx}_k"] = k
            idx += 1
    out["nrandom"] = idx
    np.savez_compressed(os.path.join(HERE, "splev_golden.npz"), **out)


def scipy_regrid(x, y, z, kx, ky, s, nxest=None, nyest=None):
    """iopt=0 smoothing fit on a grid through scipy's C translation of netlib regrid."""
    mx, my = x.size, y.size
    nxest = mx + kx + 1 if nxest is None else nxest
    nyest = my + ky + 1 if nyest is None else nyest
    lwrk = (4 + mx + my + nxest * (my + 2 * kx + 5) + nyest * (2 * ky + 5) + mx * (kx + 1)
            + my * (ky + 1) + max((nxest - kx - 1) * (nyest - ky - 1), my))
    wrk = np.zeros(lwrk)
    iwrk = np.zeros(3 + mx + my + nxest + nyest, np.int32)
    nx, tx, ny, ty, c, fp, ier = _fitpack.regrid(0, x, y, np.ascontiguousarray(z).ravel(), x[0], x[-1],
                                                 y[0], y[-1], kx, ky, s, nxest, nyest, 20, wrk, iwrk)
    return int(nx), np.asarray(tx), int(ny), np.asarray(ty), np.asarray(c), float(fp), int(ier)


def grid_surface(rng, mx, my, sigma=0.05, uniform=True):
    """C4-style synthetic grid: smooth bivariate function + noise on [-1.5,1.5]^2"""
    if uniform:
        x = np.linspace(-1.5, 1.5, mx)
        y = np.linspace(-1.5, 1.5, my)
    else:
        x = np.sort(rng.uniform(-1.5, 1.5, mx)); x[0], x[-1] = -1.5, 1.5
        y = np.sort(rng.uniform(-1.5, 1.5, my)); y[0], y[-1] = -1.5, 1.5
        x, y = np.unique(x), np.unique(y)
    X, Y = np.meshgrid(x, y, indexing="ij")
    z = np.exp(-(X ** 2 + Y ** 2)) + 0.3 * np.sin(2.0 * X) * np.cos(1.5 * Y)
    return x, y, z + sigma * rng.standard_normal(z.shape)


def make_regrid():
    """regrid (iopt=0) and bispev vectors.  NOTE: the reference deliberately differs from netlib in
    the knot-direction arbiter (new_knot_dimension_nd, core:12347-12362: ties and the first knot go
    to the LAST axis; netlib alternates starting with x), in fpknot and in the association order of
    the tensor-product sums (fpgrre residual, fpndsp).  These vectors therefore pin the oracle in
    its netlib-compatibility mode (orc_set_netlib_compat), which shares every other line of code
    with the default (reference) mode."""
    from refdata import DAREGR_X, DAREGR_Y, DAREGR_Z  # noqa: F401 (tests/ on sys.path)
    rng = np.random.default_rng(4096)
    cases = []
    for (kx, ky, s) in [(3, 3, 10.0), (3, 3, 0.22), (3, 3, 0.1), (3, 3, 0.0), (5, 5, 0.2), (1, 2, 0.3),
                        (4, 3, 0.15), (2, 5, 0.5), (1, 1, 0.05)]:
        cases.append((DAREGR_X, DAREGR_Y, DAREGR_Z.reshape(11, 11), kx, ky, s, None, None))
    for trial in range(40):
        kx, ky = int(rng.integers(1, 6)), int(rng.integers(1, 6))
        mx, my = int(rng.integers(kx + 2, 48)), int(rng.integers(ky + 2, 48))
        x, y, z = grid_surface(rng, mx, my, 0.05, uniform=rng.random() < 0.5)
        mx, my = x.size, y.size
        if mx < kx + 2 or my < ky + 2:
            continue
        s = float(rng.choice([mx * my * 0.0025, mx * my * 0.001, mx * my * 0.01, 0.0, 1e3, 1e-9]))
        nxe = nye = None
        if rng.random() < 0.25:  # storage-limited
            nxe, nye = max(2 * kx + 2, mx // 2), max(2 * ky + 2, my // 2)
        cases.append((x, y, z, kx, ky, s, nxe, nye))
    out = {"ncase": 0}
    i = 0
    for (x, y, z, kx, ky, s, nxe, nye) in cases:
        nx, tx, ny, ty, c, fp, ier = scipy_regrid(x, y, z, kx, ky, s, nxe, nye)
        if not np.isfinite(fp):
            continue
        pre = f"g{i}_"
        out[pre + "x"], out[pre + "y"], out[pre + "z"] = x, y, np.ascontiguousarray(z).ravel()
        out[pre + "par"] = np.array([kx, ky, nxe if nxe else x.size + kx + 1,
                                     nye if nye else y.size + ky + 1, nx, ny, ier], np.int64)
        out[pre + "s"], out[pre + "fp"] = s, fp
        out[pre + "tx"], out[pre + "ty"] = tx[:nx], ty[:ny]
        ncof = (nx - kx - 1) * (ny - ky - 1)
        out[pre + "c"] = c[:ncof]
        if ier <= 0 or ier in (1, 2, 3):
            xe = np.sort(rng.uniform(x[0] - 0.1, x[-1] + 0.1, 37))
            ye = np.sort(rng.uniform(y[0] - 0.1, y[-1] + 0.1, 29))
            zz, ie = _fitpack.bispev(tx[:nx], ty[:ny], c[:ncof], kx, ky, xe, ye)
            out[pre + "ex"], out[pre + "ey"], out[pre + "ez"] = xe, ye, np.asarray(zz).ravel()
        i += 1
    out["ncase"] = i
    np.savez_compressed(os.path.join(HERE, "regrid_golden.npz"), **out)
    iers = [int(out[f"g{j}_par"][6]) for j in range(i)]
    print("regrid cases:", i, "ier histogram:", {v: iers.count(v) for v in sorted(set(iers))})


def make_calculus():
    """splder / spalde / splint vectors (scipy's C translation; no reference-vs-netlib deviation
    touches these routines)."""
    rng = np.random.default_rng(31337)
    out = {}
    idx = 0
    for k in range(1, 6):
        for trial in range(5):
            nint = int(rng.integers(0, 14))
            t = np.concatenate([np.zeros(k + 1), np.sort(rng.uniform(0, 1, nint)), np.ones(k + 1)])
            n = t.size
            c = np.zeros(n)
            c[:n - k - 1] = rng.standard_normal(n - k - 1)
            x = rng.uniform(-0.2, 1.2, 64)
            x[:3] = [0.0, 1.0, 0.5]
            p = f"s{idx}_"
            out[p + "t"], out[p + "c"], out[p + "x"], out[p + "k"] = t, c, x, k
            for nu in range(0, k + 1):
                for e in (0, 1):
                    y, ier = _fitpack.splder(t, c, k, nu, x, e)
                    out[p + f"d{nu}_e{e}"] = np.asarray(y)
            xs = np.array([0.0, 0.3, 1.0, 0.77])
            out[p + "ax"] = xs
            out[p + "ad"] = np.stack([np.ravel(_fitpack.spalde(t, c, k + 1, np.array([v]))[0]) for v in xs])
            ab = np.array([[0, 1], [0.2, 0.7], [0.9, 0.1], [-1, 2], [0.3, 0.3]], float)
            out[p + "ab"] = ab
            vals = []
            for a, b in ab:
                v = _fitpack.splint(t, c, k, a, b)
                vals.append(float(v[0] if isinstance(v, tuple) else v))
            out[p + "iv"] = np.array(vals)
            idx += 1
    out["nspline"] = idx
    np.savez_compressed(os.path.join(HERE, "calculus_golden.npz"), **out)



# ------------------------------------------------------------------ parametric curves
def parcur_lwrk(m, k, nest, idim):
    return m * (k + 1) + nest * (6 + idim + 3 * k)


def scipy_parcur(iopt, ipar, idim, u, x, w, ub, ue, k, s, nest, t=None, wrk=None, iwrk=None):
    """x [m][idim] (memory order of Fortran x(idim,m)); returns t, c [idim][nk1], dict."""
    t = np.array([], float) if t is None else t
    wrk = np.array([], float) if wrk is None else wrk
    iwrk = np.array([], np.int32) if iwrk is None else iwrk
    return _fitpack.parcur(np.ascontiguousarray(x).ravel().copy(), w, u.copy(), ub, ue, k, iopt,
                           ipar, s, t, nest, wrk, iwrk, 0)


def synth_param_curve(rng, m, idim, sigma=0.05):
    th = np.sort(rng.uniform(0, 2 * np.pi, m))
    th = np.unique(th)
    cols = []
    for d in range(idim):
        f = rng.uniform(0.5, 3)
        cols.append(rng.uniform(.5, 2) * np.cos(f * th + rng.uniform(0, 6)))
    x = np.stack(cols, 1) + sigma * rng.standard_normal((th.size, idim))
    return th, x


def make_parcur():
    """parcur (ipar 0/1, iopt 0/-1 and 0 -> 1 continuations) + curev-by-splev(e=3) goldens."""
    rng = np.random.default_rng(4242)
    rec = dict(k=[], m=[], idim=[], ipar=[], iopt=[], s=[], nest=[], ub=[], ue=[], n=[], fp=[], ier=[],
               xoff=[], uoff=[], toff=[], coff=[], s0=[])
    xcat, ucat, uincat, wcat, tcat, ccat, evcat = [], [], [], [], [], [], []
    xoff = uoff = toff = coff = 0
    for trial in range(150):
        k = int(rng.integers(1, 6))
        idim = int(rng.choice([1, 2, 2, 3, 3, 4, 7, 10]))
        m = int(rng.integers(k + 3, 120))
        th, x = synth_param_curve(rng, m, idim)
        m = th.size
        w = np.ones(m) if rng.random() < 0.5 else rng.uniform(0.5, 2, m)
        ipar = int(rng.random() < 0.4)
        s = float(rng.choice([0.0, m * 0.0025 * idim, m * 0.0005, m * 0.05, 1e-9, 500.0]))
        nest = int(rng.choice([m + k + 1, max(2 * k + 2, m // 3), max(2 * k + 3, m // 2)]))
        mode = rng.choice(["fit", "lsq", "cont"], p=[0.6, 0.2, 0.2])
        u_in = th.copy() if ipar else np.zeros(m)
        ub, ue = (th[0] - 0.1 * (trial % 2), th[-1] + 0.05 * (trial % 3)) if ipar else (0.0, 1.0)
        s0 = -1.0
        if mode == "fit":
            iopt = 0
            t, c, o = scipy_parcur(0, ipar, idim, u_in, x, w, ub, ue, k, s, nest)
        elif mode == "lsq":
            iopt = -1
            nint = int(rng.integers(1, max(2, min(8, m - k - 2))))
            if ipar:
                lo, hi = ub, ue
                uu = th
            else:
                lo, hi = 0.0, 1.0
                # netlib chord-length parameter (to place interior knots inside the data)
                d = np.sqrt(((x[1:] - x[:-1]) ** 2).sum(1))
                uu = np.concatenate([[0], np.cumsum(d)]) / d.sum()
            tin = np.zeros(nint + 2 * k + 2)
            tin[k + 1:k + 1 + nint] = np.quantile(uu, np.linspace(0, 1, nint + 2)[1:-1])
            nest = tin.size
            t, c, o = scipy_parcur(-1, ipar, idim, u_in, x, w, lo, hi, k, 0.0, nest, t=tin)
            ub, ue = lo, hi
            s = 0.0
        else:
            # iopt=0 with a large s, then iopt=1 continuation with the final s
            iopt = 1
            s0 = float(max(s * 20, m * 0.5))
            t0, c0, o0 = scipy_parcur(0, ipar, idim, u_in, x, w, ub, ue, k, s0, nest)
            if o0["ier"] > 0:
                continue
            tt = np.zeros(nest)
            tt[:len(t0)] = t0
            t, c, o = scipy_parcur(1, ipar, idim, o0["u"], x, w, o0["ub"], o0["ue"], k, s, nest,
                                   t=tt[:len(t0)].copy(), wrk=o0["wrk"], iwrk=o0["iwrk"])
        ier = int(o["ier"])
        fp = float(o["fp"])
        if not np.isfinite(fp) and ier != 10:
            continue
        n = 0 if ier == 10 else len(t)
        nk1 = n - k - 1
        ev = np.zeros(0)
        if ier != 10:
            ue_pts = np.sort(rng.uniform(o["ub"] - 0.1, o["ue"] + 0.1, 23))
            cm = np.asarray(c).reshape(idim, nk1)
            ev = np.stack([np.asarray(_fitpack.splev(np.asarray(t), np.concatenate([cm[d], np.zeros(k + 1)]),
                                                      k, ue_pts, 3)[0]) for d in range(idim)], 1).ravel()
            ev = np.concatenate([ue_pts, ev])
        for key, v in (("k", k), ("m", m), ("idim", idim), ("ipar", ipar), ("iopt", iopt), ("s", s),
                       ("nest", nest), ("ub", ub), ("ue", ue), ("n", n), ("fp", fp), ("ier", ier),
                       ("xoff", xoff), ("uoff", uoff), ("toff", toff), ("coff", coff), ("s0", s0)):
            rec[key].append(v)
        xcat.append(x.ravel()); uincat.append(u_in); ucat.append(np.asarray(o["u"])); wcat.append(w)
        tcat.append(np.asarray(t)[:n]); ccat.append(np.asarray(c).ravel()[:max(nk1, 0) * idim])
        evcat.append(ev)
        xoff += x.size; uoff += m; toff += n; coff += max(nk1, 0) * idim
    ints = ("k", "m", "idim", "ipar", "iopt", "nest", "n", "ier")
    out = {key: np.array(v, np.int32 if key in ints else (np.int64 if key.endswith("off") else float))
           for key, v in rec.items()}
    evoff = np.cumsum([0] + [e.size for e in evcat])[:-1]
    np.savez_compressed(os.path.join(HERE, "parcur_golden.npz"), x=np.concatenate(xcat),
                        u_in=np.concatenate(uincat), u=np.concatenate(ucat), w=np.concatenate(wcat),
                        t=np.concatenate(tcat), c=np.concatenate(ccat), ev=np.concatenate(evcat),
                        evoff=np.array(evoff, np.int64), **out)
    print("parcur cases:", len(rec["k"]), "ier histogram:",
          {int(v): int((np.array(rec["ier"]) == v).sum()) for v in np.unique(rec["ier"])},
          "iopt:", {int(v): int((np.array(rec["iopt"]) == v).sum()) for v in np.unique(rec["iopt"])})


# ------------------------------------------------------------------ periodic curves
def percur_lwrk(m, k, nest):
    return m * (k + 1) + nest * (8 + 5 * k)


def scipy_percur(iopt, x, y, w, k, s, nest, t_in=None, wrk=None, iwrk=None):
    m = x.size
    t = np.zeros(nest) if t_in is None else t_in.copy()
    wrk = np.zeros(percur_lwrk(m, k, t.size)) if wrk is None else wrk
    iwrk = np.zeros(t.size, np.int32) if iwrk is None else iwrk
    n, t, c, fp, ier = _fitpack.percur(iopt, x, y, w, k, s, t, wrk, iwrk)
    return int(n), np.asarray(t), np.asarray(c), float(fp), int(ier), wrk, iwrk


def make_percur():
    """percur: iopt=0 over all degrees / s regimes / weights, iopt=-1 with given interior knots and
    iopt=0 -> 1 continuations (state carried in wrk/iwrk)."""
    rng = np.random.default_rng(909)
    rec = dict(k=[], m=[], iopt=[], s=[], s0=[], nest=[], n=[], fp=[], ier=[], off=[], toff=[])
    xcat, ycat, wcat, tcat, ccat = [], [], [], [], []
    off = toff = 0
    for trial in range(220):
        k = int(rng.integers(1, 6))
        m = int(rng.integers(k + 3, 140))
        x = np.unique(np.sort(rng.uniform(0, 2 * np.pi, m)))
        if rng.random() < 0.25:
            x = np.linspace(0, 2 * np.pi, m)
        m = x.size
        y = rng.uniform(.5, 2) * np.sin(x * rng.integers(1, 4) + rng.uniform(0, 6)) \
            + 0.3 * np.cos(2 * x) + 0.1 * rng.standard_normal(m)
        y[-1] = y[0]
        w = np.ones(m) if rng.random() < 0.5 else rng.uniform(0.5, 2, m)
        s = float(rng.choice([0.0, m * 0.01, m * 0.001, m * 0.1, 1e-8, 500.0]))
        nest = int(rng.choice([m + 2 * k, max(2 * k + 3, m // 3), max(2 * k + 3, m // 2)]))
        mode = rng.choice(["fit", "lsq", "cont"], p=[0.6, 0.2, 0.2])
        s0 = -1.0
        if mode == "fit":
            iopt = 0
            n, t, c, fp, ier, _, _ = scipy_percur(0, x, y, w, k, s, nest)
        elif mode == "lsq":
            iopt = -1
            nint = int(rng.integers(1, max(2, min(9, m - 3))))
            tin = np.zeros(nint + 2 * k + 2)
            tin[k + 1:k + 1 + nint] = np.quantile(x, np.linspace(0, 1, nint + 2)[1:-1])
            nest = tin.size
            n, t, c, fp, ier, _, _ = scipy_percur(-1, x, y, w, k, 0.0, nest, t_in=tin)
            s = 0.0
        else:
            # scipy's wrapper takes n = nest = len(t) on a continuation, so the replayable case is:
            # fit with a small s0 (many knots), then iopt=1 with a larger s on exactly those knots
            # (nest = n0): the state fp0/fpold/nplus is read back from wrk/iwrk and part 2 runs
            iopt = 1
            s0 = float(rng.choice([m * 0.001, m * 0.01]))
            s = float(s0 * rng.uniform(2, 30))
            n0, t0, c0, fp0, ier0, wrk, iwrk = scipy_percur(0, x, y, w, k, s0, nest)
            if ier0 > 0 or n0 <= 2 * k + 2:
                continue
            n, t, c, fp, ier = _fitpack.percur(1, x, y, w, k, s, t0[:n0].copy(), wrk, iwrk)
            n, t, c, fp, ier = int(n), np.asarray(t), np.asarray(c), float(fp), int(ier)
        if not np.isfinite(fp) and ier != 10:
            continue
        if ier == 10:
            n = 0
        for key, v in (("k", k), ("m", m), ("iopt", iopt), ("s", s), ("s0", s0), ("nest", nest), ("n", n),
                       ("fp", fp), ("ier", ier), ("off", off), ("toff", toff)):
            rec[key].append(v)
        xcat.append(x); ycat.append(y); wcat.append(w)
        tcat.append(t[:n]); ccat.append(c[:n])
        off += m
        toff += n
    ints = ("k", "m", "iopt", "nest", "n", "ier")
    out = {key: np.array(v, np.int32 if key in ints else (np.int64 if key.endswith("off") else float))
           for key, v in rec.items()}
    np.savez_compressed(os.path.join(HERE, "percur_golden.npz"), x=np.concatenate(xcat), y=np.concatenate(ycat),
                        w=np.concatenate(wcat), t=np.concatenate(tcat), c=np.concatenate(ccat), **out)
    print("percur cases:", len(rec["k"]), "ier histogram:",
          {int(v): int((np.array(rec["ier"]) == v).sum()) for v in np.unique(rec["ier"])},
          "iopt:", {int(v): int((np.array(rec["iopt"]) == v).sum()) for v in np.unique(rec["iopt"])})

if __name__ == "__main__":
    import sys
    sys.path.insert(0, os.path.dirname(HERE))
    makers = dict(mncurf=make_mncurf, curfit=make_curfit, splev=make_splev, regrid=make_regrid,
                  calculus=make_calculus, parcur=make_parcur, percur=make_percur)
    # python make_golden.py [name ...]: regenerate only the named fixtures (default: all)
    for name in (sys.argv[1:] or list(makers)):
        makers[name]()
        f = name + "_golden.npz"
        print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")
