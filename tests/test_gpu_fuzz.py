"""GPU differential (fuzz) test: many small random problems through the batch entry points against
the oracle, every output bit-exact.  Covers what the golden files only sample: all degrees, short /
odd point counts, tight and loose nest, zero / tiny / huge smoothing factors, weights, jittered and
clustered abscissae -- for curfit, parcur (idim 1, 2, 3, 5) and percur."""
import os

import numpy as np
import pytest

import oracle_lib as O

# FPB_FUZZ_SEED=<int> shifts every generator seed: `for s in 1 2 3; do FPB_FUZZ_SEED=$s pytest ...` widens the sweep
SEED = int(os.environ.get("FPB_FUZZ_SEED", "0")) * 1000003

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def F():
    import fitpack_b200
    return fitpack_b200


def _same(a, b):
    return np.array_equal(np.asarray(a), np.asarray(b), equal_nan=True)


def _abscissae(rng, B, m):
    """sorted, strictly increasing; a third of the curves clustered"""
    x = np.sort(rng.uniform(0, 1, (B, m)), axis=1)
    cl = rng.random(B) < 0.33
    x[cl] = np.sort(rng.beta(0.4, 0.4, (int(cl.sum()), m)), axis=1)
    x += np.arange(m)[None, :] * 1e-7
    return x


def _signal(rng, x):
    B = x.shape[0]
    f = rng.uniform(0.5, 6, (B, 1))
    y = rng.uniform(.2, 3, (B, 1)) * np.sin(2 * np.pi * f * x + rng.uniform(0, 6, (B, 1)))
    y += rng.uniform(0.01, 0.3, (B, 1)) * rng.standard_normal(x.shape)
    y[rng.random(B) < 0.05] = 1.5          # constant data: all residuals zero (fpknot seeding path)
    return y


def _check_fit(rg, ro, k, idim=1, tag=None):
    assert _same(rg["ier"], ro["ier"]), tag
    ok = ro["ier"] != 10
    assert _same(rg["n"][ok], ro["n"][ok]), tag
    assert _same(rg["fp"][ok], ro["fp"][ok]), tag
    for b in np.nonzero(ok)[0]:
        n = ro["n"][b]
        if ro["ier"][b] == 1 and n > ro["t"].shape[1]:
            continue                        # s=0 with nest too small: n = nmax reported, nothing fitted
        assert _same(rg["t"][b, :n], ro["t"][b, :n]), (tag, b)
        for d in range(idim):
            assert _same(rg["c"][b, d * n:d * n + n - k - 1], ro["c"][b, d * n:d * n + n - k - 1]), (tag, b, d)
    return np.bincount(ro["ier"] + 2, minlength=13)


def test_fuzz_curfit_bit_exact(F):
    rng = np.random.default_rng(20261020 + SEED)
    hist = np.zeros(13, int)
    B = 160
    for k in (1, 2, 3, 4, 5):
        for m in (k + 2, 9, 26, 61, 130):
            if m < k + 2:
                continue
            x = _abscissae(rng, B, m)
            y = _signal(rng, x)
            for weights in (False, True):
                w = rng.uniform(0.3, 3.0, (B, m)) if weights else None
                for s in (0.0, 1e-9, m * 0.002, m * 0.05, 1e4):
                    for nest in (m + k + 1, max(2 * k + 2, m // 3)):
                        ro = O.curfit_batch(0, x, y, w, k, s, nest)
                        rg = F.curfit_batch(0, x, y, w, k, s, nest)
                        hist += _check_fit(rg, ro, k, tag=("curfit", k, m, weights, s, nest))
    assert hist[0] > 0 and hist[1] > 0 and hist[2] > 0 and hist[3] > 0   # ier -2, -1, 0, 1 all exercised


def test_fuzz_parcur_bit_exact(F):
    rng = np.random.default_rng(20261021 + SEED)
    hist = np.zeros(13, int)
    B = 120
    for idim in (1, 2, 3, 5):
        for k in (1, 3, 4, 5):
            for m in (k + 2, 17, 64):
                u = _abscissae(rng, B, m) * 3.0
                x = np.stack([_signal(rng, u / 3.0) for _ in range(idim)], axis=2)
                for weights, ipar in ((False, 0), (True, 1), (True, 0)):
                    w = rng.uniform(0.3, 3.0, (B, m)) if weights else None
                    kw = dict(u=u, ub=u[:, 0] - 0.01, ue=u[:, -1] + 0.02) if ipar else {}
                    for s in (0.0, m * idim * 0.003, m * idim * 0.1, 1e5):
                        for nest in (m + k + 1, max(2 * k + 2, m // 2)):
                            ro = O.parcur_batch(0, ipar, idim, x, w, k, s, nest, **kw)
                            rg = F.parcur_batch(0, ipar, idim, x, w, k, s, nest, **kw)
                            assert _same(rg["u"], ro["u"])
                            hist += _check_fit(rg, ro, k, idim, tag=("parcur", idim, k, m, weights, ipar, s, nest))
    assert hist[0] > 0 and hist[1] > 0 and hist[2] > 0 and hist[3] > 0


def test_fuzz_percur_bit_exact(F):
    rng = np.random.default_rng(20261022 + SEED)
    hist = np.zeros(13, int)
    B = 160
    for k in (1, 2, 3, 4, 5):
        for m in (k + 3, 12, 40, 97):
            x = _abscissae(rng, B, m) * 2 * np.pi
            y = _signal(rng, x / (2 * np.pi))
            y[:, -1] = y[:, 0]
            for weights in (False, True):
                w = rng.uniform(0.3, 3.0, (B, m)) if weights else None
                for s in (0.0, 1e-9, m * 0.003, m * 0.08, 1e4):
                    for nest in (m + 2 * k, max(2 * k + 3, m // 2)):
                        ro = O.percur_batch(0, x, y, w, k, s, nest)
                        rg = F.percur_batch(0, x, y, w, k, s, nest)
                        hist += _check_fit(rg, ro, k, tag=("percur", k, m, weights, s, nest))
    assert hist[0] > 0 and hist[1] > 0 and hist[2] > 0 and hist[3] > 0


def test_fuzz_curfit_given_knots_bit_exact(F):
    """iopt = -1: random interior knot sets, some violating the Schoenberg-Whitney conditions (fpchec
    -> ier = 10 for that curve only), per-curve knot counts."""
    rng = np.random.default_rng(20261023 + SEED)
    B = 200
    n10 = nok = 0
    for k in (1, 3, 5):
        for m in (14, 50, 120):
            x = _abscissae(rng, B, m)
            y = _signal(rng, x)
            w = rng.uniform(0.3, 3.0, (B, m))
            nest = m + k + 1
            n = np.zeros(B, np.int32)
            t = np.zeros((B, nest))
            for b in range(B):
                nint = int(rng.integers(0, max(1, min(12, m - k - 1))))
                kn = np.sort(rng.uniform(x[b, 0], x[b, -1], nint))
                if rng.random() < 0.15 and nint > 2:
                    kn[1] = kn[0]                       # coincident interior knots
                if rng.random() < 0.1 and nint > 0:
                    kn[-1] = x[b, -1] + 0.5             # knot outside [xb, xe]
                n[b] = nint + 2 * k + 2
                t[b, k + 1:k + 1 + nint] = kn
            ro = O.curfit_batch(-1, x, y, w, k, 0.0, nest, n=n.copy(), t=t.copy())
            rg = F.curfit_batch(-1, x, y, w, k, 0.0, nest, n=n.copy(), t=t.copy())
            assert _same(rg["ier"], ro["ier"]), (k, m)
            ok = ro["ier"] != 10
            n10 += int((~ok).sum())
            nok += int(ok.sum())
            assert _same(rg["fp"][ok], ro["fp"][ok]) and _same(rg["n"], ro["n"]), (k, m)
            for b in np.nonzero(ok)[0]:
                nn = ro["n"][b]
                assert _same(rg["t"][b, :nn], ro["t"][b, :nn]) and _same(rg["c"][b, :nn - k - 1], ro["c"][b, :nn - k - 1]), (k, m, b)
    assert n10 > 20 and nok > 1000


def test_minimal_sizes_bit_exact(F):
    """Smallest legal problems: m = k+1 / k+2 points (curfit, parcur), m = 2, 3, 4 (percur), nest from the
    minimum 2k+2 up -- every ier class these produce (-2, -1, 0, 1, 10) must match the oracle."""
    rng = np.random.default_rng(5 + SEED)
    B = 40
    for k in (1, 2, 3, 4, 5):
        for m in (k + 1, k + 2):
            x = np.sort(rng.uniform(0, 1, (B, m)), axis=1) + np.arange(m) * 1e-6
            y = rng.standard_normal((B, m))
            xx = np.stack([y, y[:, ::-1]], 2)
            for s in (0.0, 0.01, 10.0):
                for nest in (2 * k + 2, m + k + 1, m + k + 5):
                    _check_fit(F.curfit_batch(0, x, y, None, k, s, nest), O.curfit_batch(0, x, y, None, k, s, nest),
                               k, tag=("curfit", k, m, s, nest))
                    kw = dict(u=x, ub=x[:, 0].copy(), ue=x[:, -1].copy())
                    _check_fit(F.parcur_batch(0, 1, 2, xx, None, k, s, nest, **kw),
                               O.parcur_batch(0, 1, 2, xx, None, k, s, nest, **kw), k, 2, tag=("parcur", k, m, s, nest))
        for m in (2, 3, 4, k + 2):
            x = np.sort(rng.uniform(0, 6, (B, m)), axis=1) + np.arange(m) * 1e-6
            y = rng.standard_normal((B, m))
            for s in (0.0, 0.01, 10.0):
                for nest in (2 * k + 2, 2 * k + 3, m + 2 * k, m + 2 * k + 3):
                    _check_fit(F.percur_batch(0, x, y, None, k, s, nest), O.percur_batch(0, x, y, None, k, s, nest),
                               k, tag=("percur", k, m, s, nest))
