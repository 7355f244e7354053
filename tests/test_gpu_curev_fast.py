"""GPU: curev FAST mode (one fast-splev pass per coordinate) against the oracle's curev: stated tolerance
1e-12 * max(1, max|c|), identical ier (curves whose parameters decrease are rejected with ier = 10 and left
untouched), arguments clamped to the knot range like the reference (core:1163-1165)."""
import numpy as np
import pytest

import oracle_lib as O
from refdata import synth_param_curves

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("idim,k", [(1, 3), (2, 3), (3, 5), (2, 1), (5, 2)])
def test_curev_fast_matches_exact_within_tolerance(idim, k):
    import torch
    import fitpack_b200
    from fitpack_b200.batch import Engine
    dev = torch.device("cuda", 0)
    eng = Engine(dev)
    rng = np.random.default_rng(100 + 10 * idim + k)
    B, m = 300, 64
    th, xp = synth_param_curves(rng, B, m, idim)
    nest = m + k + 1
    fit = O.parcur_batch(0, 0, idim, xp, None, k, m * idim * 0.0025, nest)
    ue = np.sort(rng.uniform(-0.05, 1.05, (B, 101)), axis=1)     # odd count: the unaligned point path too
    ue[7, 50] = ue[7, 49] - 1e-3                                  # a decreasing parameter: ier = 10
    t_d = torch.from_numpy(fit["t"]).to(dev)
    n_d = torch.from_numpy(fit["n"]).to(dev)
    c_d = torch.from_numpy(fit["c"]).to(dev)
    u_d = torch.from_numpy(ue).to(dev)
    xo, io = O.curev_batch(idim, fit["t"], fit["n"], fit["c"], k, ue)
    x0 = torch.full((B, 101, idim), -7.0, device=dev, dtype=torch.float64)
    i0 = torch.zeros(B, device=dev, dtype=torch.int32)
    xe, ie = eng.curev_batch(idim, t_d, n_d, c_d, k, u_d, mode=0)
    xf, jf = eng.curev_batch(idim, t_d, n_d, c_d, k, u_d, out=(x0, i0), mode=1)
    torch.cuda.synchronize()
    xe, xf, ie, jf = xe.cpu().numpy(), xf.cpu().numpy(), ie.cpu().numpy(), jf.cpu().numpy()
    assert np.array_equal(ie, io) and np.array_equal(jf, io) and io[7] == 10 and (np.delete(io, 7) == 0).all()
    assert (xf[7] == -7.0).all()                                   # rejected curve: nothing written
    ok = io == 0
    assert np.array_equal(xe[ok], xo[ok])                          # exact mode: the oracle's bits
    scale = np.maximum(1.0, np.abs(fit["c"]).max(axis=1))[ok, None, None]
    assert (np.abs(xf[ok] - xo[ok]) / scale).max() < 1e-12
    # even point count (TMA point tiles) gives the same values as the odd one on the shared prefix
    u2 = torch.from_numpy(np.ascontiguousarray(ue[:, :100])).to(dev)
    x2, j2 = eng.curev_batch(idim, t_d, n_d, c_d, k, u2, mode=1)
    torch.cuda.synchronize()
    assert np.array_equal(x2.cpu().numpy()[ok][:, :49], xf[ok][:, :49])
    eng.close()


def test_bispev_fast_matches_exact_within_tolerance():
    """bispev FAST (the reference's sums as fused multiply-adds): values within 1e-12 * max(1, max|c|) of the
    exact mode, which itself is the oracle's bits."""
    import torch
    import fitpack_b200
    from fitpack_b200.batch import Engine
    dev = torch.device("cuda", 0)
    eng = Engine(dev)
    rng = np.random.default_rng(4)
    for kx, ky, mx, my in ((3, 3, 70, 65), (1, 2, 40, 33), (5, 4, 64, 50)):
        gx, gy = np.linspace(-1.0, 1.0, mx), np.linspace(0.0, 2.0, my)
        GX, GY = np.meshgrid(gx, gy, indexing="ij")
        z = np.cos(2 * GX) * np.sin(GY) + 0.05 * rng.standard_normal(GX.shape)
        nxe, nye = mx + kx + 1, my + ky + 1
        o = O.regrid(0, gx, gy, z, -1, 1, 0, 2, kx, ky, mx * my * 0.0025, nxe, nye)
        ex = np.sort(rng.uniform(-1, 1, 90))
        ey = np.sort(rng.uniform(0, 2, 77))
        zo, io = O.bispev(o["tx"], o["nx"], o["ty"], o["ny"], o["c"], kx, ky, ex, ey)
        out0 = torch.empty((90, 77), device=dev, dtype=torch.float64)
        out1 = torch.empty((90, 77), device=dev, dtype=torch.float64)
        fitpack_b200.bispev(o["tx"], o["nx"], o["ty"], o["ny"], o["c"], kx, ky, ex, ey, out=out0, engine=eng, mode=0)
        fitpack_b200.bispev(o["tx"], o["nx"], o["ty"], o["ny"], o["c"], kx, ky, ex, ey, out=out1, engine=eng, mode=1)
        torch.cuda.synchronize()
        assert io == 0 and np.array_equal(out0.cpu().numpy(), zo)
        scale = max(1.0, np.abs(o["c"]).max())
        assert np.abs(out1.cpu().numpy() - zo).max() / scale < 1e-12
    eng.close()


@pytest.mark.parametrize("k", [2, 3, 5])
def test_exact_evaluators_with_and_without_the_span_table(k):
    """Exact splev / curev use a per-interval span table for splines with up to 96 knot intervals and plain
    divisions beyond: both paths, and a batch that mixes them, must give the oracle's bits."""
    import torch
    from fitpack_b200.batch import Engine
    dev = torch.device("cuda", 0)
    eng = Engine(dev)
    rng = np.random.default_rng(50 + k)
    B, ld, m = 64, 420, 333
    t = np.zeros((B, ld))
    c = rng.standard_normal((B, ld))
    n = np.zeros(B, np.int32)
    for b in range(B):
        nint = int(rng.choice([3, 40, 95, 96, 97, 150, 400 - 2 * k]))     # knot intervals: around the table's cap
        nb = nint + 2 * k + 1
        inner = np.sort(rng.uniform(0.0, 1.0, nint - 1))
        t[b, :k + 1] = 0.0
        t[b, k + 1:k + nint] = inner
        t[b, k + nint:nb] = 1.0
        n[b] = nb
    x = np.sort(rng.uniform(-0.05, 1.05, (B, m)), axis=1)
    yo, io = O.splev_batch(t, n, c, k, x)
    r = eng.splev_batch(torch.from_numpy(t).to(dev), torch.from_numpy(n).to(dev), torch.from_numpy(c).to(dev), k,
                        torch.from_numpy(x).to(dev), e=0, mode=0)
    torch.cuda.synchronize()
    assert np.array_equal(r["y"].cpu().numpy(), yo) and not r["ier"].cpu().numpy().any()
    # curev on the same splines (idim = 1: c(1:n-k-1) of each row), parameters clamped to [0, 1] by curev itself
    xo, jo = O.curev_batch(1, t, n, c, k, x)
    xe, je = eng.curev_batch(1, torch.from_numpy(t).to(dev), torch.from_numpy(n).to(dev),
                             torch.from_numpy(c).to(dev), k, torch.from_numpy(x).to(dev), mode=0)
    torch.cuda.synchronize()
    assert np.array_equal(xe.cpu().numpy(), xo) and np.array_equal(je.cpu().numpy(), jo)
    eng.close()
