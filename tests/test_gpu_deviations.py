"""GPU: the CUDA path, through the C-ABI, on the directed reference-vs-netlib deviation cases
(tests/deviation_cases.py): it must give the hand-derived result AND the oracle's reference-mode bits."""
import numpy as np
import pytest

import deviation_cases as D
import oracle_lib as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def F():
    import fitpack_b200
    return fitpack_b200


def _same(a, b):
    return np.array_equal(np.asarray(a), np.asarray(b), equal_nan=True)


def test_fpknot_first_eligible_interval_with_nan_residuals(F):
    c = D.nan_case()
    n_exp, at = D.nan_case_expected()
    g = F.curfit(0, c["x"], c["y"], None, c["xb"], c["xe"], c["k"], c["s"], c["nest"])
    o = O.curfit(0, c["x"], c["y"], None, c["xb"], c["xe"], c["k"], c["s"], c["nest"])
    assert (g["n"], g["ier"]) == (n_exp, 1) == (o["n"], o["ier"])
    assert _same(g["t"][4:8], c["x"][np.array(at) - 1])
    assert _same(g["t"][:12], o["t"][:12]) and _same(g["c"][:8], o["c"][:8]) and _same(g["fp"], o["fp"])


def test_subnormal_pivot_is_skipped_like_a_zero_pivot(F):
    res = {}
    for wf in (0.0, 1e-310):
        c = D.subnormal_case(wf)
        res[wf] = F.curfit(-1, c["x"], c["y"], c["w"], c["xb"], c["xe"], c["k"], 0.0, c["nest"], n=c["n"],
                           t=c["t"].copy())
        o = O.curfit(-1, c["x"], c["y"], c["w"], c["xb"], c["xe"], c["k"], 0.0, c["nest"], n=c["n"],
                     t=c["t"].copy())
        assert res[wf]["ier"] == o["ier"] == 0 and res[wf]["fp"] == o["fp"] and _same(res[wf]["c"][:8], o["c"][:8])
    a, b = res[0.0], res[1e-310]
    assert np.isnan(a["c"][0]) and np.isnan(b["c"][0]) and np.isfinite(a["c"][1:8]).all()
    assert _same(a["c"][:8], b["c"][:8]) and a["fp"] == b["fp"]


def test_main_loop_runs_m_plus_one_knot_sets_when_fpknot_cannot_refine(F):
    import torch
    from fitpack_b200.batch import Engine
    c = D.spin_case()
    k, nest, m = c["k"], c["nest"], c["x"].size
    o1 = O.curfit(0, c["x"], c["y"], None, c["xb"], c["xe"], k, c["s_first"], nest)
    n = o1["n"]
    iw = o1["iwrk"].copy()
    iw[:n - 2 * (k + 1) + 1] = 0
    o2 = O.curfit(1, c["x"], c["y"], None, c["xb"], c["xe"], k, c["s_second"], nest, n=n, t=o1["t"].copy(),
                  wrk=o1["wrk"].copy(), iwrk=iw.copy())
    # flat ABI (state in wrk/iwrk)
    g2 = F.curfit(1, c["x"], c["y"], None, c["xb"], c["xe"], k, c["s_second"], nest, n=n, t=o1["t"].copy(),
                  wrk=o1["wrk"].copy(), iwrk=iw.copy(), lwrk=o1["wrk"].size)
    assert (g2["n"], g2["ier"]) == (n, o2["ier"]) and g2["ier"] in (2, 3) and g2["fp"] == o2["fp"]
    assert _same(g2["t"][:n], o1["t"][:n]) and _same(g2["c"][:n - k - 1], o2["c"][:n - k - 1])
    # batch ABI with per-curve statistics: m+1 least-squares passes (core:4848)
    dev = torch.device("cuda", 0)
    eng = Engine(dev)
    B = 3
    xd = torch.from_numpy(np.tile(c["x"], (B, 1))).to(dev)
    yd = torch.from_numpy(np.tile(c["y"], (B, 1))).to(dev)
    r = eng.curfit_batch(xd, yd, None, k=k, s=c["s_second"], nest=nest, iopt=1, xb=c["xb"], xe=c["xe"],
                         layout="curve_major",
                         n=torch.full((B,), n, dtype=torch.int32, device=dev),
                         t=torch.from_numpy(np.tile(o1["t"], (B, 1))).to(dev),
                         fpint=torch.from_numpy(np.tile(o1["wrk"][:nest], (B, 1))).to(dev),
                         nrdata=torch.from_numpy(np.tile(iw[:nest], (B, 1))).to(dev), want_stats=True)
    torch.cuda.synchronize()
    st = r["stats"].cpu().numpy()
    assert (st[:, 0] == m + 1).all() and (st[:, 0] == o2["stats"][0]).all(), st
    assert (r["n"].cpu().numpy() == n).all() and (r["ier"].cpu().numpy() == o2["ier"]).all()
    eng.close()


def test_fpchec_tests_before_advancing(F):
    """iopt=-1 with the knot set netlib's walk rejects: the reference accepts it (the rank-deficient fit
    that follows yields non-finite coefficients in both the oracle and the CUDA path -- compared as bits)."""
    c = D.fpchec_case()
    y = np.array([1.0, 2.0, 3.0, 4.0])
    t = np.zeros(8)
    t[:6] = c["t"]
    g = F.curfit(-1, c["x"], y, None, 0.0, 10.0, c["k"], 0.0, 8, n=c["n"], t=t.copy())
    o = O.curfit(-1, c["x"], y, None, 0.0, 10.0, c["k"], 0.0, 8, n=c["n"], t=t.copy())
    assert g["ier"] == o["ier"] != 10
    assert _same(g["c"][:4], o["c"][:4]) and _same(g["fp"], o["fp"])


def test_parcur_chord_lengths_use_the_scaled_norm2(F):
    c = D.norm2_case()
    g = F.parcur(0, 0, c["idim"], c["x"], None, c["k"], c["s"], c["nest"])
    o = O.parcur(0, 0, c["idim"], c["x"], None, c["k"], c["s"], c["nest"])
    assert g["ier"] == o["ier"] == -1 and np.isfinite(g["u"]).all() and np.abs(g["u"] - c["u"]).max() < 1e-15
    assert _same(g["u"], o["u"]) and g["n"] == o["n"] and _same(g["c"], o["c"])


def test_regrid_ties_go_to_the_last_axis(F):
    c = D.regrid_symmetric_case()
    g, z, k, nest = c["g"], c["z"], c["k"], c["nest"]
    for s in c["s_values"]:
        r = F.regrid(0, g, g, z, 0, 1, 0, 1, k, k, s, nest, nest)
        o = O.regrid(0, g, g, z, 0, 1, 0, 1, k, k, s, nest, nest)
        assert (r["nx"], r["ny"], r["ier"]) == (o["nx"], o["ny"], o["ier"]) and r["ny"] >= r["nx"]
        assert _same(r["tx"][:r["nx"]], o["tx"][:o["nx"]]) and _same(r["ty"][:r["ny"]], o["ty"][:o["ny"]])
