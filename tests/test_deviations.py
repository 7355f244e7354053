"""CPU: the oracle's reference mode on the directed reference-vs-netlib deviation cases, against results
derived by hand from the reference source (tests/deviation_cases.py).  The netlib switch must change the
outcome of every case it covers -- proof that the case really exercises the deviation."""
import numpy as np

import deviation_cases as D
import oracle_lib as O


def _netlib(on):
    O.set_netlib_compat(1 if on else 0)


def test_fpknot_first_eligible_interval_with_nan_residuals():
    c = D.nan_case()
    n_exp, at = D.nan_case_expected()
    assert (n_exp, at) == (12, [2, 3, 4, 7])           # the hand trace in DESIGN.md section 2
    r = O.curfit(0, c["x"], c["y"], None, c["xb"], c["xe"], c["k"], c["s"], c["nest"])
    assert (r["n"], r["ier"]) == (n_exp, 1)
    assert np.array_equal(r["t"][4:8], c["x"][np.array(at) - 1])
    try:
        _netlib(True)
        q = O.curfit(0, c["x"], c["y"], None, c["xb"], c["xe"], c["k"], c["s"], c["nest"])
    finally:
        _netlib(False)
    assert q["n"] == 12 and not np.array_equal(q["t"][4:8], r["t"][4:8])   # netlib ends on the LAST interval


def test_subnormal_pivot_is_skipped_like_a_zero_pivot():
    res = {}
    for wf in (0.0, 1e-310):
        c = D.subnormal_case(wf)
        res[wf] = O.curfit(-1, c["x"], c["y"], c["w"], c["xb"], c["xe"], c["k"], 0.0, c["nest"], n=c["n"],
                           t=c["t"].copy())
    a, b = res[0.0], res[1e-310]
    assert a["ier"] == b["ier"] == 0 and a["fp"] == b["fp"]
    assert np.isnan(a["c"][0]) and np.isnan(b["c"][0])
    assert np.array_equal(a["c"][:8], b["c"][:8], equal_nan=True) and np.isfinite(a["c"][1:8]).all()
    try:
        _netlib(True)
        c = D.subnormal_case(1e-310)
        q = O.curfit(-1, c["x"], c["y"], c["w"], c["xb"], c["xe"], c["k"], 0.0, c["nest"], n=c["n"], t=c["t"].copy())
    finally:
        _netlib(False)
    assert q["c"][0] == c["y"][0]                          # netlib rotates the subnormal pivot in


def test_main_loop_runs_m_plus_one_knot_sets_when_fpknot_cannot_refine():
    c = D.spin_case()
    r = O.curfit(0, c["x"], c["y"], None, c["xb"], c["xe"], c["k"], c["s_first"], c["nest"])
    n, k = r["n"], c["k"]
    assert r["ier"] == 0 and n > 2 * (k + 1)
    iw = r["iwrk"].copy()
    iw[:n - 2 * (k + 1) + 1] = 0
    r2 = O.curfit(1, c["x"], c["y"], None, c["xb"], c["xe"], k, c["s_second"], c["nest"], n=n, t=r["t"].copy(),
                  wrk=r["wrk"].copy(), iwrk=iw)
    assert r2["stats"][0] == c["passes"] == c["x"].size + 1
    assert r2["n"] == n and np.array_equal(r2["t"][:n], r["t"][:n]) and r2["ier"] in (2, 3)


def test_fpchec_tests_before_advancing():
    c = D.fpchec_case()
    assert O.fpchec(c["x"], c["t"], c["n"], c["k"]) == c["expected"]


def test_parcur_chord_lengths_use_the_scaled_norm2():
    c = D.norm2_case()
    r = O.parcur(0, 0, c["idim"], c["x"], None, c["k"], c["s"], c["nest"])
    assert r["ier"] == -1 and np.isfinite(r["u"]).all() and np.abs(r["u"] - c["u"]).max() < 1e-15
    try:
        _netlib(True)
        q = O.parcur(0, 0, c["idim"], c["x"], None, c["k"], c["s"], c["nest"])
    finally:
        _netlib(False)
    assert np.isnan(q["u"][1:-1]).all()                  # sqrt(sum d**2) overflowed


def test_regrid_ties_go_to_the_last_axis():
    c = D.regrid_symmetric_case()
    g, z, k, nest = c["g"], c["z"], c["k"], c["nest"]
    seen_unequal = False
    for s in c["s_values"]:
        r = O.regrid(0, g, g, z, 0, 1, 0, 1, k, k, s, nest, nest)
        assert r["ier"] == 0 and r["ny"] >= r["nx"]
        try:
            _netlib(True)
            q = O.regrid(0, g, g, z, 0, 1, 0, 1, k, k, s, nest, nest)
        finally:
            _netlib(False)
        assert q["nx"] >= q["ny"]
        seen_unequal = seen_unequal or r["ny"] > r["nx"]
    assert seen_unequal
