"""The warp-per-curve kernel behind fp_curfit_c / the handle API (fitpack_b200/csrc/fpb_curfit_solo.cu) against the
CPU oracle -- n, knots, ier, c, fp and the continuation state bit for bit -- and the same calls through the batch
kernels (fpb_set_single_curve_kernel(0)), so that both implementations keep replaying the reference's flat-ABI tests
(mncurf test/fitpack_tests.f90:1138-1269, the golden vectors, the deviation cases, the handle API)."""
import contextlib

import numpy as np
import pytest

import oracle_lib as O
import test_gpu_deviations
import test_gpu_parity
from test_gpu_parity import F, eng, _same  # noqa: F401  (fixtures)

pytestmark = pytest.mark.gpu


@contextlib.contextmanager
def single_curve_kernel(Fm, on):
    prev = Fm.lib().fpb_set_single_curve_kernel(1 if on else 0)
    try:
        yield
    finally:
        Fm.lib().fpb_set_single_curve_kernel(prev)


def _check(o, g, k):
    assert (g["n"], g["ier"]) == (o["n"], o["ier"])
    if o["ier"] == 10:
        return
    n = int(o["n"])
    assert g["fp"] == o["fp"]
    assert _same(g["t"][:n], o["t"][:n]) and _same(g["c"][:n - k - 1], o["c"][:n - k - 1])


def test_default_is_the_warp_per_curve_kernel(F):
    assert F.lib().fpb_set_single_curve_kernel(1) == 1


@pytest.mark.parametrize("name", ["test_mncurf_sequence_bit_exact", "test_curfit_golden_vectors_bit_exact",
                                  "test_curfit_input_errors_leave_outputs_untouched", "test_curve_handle_api",
                                  "test_host_batch_api_matches_oracle",
                                  "test_parcur_golden_cases_flat_abi_bit_exact",
                                  "test_percur_golden_cases_flat_abi_bit_exact"])
def test_flat_abi_tests_also_pass_on_the_batch_kernels(F, name):
    """the tests of test_gpu_parity run on the warp-per-curve kernel by default; here on one thread per curve"""
    with single_curve_kernel(F, False):
        getattr(test_gpu_parity, name)(F)


def test_deviation_cases_on_both_kernels(F):
    names = [n for n in dir(test_gpu_deviations) if n.startswith("test_")]
    assert names
    for on in (True, False):
        with single_curve_kernel(F, on):
            for n in names:
                getattr(test_gpu_deviations, n)(F)


@pytest.mark.parametrize("k", [1, 2, 3, 4, 5])
def test_random_fits_continuations_and_given_knots_bit_exact(F, k):
    """iopt = 0, then iopt = 1 with a smaller s on the returned state, then iopt = -1 on the knots found; weights
    with zeros (skipped rotations), s = 0 (interpolation), s so large that the polynomial is accepted, small nest
    (ier = 1), m down to 2k+2."""
    rng = np.random.default_rng(1000 + k)
    for trial in range(60):
        m = int(rng.integers(2 * k + 2, 330))
        x = np.sort(rng.uniform(0, 1, m))
        x[0], x[-1] = 0.0, 1.0
        y = np.sin(2 * np.pi * rng.uniform(0.5, 5) * x) + 0.05 * rng.standard_normal(m)
        w = rng.uniform(0.5, 2.0, m) if trial % 3 else np.ones(m)
        if trial % 11 == 0:
            w[rng.integers(0, m)] = 0.0
        s = float(m * 0.0025 * rng.choice([0.0, 0.03, 0.3, 1.0, 3.0, 3000.0]))
        nest = int(rng.choice([m + k + 1, max(2 * k + 2, m // 2), max(102, int(np.ceil(1.4 * m)))]))
        o = O.curfit(0, x, y, w, 0.0, 1.0, k, s, nest)
        g = F.curfit(0, x, y, w, 0.0, 1.0, k, s, nest)
        _check(o, g, k)
        if o["ier"] > 0 or s == 0.0:
            continue
        n = int(o["n"])
        nrint = n - 2 * (k + 1) + 1
        assert _same(g["wrk"][n - 2:n], o["wrk"][n - 2:n]) and _same(g["iwrk"][:nrint], o["iwrk"][:nrint])
        assert g["iwrk"][n - 1] == o["iwrk"][n - 1]
        s2 = s * 0.4
        o1 = O.curfit(1, x, y, w, 0.0, 1.0, k, s2, nest, n=o["n"], t=o["t"], c=o["c"], fp=o["fp"], wrk=o["wrk"],
                      iwrk=o["iwrk"])
        g1 = F.curfit(1, x, y, w, 0.0, 1.0, k, s2, nest, n=g["n"], t=g["t"], c=g["c"], fp=g["fp"], wrk=g["wrk"],
                      iwrk=g["iwrk"])
        _check(o1, g1, k)
        o2 = O.curfit(-1, x, y, w, 0.0, 1.0, k, s2, nest, n=o1["n"], t=o1["t"].copy())
        g2 = F.curfit(-1, x, y, w, 0.0, 1.0, k, s2, nest, n=g1["n"], t=g1["t"].copy())
        _check(o2, g2, k)


def test_given_knots_rejected_by_fpchec_and_bad_arguments(F):
    """the data checks of curfit (core:1265-1285) run on the host for a batch of one: same ier, outputs untouched
    except the clamped end knots (core:1278-1284)"""
    rng = np.random.default_rng(5)
    m, k, nest = 40, 3, 50
    x = np.sort(rng.uniform(0, 1, m))
    y = rng.standard_normal(m)
    n = 12
    for variant in range(4):
        t = np.zeros(nest)
        t[k + 1:n - k - 1] = np.linspace(0.1, 0.9, n - 2 * k - 2)
        if variant == 1:
            t[k + 2] = t[k + 1]                       # not strictly increasing
        if variant == 2:
            t[k + 1:n - k - 1] = x[0] + 1e-9 * np.arange(1, n - 2 * k - 1)   # Schoenberg-Whitney fails
        if variant == 3:
            t[k + 1] = 2.0                            # beyond xe
        to, tg = t.copy(), t.copy()
        o = O.curfit(-1, x, y, None, x[0], x[-1], k, 0.0, nest, n=n, t=to)
        c0 = np.full(nest, 7.0)
        g = F.curfit(-1, x, y, None, x[0], x[-1], k, 0.0, nest, n=n, t=tg, c=c0)
        assert g["ier"] == o["ier"], variant
        assert _same(tg[:n], to[:n]), variant
        if o["ier"] == 10:
            assert np.all(c0 == 7.0), variant
        else:
            _check(o, g, k)
    for bad in (dict(xb=x[0] + 0.01), dict(xe=x[-1] - 0.01), dict(s=-1.0), dict(s=0.0, nest=m + k)):
        a = dict(xb=x[0], xe=x[-1], s=1.0, nest=nest)
        a.update(bad)
        o = O.curfit(0, x, y, None, a["xb"], a["xe"], k, a["s"], a["nest"])
        g = F.curfit(0, x, y, None, a["xb"], a["xe"], k, a["s"], a["nest"])
        assert g["ier"] == o["ier"] == 10, bad


def test_curves_too_large_for_shared_memory_take_the_batch_kernels(F):
    """the warp-per-curve kernel keeps the curve in shared memory (200 KB); beyond that the same call runs on the batch
    kernels -- same bits"""
    rng = np.random.default_rng(9)
    # 820 points with the handle API's nest = 1.4 m still fit (the shared arrays are min(nest, m+k+1) long) ...
    m, k = 820, 3
    x = np.sort(rng.uniform(0, 1, m))
    y = np.sin(9 * x) + 0.1 * rng.standard_normal(m)
    nest = max(102, int(np.ceil(np.float32(1.4) * m)))
    _check(O.curfit(0, x, y, None, x[0], x[-1], k, 8.0, nest), F.curfit(0, x, y, None, x[0], x[-1], k, 8.0, nest), k)
    # ... 3000 points do not
    m, k = 3000, 3
    x = np.sort(rng.uniform(0, 1, m))
    y = np.sin(9 * x) + 0.1 * rng.standard_normal(m)
    nest = m + k + 1
    o = O.curfit(0, x, y, None, x[0], x[-1], k, 30.0, nest)
    g = F.curfit(0, x, y, None, x[0], x[-1], k, 30.0, nest)
    _check(o, g, k)


def test_batch_entry_point_with_one_curve(F):
    """fp_curfit_batch_c with nbatch = 1 is the same call pattern"""
    rng = np.random.default_rng(10)
    m, k, nest = 200, 3, 204
    x = np.sort(rng.uniform(0, 1, m))
    y = np.cos(5 * x) + 0.05 * rng.standard_normal(m)
    o = O.curfit(0, x, y, None, x[0], x[-1], k, 0.5, nest)
    for on in (True, False):
        with single_curve_kernel(F, on):
            g = F.curfit_batch(0, x[None, :], y[None, :], None, k, 0.5, nest)
            n = int(o["n"])
            assert int(g["n"][0]) == n and int(g["ier"][0]) == o["ier"] and g["fp"][0] == o["fp"]
            assert _same(g["t"][0, :n], o["t"][:n]) and _same(g["c"][0, :n - k - 1], o["c"][:n - k - 1])


@pytest.mark.parametrize("weights", [False, True])
def test_small_host_batches_run_one_warp_per_curve_and_match_the_oracle(F, weights):
    """fp_curfit_batch_c with up to a few thousand curves: one CTA of one warp per curve instead of one thread per
    curve; rejected curves (unsorted x) keep the caller's n / t / c / fp; iopt = -1 with per-curve knots"""
    rng = np.random.default_rng(77 + weights)
    B, m, k, nest = 300, 120, 3, 124
    x = np.sort(rng.uniform(0, 1, (B, m)), axis=1)
    y = np.sin(2 * np.pi * rng.uniform(0.5, 3, (B, 1)) * x) + 0.05 * rng.standard_normal((B, m))
    w = rng.uniform(0.5, 2.0, (B, m)) if weights else None
    x[17, 5], x[17, 6] = x[17, 6], x[17, 5]          # curve 17: not sorted -> ier = 10
    s = 0.3
    o = O.curfit_batch(0, x, y, w, k, s, nest)
    t_in = np.full((B, nest), -3.0)
    g = F.curfit_batch(0, x, y, w, k, s, nest, t=t_in.copy())
    assert np.array_equal(g["n"], o["n"]) and np.array_equal(g["ier"], o["ier"]) and g["ier"][17] == 10
    for b in range(B):
        if b == 17:
            assert np.all(g["t"][b] == -3.0) and g["n"][b] == 0
            continue
        n = int(o["n"][b])
        assert g["fp"][b] == o["fp"][b]
        assert _same(g["t"][b, :n], o["t"][b, :n]) and _same(g["c"][b, :n - k - 1], o["c"][b, :n - k - 1])
        assert np.all(g["t"][b, n:] == -3.0)          # beyond n the caller's array is left alone
    # least squares on the knots just found (iopt = -1, per-curve n and t)
    o2 = O.curfit_batch(-1, x, y, w, k, s, nest, n=o["n"].copy(), t=o["t"].copy())
    g2 = F.curfit_batch(-1, x, y, w, k, s, nest, n=g["n"].copy(), t=g["t"].copy())
    assert np.array_equal(g2["ier"], o2["ier"])
    for b in range(B):
        if o2["ier"][b] == 10:
            continue
        n = int(o2["n"][b])
        assert g2["fp"][b] == o2["fp"][b] and _same(g2["c"][b, :n - k - 1], o2["c"][b, :n - k - 1])


def _same_parcur(o, g, idim, k):
    assert (g["n"], g["ier"]) == (o["n"], o["ier"])
    assert _same(g["u"], o["u"]) and g["ub"] == o["ub"] and g["ue"] == o["ue"]
    if o["ier"] == 10:
        return
    n = int(o["n"])
    assert g["fp"] == o["fp"] and _same(g["t"][:n], o["t"][:n])
    for d in range(idim):
        assert _same(g["c"][d * n:d * n + n - k - 1], o["c"][d * n:d * n + n - k - 1])


@pytest.mark.parametrize("idim", [2, 3])
def test_parcur_flat_abi_on_the_warp_per_curve_kernel(F, idim):
    """fp_parcur_c (idim 2, 3) = the same kernel with idim right-hand sides per band row (fppara core:8822-9177):
    chord-length and given parameters, iopt 0 -> 1 -> -1, weights, s from interpolation to 'polynomial accepted';
    n, ier, fp, t, c (in the reference's c(d*n + i) layout), u, ub, ue bit for bit"""
    rng = np.random.default_rng(500 + idim)
    for trial in range(80):
        k = int(rng.integers(1, 6))
        m = int(rng.integers(2 * k + 2, 260))
        ph = np.sort(rng.uniform(0, 1, m)) * rng.uniform(1, 6)
        x = np.stack([np.cos(ph), np.sin(ph), ph * 0.3][:idim], 1) + 0.03 * rng.standard_normal((m, idim))
        w = rng.uniform(0.5, 2.0, m) if trial % 3 else None
        s = float(m * 0.001 * rng.choice([0.0, 0.1, 1.0, 10.0, 1000.0]))
        nest = int(rng.choice([m + k + 1, max(2 * k + 2, m // 2)]))
        ipar = int(trial % 4 == 1)
        kw = {}
        if ipar:
            u = np.cumsum(rng.uniform(0.1, 1.0, m))
            kw = dict(u=u, ub=float(u[0]) - (trial % 2) * 0.1, ue=float(u[-1]))
        o = O.parcur(0, ipar, idim, x, w, k, s, nest, **kw)
        g = F.parcur(0, ipar, idim, x, w, k, s, nest, **kw)
        _same_parcur(o, g, idim, k)
        if o["ier"] > 0 or s == 0.0 or trial % 2:
            continue
        s2 = s * 0.4
        o1 = O.parcur(1, ipar, idim, x, w, k, s2, nest, u=o["u"], ub=o["ub"], ue=o["ue"], n=o["n"], t=o["t"],
                      wrk=o["wrk"], iwrk=o["iwrk"], c=o["c"])
        g1 = F.parcur(1, ipar, idim, x, w, k, s2, nest, u=g["u"], ub=g["ub"], ue=g["ue"], n=g["n"], t=g["t"],
                      wrk=g["wrk"], iwrk=g["iwrk"], c=g["c"])
        _same_parcur(o1, g1, idim, k)
        o2 = O.parcur(-1, ipar, idim, x, w, k, s2, nest, u=o1["u"], ub=o1["ub"], ue=o1["ue"], n=o1["n"], t=o1["t"])
        g2 = F.parcur(-1, ipar, idim, x, w, k, s2, nest, u=g1["u"], ub=g1["ub"], ue=g1["ue"], n=g1["n"], t=g1["t"])
        _same_parcur(o2, g2, idim, k)


def test_parcur_rejected_inputs_on_the_warp_per_curve_kernel(F):
    """parcur's checks (core:15255-15270) run inside the kernel: ier = 10, n / t / c / fp untouched, the computed
    parameters and the clamped end knots visible as in the reference"""
    m, k, idim, nest = 60, 3, 2, 70
    ph = np.linspace(0, 3, m)
    x = np.stack([np.cos(ph), np.sin(ph)], 1)
    u = np.linspace(0, 1, m)
    ubad = u.copy()
    ubad[10] = ubad[9]
    wbad = np.ones(m)
    wbad[5] = 0.0
    tbad = np.zeros(nest)
    tbad[4:8] = [0.2, 0.2, 0.5, 0.7]
    for a in (dict(ipar=1, u=ubad, ub=0.0, ue=1.0), dict(ipar=0, w=wbad),
              dict(ipar=1, iopt=-1, u=u, ub=0.0, ue=1.0, n=12, t=tbad), dict(ipar=1, u=u, ub=0.5, ue=1.0)):
        iopt, ipar, w = a.pop("iopt", 0), a.pop("ipar"), a.pop("w", None)
        o = O.parcur(iopt, ipar, idim, x, w, k, 0.1, nest, **a)
        g = F.parcur(iopt, ipar, idim, x, w, k, 0.1, nest, **a)
        assert o["ier"] == g["ier"] == 10
        _same_parcur(o, g, idim, k)
        assert _same(o["t"], g["t"]) and _same(o["c"], g["c"])


@pytest.mark.parametrize("idim,k,weights,ipar", [(2, 3, False, 0), (3, 5, True, 1)])
def test_parcur_host_batches_also_pass_on_the_batch_kernels(F, eng, idim, k, weights, ipar):
    """fp_parcur_batch_c slices of up to 4096 curves (idim 2, 3) run one warp per curve by default -- which is what
    test_gpu_parity.test_parcur_batch_bit_exact_vs_oracle (1500 curves) now exercises; here the same test on the batch
    kernels, and the rejected-curve test on both"""
    with single_curve_kernel(F, False):
        test_gpu_parity.test_parcur_batch_bit_exact_vs_oracle(F, eng, idim, k, weights, ipar)
    for on in (True, False):
        with single_curve_kernel(F, on):
            test_gpu_parity.test_parcur_percur_batches_with_invalid_curves(F)
            import test_gpu_boundary
            test_gpu_boundary.test_parcur_batch_rejected_rows_keep_the_callers_n_t_c(F)


def test_parcur_small_batch_continuation_and_given_knots(F):
    """iopt = 1 and iopt = -1 for a small batch of parametric curves through fp_parcur_batch_c"""
    from refdata import synth_param_curves
    rng = np.random.default_rng(314)
    B, m, idim, k = 200, 80, 2, 3
    th, x = synth_param_curves(rng, B, m, idim)
    nest, s = m + k + 1, m * 0.0025 * idim
    o = O.parcur_batch(0, 0, idim, x, None, k, s, nest)
    g = F.parcur_batch(0, 0, idim, x, None, k, s, nest)
    assert _same(g["n"], o["n"]) and _same(g["fp"], o["fp"]) and _same(g["u"], o["u"])
    # least squares on the knots found, per curve (the oracle one curve at a time)
    g2 = F.parcur_batch(-1, 0, idim, x, None, k, s, nest, u=g["u"], ub=g["ub"], ue=g["ue"], n=g["n"], t=g["t"])
    g1 = F.parcur_batch(1, 0, idim, x, None, k, s * 0.3, nest, u=g["u"], ub=g["ub"], ue=g["ue"], n=g["n"], t=g["t"],
                        fpint=g["fpint"], nrdata=g["nrdata"], c=g["c"])
    for b in range(0, B, 7):
        ob = O.parcur(0, 0, idim, x[b], None, k, s, nest)
        o2 = O.parcur(-1, 0, idim, x[b], None, k, s, nest, u=ob["u"], ub=ob["ub"], ue=ob["ue"], n=ob["n"], t=ob["t"])
        o1 = O.parcur(1, 0, idim, x[b], None, k, s * 0.3, nest, u=ob["u"], ub=ob["ub"], ue=ob["ue"], n=ob["n"], t=ob["t"],
                      wrk=ob["wrk"], iwrk=ob["iwrk"], c=ob["c"])
        for gg, oo in ((g2, o2), (g1, o1)):
            n = int(oo["n"])
            assert int(gg["n"][b]) == n and int(gg["ier"][b]) == oo["ier"] and gg["fp"][b] == oo["fp"], b
            assert _same(gg["t"][b, :n], oo["t"][:n])
            for d in range(idim):
                assert _same(gg["c"][b, d * n:d * n + n - k - 1], oo["c"][d * n:d * n + n - k - 1])


def test_percur_on_the_curve_per_cta_kernel(F):
    """fp_percur_c and fp_percur_batch_c slices of up to 1024 curves: fpperi (core:9668-10192) per curve in one launch,
    the curve's workspace in shared memory (fpb_percur_solo.cu: the per-curve code of the batch kernels compiled with
    unit stride).  Random fits with continuations, and a small batch, bit for bit against the oracle."""
    rng = np.random.default_rng(31)
    for trial in range(60):
        k = int(rng.integers(1, 6))
        m = int(rng.integers(2 * k + 4, 200))
        x = np.sort(rng.uniform(0, 2 * np.pi, m))
        x[0], x[-1] = 0.0, 2 * np.pi
        y = np.sin(x * int(rng.integers(1, 5))) + 0.05 * rng.standard_normal(m)
        y[-1] = y[0]
        w = rng.uniform(0.5, 2.0, m) if trial % 3 else None
        s = float(m * 0.0025 * rng.choice([0.0, 0.1, 1.0, 10.0, 1000.0]))
        nest = int(rng.choice([m + 2 * k, max(2 * k + 4, m // 2)]))
        o = O.percur(0, x, y, w, k, s, nest)
        g = F.percur(0, x, y, w, k, s, nest)
        _check(o, g, k)
        if o["ier"] > 0 or s == 0.0 or trial % 2:
            continue
        o1 = O.percur(1, x, y, w, k, s * 0.4, nest, n=o["n"], t=o["t"], wrk=o["wrk"], iwrk=o["iwrk"], c=o["c"], fp=o["fp"])
        g1 = F.percur(1, x, y, w, k, s * 0.4, nest, n=g["n"], t=g["t"], wrk=g["wrk"], iwrk=g["iwrk"], c=g["c"], fp=g["fp"])
        _check(o1, g1, k)
    B, m, k, nest = 300, 90, 3, 96
    th = np.sort(rng.uniform(0, 2 * np.pi, (B, m)), axis=1)
    th[:, 0], th[:, -1] = 0.0, 2 * np.pi
    yy = np.sin(2 * th) + 0.1 * rng.standard_normal(th.shape)
    yy[:, -1] = yy[:, 0]
    th[5, 10] = th[5, 9]                                  # curve 5: x not increasing -> ier = 10
    ob = O.percur_batch(0, th, yy, None, k, 0.9, nest)
    for on in (True, False):
        with single_curve_kernel(F, on):
            gb = F.percur_batch(0, th, yy, None, k, 0.9, nest)
            assert _same(gb["n"], ob["n"]) and _same(gb["ier"], ob["ier"]) and gb["ier"][5] == 10
            for b in range(B):
                if b == 5:
                    continue
                n = int(ob["n"][b])
                assert gb["fp"][b] == ob["fp"][b], b
                assert _same(gb["t"][b, :n], ob["t"][b, :n]) and _same(gb["c"][b, :n - k - 1], ob["c"][b, :n - k - 1]), b
