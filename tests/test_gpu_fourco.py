"""fp_fourco_c / fitpack_curve_c_fourier on the GPU against the CPU oracle (reference core:1474-1520) and the reference's
known-answer test mnfour.  Tolerance, stated: 1e-12 absolute against the oracle on the reference's mnfour case, 5e-9 * max(1, |value|) on
random splines with close knots (the kernel follows the reference's operation order but its sin / cos are CUDA's,
<= 2 ulp, libm's in the reference, and the divided differences amplify that) -- and mnfour's own 1e-10 + 1e-6 |exact|
against the exact integrals."""
import numpy as np
import pytest

import oracle_lib as O
from test_fourco import exfour, mnfour_spline

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def F():
    import fitpack_b200
    return fitpack_b200


def test_mnfour_known_answers_on_the_gpu(F):
    t, n, c = mnfour_spline()
    alfa = [0.0, 0.001]
    for _ in range(6):
        alfa.append(-10 * alfa[-1])
    ress, resc, ier, w1, w2 = F.fourco(t, n, c, np.array(alfa))
    assert ier == 0
    for a, rs, rc in zip(alfa, ress, resc):
        es, ec = exfour(a)
        assert abs(rs - es) < 1e-10 + 1e-6 * abs(es) and abs(rc - ec) < 1e-10 + 1e-6 * abs(ec), a
    os_, oc_, _, ow1, ow2 = O.fourco(t, n, c, np.array(alfa))
    assert np.allclose(ress, os_, rtol=0, atol=1e-12) and np.allclose(resc, oc_, rtol=0, atol=1e-12)
    assert np.allclose(w1[:n - 4], ow1[:n - 4], rtol=0, atol=1e-12) and np.allclose(w2[:n - 4], ow2[:n - 4], rtol=0, atol=1e-12)


def test_random_cubic_splines_against_the_oracle(F):
    rng = np.random.default_rng(8)
    worst = 0.0
    for trial in range(40):
        nint = int(rng.integers(3, 60))
        span = float(rng.choice([1.0, 6.0, 50.0]))
        interior = np.sort(rng.uniform(0.02, 0.98, nint - 1)) * span
        t = np.concatenate([np.zeros(4), interior, np.full(4, span)])
        n = t.size
        c = np.concatenate([rng.standard_normal(n - 4), np.zeros(4)])
        alfa = np.concatenate([[0.0], rng.uniform(-30, 30, 70), 10.0 ** rng.uniform(-6, 3, 30)])
        ress, resc, ier, _, _ = F.fourco(t, n, c, alfa)
        os_, oc_, oier, _, _ = O.fourco(t, n, c, alfa)
        assert ier == oier == 0
        # the divided differences divide by (alfa * knot gap)^3: with close knots the 2-ulp difference between CUDA's and
        # libm's sin / cos is amplified (the algorithm's own accuracy is ~1e-8, the cut-off of its series)
        tol = 5e-9 * np.maximum(1.0, np.maximum(np.abs(os_), np.abs(oc_)))   # largest seen: 5.6e-10
        err = max(np.max(np.abs(ress - os_)), np.max(np.abs(resc - oc_)))
        worst = max(worst, err)
        assert np.all(np.abs(ress - os_) <= tol) and np.all(np.abs(resc - oc_) <= tol), (trial, err)
    print("largest |gpu - oracle| over all cases: %.3e" % worst)


def test_input_checks_and_handle(F):
    t, n, c = mnfour_spline()
    assert F.fourco(t[:9], 9, c[:9], np.array([1.0]))[2] == 13
    tb = t.copy()
    tb[5] = tb[4]
    assert F.fourco(tb, n, c, np.array([1.0]))[2] == 10
    # handle: an interpolating cubic curve through y = x reproduces mnfour's integrals
    x = np.linspace(0.0, 1.0, 41)
    cv = F.FitpackCurve()
    assert F.FITPACK_SUCCESS(cv.new_fit(x, x.copy(), smoothing=0.0))
    a, b, ierr = cv.fourier(np.array([0.0, 0.1, -1.0, 10.0]))
    assert ierr <= 0
    for al, rs, rc in zip([0.0, 0.1, -1.0, 10.0], a, b):
        es, ec = exfour(al)
        assert abs(rs - es) < 1e-9 and abs(rc - ec) < 1e-9, al
