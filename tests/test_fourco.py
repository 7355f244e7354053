"""fourco (Fourier integrals of a cubic spline, reference core:1474-1520, fpbfou 1947-2123, fpcsin 4636-4686): the
oracle against the reference's own known-answer test mnfour (test/fitpack_tests.f90:1536-1680: y = x on [0, 1] as a
cubic spline, alfa = 0, 1e-3, -1e-2, ..., 1e3, exact integrals, tolerance 1e-10 + 1e-6 |exact|) and against numerical
quadrature on random cubic splines.  CPU only; the GPU entry point is compared with the oracle in test_gpu_fourco.py."""
import numpy as np
import pytest
from scipy import integrate, interpolate

import oracle_lib as O


def mnfour_spline():
    """knots and coefficients of y = x exactly as mnfour builds them (test/fitpack_tests.f90:1561-1580)"""
    k, k1 = 3, 4
    n = 2 * k1 + 4
    t = np.zeros(n)
    t[n - k - 1:] = 1.0
    t[4:8] = [0.1, 0.3, 0.4, 0.8]
    c = np.zeros(n)
    for i in range(2, n - k1 + 1):          # c(i) = c(i-1) + (t(i+k) - t(i)) / k, 1-based
        c[i - 1] = c[i - 2] + (t[i + k - 1] - t[i - 1]) / k
    return t, n, c


def exfour(alfa):
    """exact int_0^1 x sin(alfa x) dx, int_0^1 x cos(alfa x) dx (test/fitpack_tests.f90: exfour, series for small alfa)"""
    if abs(alfa) < 1e-3:
        return alfa / 3.0 - alfa ** 3 / 30.0, 0.5 - alfa ** 2 / 8.0
    s, co = np.sin(alfa), np.cos(alfa)
    return (s - alfa * co) / alfa ** 2, (co + alfa * s - 1.0) / alfa ** 2


def test_mnfour_known_answers():
    t, n, c = mnfour_spline()
    alfa = [0.0, 0.001]
    for _ in range(6):
        alfa.append(-10 * alfa[-1])
    ress, resc, ier, _, _ = O.fourco(t, n, c, np.array(alfa))
    assert ier == 0
    for a, rs, rc in zip(alfa, ress, resc):
        es, ec = exfour(a)
        assert abs(rs - es) < 1e-10 + 1e-6 * abs(es), a
        assert abs(rc - ec) < 1e-10 + 1e-6 * abs(ec), a


def test_against_quadrature_on_random_cubic_splines():
    rng = np.random.default_rng(3)
    for trial in range(12):
        nint = int(rng.integers(3, 20))
        interior = np.sort(rng.uniform(0.05, 0.95, nint - 1)) * 4.0
        t = np.concatenate([np.zeros(4), interior, np.full(4, 4.0)])
        n = t.size
        c = np.concatenate([rng.standard_normal(n - 4), np.zeros(4)])
        spl = interpolate.BSpline(t, c[:n - 4], 3)
        alfa = np.array([0.0, 0.05, 0.7, -2.5, 9.0, 40.0])
        ress, resc, ier, _, _ = O.fourco(t, n, c, alfa)
        assert ier == 0
        for a, rs, rc in zip(alfa, ress, resc):
            qs = integrate.quad(lambda x: spl(x) * np.sin(a * x), 0.0, 4.0, points=interior, limit=400, epsabs=1e-13, epsrel=1e-13)[0]
            qc = integrate.quad(lambda x: spl(x) * np.cos(a * x), 0.0, 4.0, points=interior, limit=400, epsabs=1e-13, epsrel=1e-13)[0]
            assert abs(rs - qs) < 2e-8 * max(1.0, abs(qs)), (trial, a)   # the series stop at 1e-8 (eps of fpbfou)
            assert abs(rc - qc) < 2e-8 * max(1.0, abs(qc)), (trial, a)


def test_input_checks():
    t, n, c = mnfour_spline()
    assert O.fourco(t[:9], 9, c[:9], np.array([1.0]))[2] == 13          # n < 10: FITPACK_INSUFFICIENT_KNOTS
    tb = t.copy()
    tb[5] = tb[4]                                                        # interior knots not strictly increasing
    assert O.fourco(tb, n, c, np.array([1.0]))[2] == 10
    tb = t.copy()
    tb[1] = 0.5                                                          # boundary knots decreasing
    assert O.fourco(tb, n, c, np.array([1.0]))[2] == 10
