"""Datasets the reference's own tests use for this path (values restated, not code)."""
import numpy as np

# test/fitpack_tests.f90:1155-1158 (mncurf)
MNCURF_X = np.arange(25.0)
MNCURF_Y = np.array([1.0, 1.0, 1.4, 1.1, 1.0, 1.0, 4.0, 9.0, 13.0, 13.4, 12.8, 13.1, 13.0, 14.0,
                     13.0, 13.5, 10.0, 2.0, 3.0, 2.5, 2.5, 2.5, 3.0, 4.0, 3.5])

# test/fitpack_curve_tests.f90:61-85 (test_fpknot_crash, scipy issue 3691)
FPKNOT_X = np.arange(109.0)
FPKNOT_Y = np.array([
    0., 0., 0., 0., 0., 10.9, 0., 11., 0.,
    0., 0., 10.9, 0., 0., 0., 0., 0., 0.,
    10.9, 0., 0., 0., 11., 0., 0., 0., 10.9,
    0., 0., 0., 10.5, 0., 0., 0., 10.7, 0.,
    0., 0., 11., 0., 0., 0., 0., 0., 0.,
    10.9, 0., 0., 10.7, 0., 0., 0., 10.6, 0.,
    0., 0., 10.5, 0., 0., 10.7, 0., 0., 10.5,
    0., 0., 11.5, 0., 0., 0., 10.7, 0., 0.,
    10.7, 0., 0., 10.9, 0., 0., 10.8, 0., 0.,
    0., 10.7, 0., 0., 10.6, 0., 0., 0., 10.4,
    0., 0., 10.6, 0., 0., 10.5, 0., 0., 0.,
    10.7, 0., 0., 0., 10.4, 0., 0., 0., 10.8, 0.])

# mncurf call sequence (iopt, s); step 7 is the iopt=-1 fit with knots 3,6,..,21
MNCURF_STEPS = [(0, 1000.0), (1, 60.0), (1, 10.0), (1, 30.0), (0, 30.0), (0, 0.0), (-1, 0.0)]


def synth_curves(rng, B, m, sigma=0.1):
    """C2-style synthetic batch."""
    u = rng.uniform(-1, 1, (B, m))
    x = (np.arange(m)[None, :] + 0.3 * u) / (m - 1)
    x[:, 0] = 0.0
    x[:, -1] = 1.0
    A = rng.uniform(0.5, 2, (B, 1))
    f = rng.uniform(1, 4, (B, 1))
    ph = rng.uniform(0, 2 * np.pi, (B, 1))
    y = A * np.sin(2 * np.pi * f * x + ph) + sigma * rng.standard_normal((B, m))
    return np.ascontiguousarray(x), np.ascontiguousarray(y)
