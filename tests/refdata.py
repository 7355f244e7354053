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

# regrid test data `daregr` (test/fitpack_test_data.f90:175-192): 11 x 11 grid on [-1.5, 1.5]^2;
# flat list order == z(my*(i-1)+j) as mnregr passes it (test/fitpack_tests.f90:3170-3362)
DAREGR_X = np.array([-1.5, -1.2, -0.9, -0.6, -0.3, 0.0, 0.3, 0.6, 0.9, 1.2, 1.5])
DAREGR_Y = DAREGR_X.copy()
DAREGR_Z = np.array([
    -0.0325, 0.0784, 0.0432, 0.0092, 0.1523, 0.0802, 0.0925, -0.0098, 0.0810, -0.0146, -0.0019,
    0.1276, 0.0223, 0.0357, 0.1858, 0.2818, 0.1675, 0.2239, 0.1671, 0.0843, 0.0151, 0.0427,
    0.0860, 0.1267, 0.1839, 0.3010, 0.5002, 0.4683, 0.4562, 0.2688, 0.1276, 0.1244, 0.0377,
    0.0802, 0.1803, 0.3055, 0.4403, 0.6116, 0.7178, 0.6797, 0.5218, 0.2624, 0.1341, -0.0233,
    0.1321, 0.2023, 0.4446, 0.7123, 0.7944, 0.9871, 0.8430, 0.6440, 0.4682, 0.1319, 0.1075,
    0.2561, 0.1900, 0.4614, 0.7322, 0.9777, 1.0463, 0.9481, 0.6649, 0.4491, 0.2442, 0.1341,
    0.0981, 0.2009, 0.4616, 0.5514, 0.7692, 0.9831, 0.7972, 0.5937, 0.4190, 0.1436, 0.0995,
    0.0991, 0.1545, 0.3399, 0.4940, 0.6328, 0.7168, 0.6886, 0.3925, 0.3015, 0.1758, 0.0928,
    -0.0197, 0.1479, 0.1225, 0.3254, 0.3847, 0.4767, 0.4324, 0.2827, 0.2287, 0.0999, 0.0785,
    0.0032, 0.0917, 0.0246, 0.1780, 0.2394, 0.1765, 0.1642, 0.2081, 0.1049, 0.0493, -0.0502,
    0.0101, 0.0297, 0.0468, 0.0221, 0.1074, 0.0433, 0.0626, 0.1436, 0.1092, -0.0232, 0.0132])
# mnregr call sequence: (iopt, kx, ky, s); the last one is iopt=-1 with interior knots -0.5, 0, 0.5
MNREGR_STEPS = [(0, 3, 3, 10.0), (1, 3, 3, 0.22), (1, 3, 3, 0.1), (1, 3, 3, 0.0), (0, 5, 5, 0.2),
                (-1, 3, 3, 0.0)]


def synth_param_curves(rng, B, m, idim, sigma=0.05):
    """Noisy Lissajous-like parametric curves, [B][m][idim] (memory order of Fortran x(idim,m))."""
    th = np.sort(rng.uniform(0, 2 * np.pi, (B, m)), axis=1)
    th += np.arange(m)[None, :] * 1e-9  # strictly increasing
    cols = []
    for d in range(idim):
        f = rng.uniform(0.5, 3, (B, 1))
        cols.append(rng.uniform(.5, 2, (B, 1)) * np.cos(f * th + rng.uniform(0, 6, (B, 1))))
    x = np.stack(cols, 2) + sigma * rng.standard_normal((B, m, idim))
    return np.ascontiguousarray(th), np.ascontiguousarray(x)
