"""sproot (zeros of a cubic spline, reference core:17962-18109, fpcuro 5108-5198): the oracle against scipy's
translation of the same netlib routine, bit for bit (same libm), and against the defining property s(zero) = 0.
CPU only; fp_sproot_c is compared with the oracle in test_gpu_sproot.py."""
import numpy as np
from scipy import interpolate

import oracle_lib as O


def test_sproot_equals_scipy_netlib_bit_for_bit():
    rng = np.random.default_rng(4)
    for trial in range(120):
        m = int(rng.integers(12, 80))
        x = np.sort(rng.uniform(0, 10, m))
        y = np.sin(x * rng.uniform(0.5, 3)) + 0.3 * rng.standard_normal(m) * (trial % 3 == 0)
        t, c, _ = interpolate.splrep(x, y, k=3, s=float(rng.choice([0.05, 1.0, 5.0])))
        n = len(t)
        zo, mo, ier = O.sproot(t, n, c, 4 * n)
        zs = interpolate.sproot((t, c, 3), mest=4 * n)
        assert ier == 0 and mo == len(zs) and np.array_equal(zo, zs), trial
        if mo:
            assert np.max(np.abs(interpolate.splev(zo, (t, c, 3)))) < 1e-9


def test_sproot_checks_and_storage():
    x = np.linspace(0, 10, 60)
    t, c, _ = interpolate.splrep(x, np.sin(3 * x), k=3, s=0)
    n = len(t)
    z, m, ier = O.sproot(t, n, c, 3)
    assert ier == 1 and m == 3                            # more zeros than mest: the first mest found
    assert O.sproot(t[:7], 7, c[:7], 10)[2] == 13         # n < 8
    tb = t.copy()
    tb[6] = tb[5]
    assert O.sproot(tb, n, c, 40)[2] == 10
