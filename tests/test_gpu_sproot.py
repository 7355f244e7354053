"""fp_sproot_c on the GPU against the CPU oracle (reference core:17962-18109).  The kernel follows the reference's
operation order; its pow / atan2 / cos are CUDA's, so the zeros agree to rounding level: tolerance 1e-12 * max(1, |z|),
same count on well-separated zeros."""
import numpy as np
import pytest
from scipy import interpolate

import oracle_lib as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def F():
    import fitpack_b200
    return fitpack_b200


def test_sproot_against_the_oracle(F):
    rng = np.random.default_rng(4)
    for trial in range(60):
        m = int(rng.integers(12, 80))
        x = np.sort(rng.uniform(0, 10, m))
        y = np.sin(x * rng.uniform(0.5, 3)) + 0.3 * rng.standard_normal(m) * (trial % 3 == 0)
        t, c, _ = interpolate.splrep(x, y, k=3, s=float(rng.choice([0.05, 1.0, 5.0])))
        n = len(t)
        zo, mo, io = O.sproot(t, n, c, 4 * n)
        zg, mg, ig = F.sproot(t, n, c, 4 * n)
        assert (mg, ig) == (mo, io), trial
        assert np.all(np.abs(zg - zo) <= 1e-12 * np.maximum(1.0, np.abs(zo))), trial


def test_sproot_checks_and_storage(F):
    x = np.linspace(0, 10, 60)
    t, c, _ = interpolate.splrep(x, np.sin(3 * x), k=3, s=0)
    n = len(t)
    z, m, ier = F.sproot(t, n, c, 3)
    assert ier == 1 and m == 3
    assert F.sproot(t[:7], 7, c[:7], 10)[2] == 13
    tb = t.copy()
    tb[6] = tb[5]
    assert F.sproot(tb, n, c, 40)[2] == 10
