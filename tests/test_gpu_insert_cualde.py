"""fp_insert_c and fp_cualde_c on the GPU against the CPU oracle (reference core:15064-15159 / 7961-8026 and 1062-1101),
bit for bit, on open and periodic splines, every degree, accepted and rejected arguments."""
import numpy as np
import pytest
from scipy import interpolate

import oracle_lib as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def F():
    import fitpack_b200
    return fitpack_b200


def test_insert_bit_exact(F):
    rng = np.random.default_rng(1)
    for k in (1, 2, 3, 4, 5):
        for per in (0, 1):
            x = np.sort(rng.uniform(0, 1, 50))
            x[0], x[-1] = 0.0, 1.0
            y = np.sin(2 * np.pi * x) + 0.02 * rng.standard_normal(50)
            if per:
                y[-1] = y[0]
            t, c, _ = interpolate.splrep(x, y, k=k, s=0.02, per=per)
            n = len(t)
            for xx in (0.013, 0.37, 0.5, 0.981, -0.2, 1.0, float(t[min(k + 2, n - 1)])):
                for nest in (n + 4, n):
                    to, no, co, io = O.insert(per, t, n, c, k, xx, nest)
                    tg, ng, cg, ig = F.insert(per, t, n, c, k, xx, nest)
                    assert (ng, ig) == (no, io), (k, per, xx, nest)
                    assert np.array_equal(tg[:ng], to[:no]) and np.array_equal(cg[:max(ng - k - 1, 0)], co[:max(no - k - 1, 0)])
            # several insertions in a row (the handle API's insert_knot(many) pattern)
            to, no, co = np.concatenate([t, np.zeros(8)]), n, np.concatenate([c, np.zeros(8)])
            tg, ng, cg = to.copy(), n, co.copy()
            for xx in (0.11, 0.11, 0.62, 0.9):
                to, no, co, io = O.insert(per, to, no, co, k, xx, n + 8)
                tg, ng, cg, ig = F.insert(per, tg, ng, cg, k, xx, n + 8)
                assert (ng, ig) == (no, io) and np.array_equal(tg[:ng], to[:no]) and np.array_equal(cg[:ng - k - 1], co[:no - k - 1])


def test_cualde_bit_exact(F):
    rng = np.random.default_rng(2)
    for idim in (1, 2, 3):
        for k in (1, 3, 5):
            u = np.linspace(0, 1, 60)
            x = np.stack([np.cos(3 * u), np.sin(3 * u), u ** 2][:idim], 1) + 0.01 * rng.standard_normal((60, idim))
            o = O.parcur(0, 1, idim, x, None, k, 0.01, 70, u=u, ub=0.0, ue=1.0)
            n = o["n"]
            for uu in (0.0, 0.21, 0.5, 1.0, 1.5, -0.1):
                do, io = O.cualde(idim, o["t"], n, o["c"], k + 1, uu)
                dg, ig = F.cualde(idim, o["t"], n, o["c"], k + 1, uu)
                assert ig == io, (idim, k, uu)
                if io == 0:
                    assert np.array_equal(dg, do), (idim, k, uu)
