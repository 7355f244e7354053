"""insert (knot insertion, reference core:15064-15159 + fpinst 7961-8026) and cualde (all derivatives of a parametric
curve, core:1062-1101): the oracle against scipy's translation of the same netlib routines -- bit for bit for insert
(open and periodic splines), against spalde per coordinate for cualde -- and against the defining property (the spline
does not change).  CPU only; the GPU entry points are compared with the oracle in test_gpu_insert_cualde.py."""
import numpy as np
from scipy import interpolate

import oracle_lib as O


def test_insert_equals_scipy_netlib_bit_for_bit():
    rng = np.random.default_rng(1)
    for k in (1, 2, 3, 4, 5):
        for per in (0, 1):
            x = np.sort(rng.uniform(0, 1, 50))
            x[0], x[-1] = 0.0, 1.0
            y = np.sin(2 * np.pi * x) + 0.02 * rng.standard_normal(50)
            if per:
                y[-1] = y[0]
            tck = interpolate.splrep(x, y, k=k, s=0.02, per=per)
            t, c, _ = tck
            n = len(t)
            for xx in (0.013, 0.37, 0.5, 0.981, float(t[k + 2]) if n > 2 * k + 4 else 0.4):
                tt, nn, cc, ier = O.insert(per, t, n, c, k, xx, n + 4)
                try:
                    t2, c2, _ = interpolate.insert(xx, tck, per=per)
                except ValueError:                       # netlib's insert returned ier = 10
                    assert ier == 10 and nn == n, (k, per, xx)
                    continue
                assert ier == 0 and nn == len(t2), (k, per, xx)
                assert np.array_equal(tt[:nn], t2) and np.array_equal(cc[:nn - k - 1], c2[:nn - k - 1]), (k, per, xx)
                xe = np.linspace(0, 1, 200)
                assert np.allclose(interpolate.splev(xe, (tt[:nn], cc[:nn], k)), interpolate.splev(xe, tck), atol=1e-12)


def test_insert_rejections_leave_everything_untouched():
    t = np.array([0, 0, 0, 0, 0.5, 1, 1, 1, 1.0])
    c = np.array([1.0, 2, 3, 4, 5, 0, 0, 0, 0])
    for args in (dict(x=0.3, nest=9), dict(x=-0.1, nest=12), dict(x=1.5, nest=12)):
        tt, nn, cc, ier = O.insert(0, t, 9, c, 3, args["x"], args["nest"])
        assert ier == 10 and nn == 9 and np.array_equal(tt[:9], t) and np.array_equal(cc[:9], c)


def test_cualde_is_spalde_per_coordinate_interleaved():
    rng = np.random.default_rng(2)
    for idim in (1, 2, 3):
        for k in (1, 3, 5):
            u = np.linspace(0, 1, 60)
            x = np.stack([np.cos(3 * u), np.sin(3 * u), u ** 2][:idim], 1) + 0.01 * rng.standard_normal((60, idim))
            o = O.parcur(0, 1, idim, x, None, k, 0.01, 70, u=u, ub=0.0, ue=1.0)
            n = o["n"]
            for uu in (0.0, 0.21, 0.5, 1.0):
                d, ier = O.cualde(idim, o["t"], n, o["c"], k + 1, uu)
                assert ier == 0
                for i in range(idim):
                    di, iei = O.spalde(o["t"], n, o["c"][i * n:(i + 1) * n], k + 1, uu)
                    assert iei == 0 and np.array_equal(d[:, i], di[:k + 1])
            assert O.cualde(idim, o["t"], n, o["c"], k + 1, 1.5)[1] == 10
