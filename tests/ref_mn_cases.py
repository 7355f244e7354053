"""Two of the reference's own mn* test programs (test/fitpack_tests.f90), re-enacted against any implementation `I` that
offers insert / splint / sproot / curev with the oracle_lib signatures (the CPU oracle, or fitpack_b200 on the GPU).

mnspin (:4100-4186): the B-spline form of (1-x)^k, k = 1..5, four knots inserted with insert_inplace, then splint over
four ranges against the exact integral ((1-a)^(k+1) - (1-b)^(k+1)) / (k+1).  The reference asserts 1e-3 relative; the
spline IS the polynomial, so 1e-13 holds and is what is asserted here.
mnspro (:4190-4320): a closed planar cubic spline curve cut by five straight lines alfa x + beta y = gamma: sproot of
alfa sx + beta sy - gamma, then curev at the zeros.  The reference asserts only ier; here also that every
intersection point lies on its line (1e-12) and that the zeros are strictly increasing inside the parameter range."""
import numpy as np


def run_mnspin(I):
    worst = 0.0
    for k in range(1, 6):
        k1 = k + 1
        n, nest = 2 * k1, 20
        t, c = np.zeros(nest), np.zeros(nest)
        c[0] = 1.0
        t[n - k - 1:n] = 1.0
        for xk in (0.8, 0.4, 0.3, 0.1):
            t, n, c, ier = I.insert(0, t, n, c, k, xk, nest)
            assert ier == 0
        assert n == 2 * k1 + 4
        a, b = 0.0, 1.0
        for _ in range(4):
            aint = I.splint(t, n, c, k, a, b)[0]
            exint = ((1.0 - a) ** k1 - (1.0 - b) ** k1) / k1
            assert abs(aint - exint) < 1e-3 * abs(exint)            # the reference's own assertion
            assert abs(aint - exint) < 1e-13 * max(1.0, abs(exint)), (k, a, b)
            worst = max(worst, abs(aint - exint))
            a, b = a + 0.1, b - 0.3
    return worst


def mnspro_curve():
    n = 13
    t = np.zeros(n)
    t[3:10] = [0.0, 0.2, 0.3, 0.5, 0.6, 0.7, 1.0]
    c = np.zeros(26)
    c[0:6] = [1.0, 3.0, 4.0, 5.0, 6.0, -1.0]
    c[13:19] = [1.0, 2.0, -3.0, 2.0, 1.0, 4.0]
    per = t[9] - t[3]
    t[0:3] = t[6:9] - per
    t[10:13] = t[4:7] + per
    c[6:9] = c[0:3]
    c[19:22] = c[13:16]
    return t, n, c


def run_mnspro(I):
    t, n, c = mnspro_curve()
    nk1 = n - 4
    lines = [(0.0, 1.0, 0.0), (1.0, 0.0, 0.0), (1.0, -1.0, 0.0), (0.4, 0.3, 1.2), (0.4, 0.4, 0.0)]
    counts = []
    for alfa, beta, gamma in lines:
        cc = np.zeros(n)
        cc[:nk1] = alfa * c[:nk1] + beta * c[n:n + nk1] - gamma
        z, m, ier = I.sproot(t, n, cc, 20)
        assert ier <= 0
        counts.append(m)
        if m:
            assert np.all(np.diff(z) > 0) and z[0] >= t[3] and z[-1] <= t[9]
            xy, ie = I.curev(2, t, n, c, 3, z)
            assert ie == 0
            xy = np.asarray(xy).reshape(m, 2)
            assert np.max(np.abs(alfa * xy[:, 0] + beta * xy[:, 1] - gamma)) < 1e-12
    return counts
