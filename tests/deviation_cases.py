"""Directed cases for the places where the reference (perazz/fitpack) deliberately departs from netlib
FITPACK -- exactly the places no scipy golden can pin (scipy wraps netlib).  Each case comes with the
result DERIVED BY HAND from the reference source (cited), so that the oracle's reference mode and the
CUDA path are both checked against something that is neither of them.

Used by tests/test_deviations.py (oracle, CPU) and tests/test_gpu_deviations.py (CUDA path).
"""
import numpy as np


# ---------------------------------------------------------------------------------------------------
# (1) fpknot seeding, core:8218-8240.  `if (jpoint/=0 .and. (number==0 .or. fpint(j)>fpmax))`: the FIRST
# interval that holds an interior data point seeds the search and is only displaced by a strictly larger
# residual.  With a NaN ordinate every residual is NaN, `fpint(j) > fpmax` is false for all j, so the
# reference always splits the FIRST eligible interval (netlib's `fpmax >= fpint(j) -> skip` is false as
# well and ends on the LAST one).  The knot sequence then follows from the nrdata bookkeeping alone:
#   split interval `number` holding maxpt interior points at data point nrx = maxbeg + maxpt/2 + 1,
#   nrdata(number) = ihalf-1, nrdata(number+1) = maxpt-ihalf        (core:8243-8270)
# and from the nplus rule (core:4927-4935): first pass nplus = 1 (ier = -2), afterwards
# `fpold-fp > acc` is false for NaN, so npl1 = 2*nplus and nplus doubles; the loop leaves as soon as
# n == nest (core:4993), and the next pass returns ier = 1 (core:4916-4918).
def nan_case():
    m, k, nest = 12, 3, 12
    x = np.arange(m, dtype=np.float64)
    y = np.sin(x)
    y[5] = np.nan
    return dict(x=x, y=y, k=k, s=0.5, nest=nest, xb=0.0, xe=11.0)


def nan_case_expected(m=12, k=3, nest=12):
    """hand simulation of the bookkeeping above; returns (n, interior knots as data-point indices (1-based))"""
    nmin = 2 * (k + 1)
    nrdata = [m - 2]          # one interval holding all interior points (core:4838-4842)
    knots_at = []             # 1-based indices of the data points that became knots, by interval order
    n, nplus, first = nmin, 0, True
    while n < nest:
        nplus = 1 if first else nplus * 2
        first = False
        for _ in range(nplus):
            number = next((j for j, c in enumerate(nrdata) if c != 0), None)
            if number is None:
                break
            maxpt = nrdata[number]
            maxbeg = 1 + sum(c + 1 for c in nrdata[:number])
            ihalf = maxpt // 2 + 1
            nrx = maxbeg + ihalf
            nrdata[number:number + 1] = [ihalf - 1, maxpt - ihalf]
            knots_at.insert(number, nrx)
            n += 1
            if n == nest:
                break
    return n, knots_at


# ---------------------------------------------------------------------------------------------------
# (2) zero-pivot test `equal(piv, zero)`, core:5706 (fp_rotate_row) / 5810 (fp_rotate_shifted), with
# equal() = abs(a-b) < spacing(...) (core:18644-18647): a SUBNORMAL pivot counts as zero and the rotation
# is skipped (netlib tests piv == 0 and rotates).  A data point whose weight is subnormal therefore
# contributes exactly what a zero-weight point contributes: h = w*b is subnormal -> every pivot skipped,
# yi**2 and (w*(s(x)-y))**2 underflow to 0.  With the first interior knot between x(1) and x(2) the first
# B-spline has no other support point: its diagonal stays 0 and fpback (core:1808-1838) returns
# c(1) = 0/0 = NaN for BOTH weights -- netlib's rotation would instead leave a(1,1) = w*1 and give c(1) = y(1).
def subnormal_case(wfirst):
    m, k = 12, 3
    x = np.linspace(0.0, 1.0, m)
    y = np.cos(3.0 * x) + 2.0
    w = np.ones(m)
    w[0] = wfirst
    nest, n = 20, 12
    t = np.zeros(nest)
    t[4:8] = [0.05, 0.3, 0.6, 0.8]
    return dict(x=x, y=y, w=w, k=k, nest=nest, n=n, t=t, xb=0.0, xe=1.0)


# ---------------------------------------------------------------------------------------------------
# (3) `main_loop: do while (iter<=m)`, core:4847-4850 with iter starting at 0 and incremented at the top:
# up to m+1 knot sets (netlib: do iter=1,m).  It only shows when fpknot cannot add a knot: every nrdata(j)
# is 0, `if (number==0) return` (core:8239) leaves n as it is, and the loop spins.  An iopt=1 continuation
# whose iwrk(1:nrint) the caller zeroed does exactly that: m+1 identical least-squares passes, then part 2
# (core:5004-5084) on an unreachable target (f(p) >= fp_lsq > s), which ends with ier = 2 or 3; n and the
# knots are unchanged.  (netlib reads uninitialised memory here -- the crash the reference fixed.)
def spin_case():
    m, k = 16, 3
    rng = np.random.default_rng(3)
    x = np.linspace(0.0, 1.0, m)
    y = np.sin(7.0 * x) + 0.1 * rng.standard_normal(m)
    return dict(x=x, y=y, k=k, s_first=0.2, s_second=0.01, nest=m + k + 1, xb=0.0, xe=1.0, passes=m + 1)


# ---------------------------------------------------------------------------------------------------
# (4) fpchec Schoenberg-Whitney walk, core:2555-2564: `do while (i<m .and. x(i)<=tj)` TESTS x(i) before
# advancing, so one data point may serve two consecutive j (netlib advances first: `i=i+1` then tests).
# k=1, x = (0,5,10,10), t = (0,0,4,6,10,10): j=2 (tj=0, tl=6): x(1)=0<=0 -> i=2, x(2)=5>0, 5<6 ok;
# j=3 (tj=4, tl=10): x(2)=5>4 no advance, 5<10 ok -> accepted (0).  netlib: j=3 advances to x(3)=10>=tl -> 10.
def fpchec_case():
    return dict(x=np.array([0.0, 5.0, 10.0, 10.0]), t=np.array([0.0, 0.0, 4.0, 6.0, 10.0, 10.0]), n=6, k=1,
                expected=0)


# ---------------------------------------------------------------------------------------------------
# (5) parcur chord lengths through NORM2, core:15243: gfortran's NORM2 scales by the running maximum, so
# coordinate differences of 3e200 / 4e200 give 5e200 where sqrt(sum d**2) (netlib) overflows to Inf.
# Six collinear points at equal spacing => u = (0, .2, .4, .6, .8, 1) (each chord = 5e200 exactly? no:
# the quotients acc/total are the correctly rounded i/5 only if every partial sum is exact; the test
# asserts 1e-15 and finiteness, and bit-equality between oracle and GPU).
def norm2_case():
    X = np.array([[0, 0], [3e200, 4e200], [6e200, 8e200], [9e200, 1.2e201], [1.2e201, 1.6e201],
                  [1.5e201, 2e201]], dtype=np.float64)
    return dict(x=X, idim=2, k=1, s=0.0, nest=20, u=np.linspace(0.0, 1.0, 6))


# ---------------------------------------------------------------------------------------------------
# (6) regrid's axis arbiter new_knot_dimension_nd, core:12347-12362: the first knot and every tie go to the
# LAST axis (netlib's fpregr alternates starting with x).  On data symmetric in (x,y) over identical grids
# the two axes see identical residual sums at every step, so every decision is a tie: the reference can only
# produce ny >= nx and ny - nx <= 1... (each tie adds to y, the next step the sums differ by the new knot);
# what is derived by hand and asserted: ny >= nx, and netlib mode gives the transposed pair.
def regrid_symmetric_case():
    g = np.linspace(0.0, 1.0, 15)
    GX, GY = np.meshgrid(g, g, indexing="ij")
    z = np.sin(3 * GX) * np.sin(3 * GY) + GX * GY
    return dict(g=g, z=z, k=3, nest=19, s_values=(0.05, 0.005, 0.0005))
