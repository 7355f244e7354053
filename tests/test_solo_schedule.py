"""The wavefront schedules of the warp-per-curve kernel (fitpack_b200/csrc/fpb_curfit_solo.cu), replayed on the CPU.

tests/solo_sim/solo_sim.c runs the kernel's tick schedule -- which row meets which band row at which tick, G lane
groups, every tick reading what the previous tick left -- with the oracle's own fpgivs / fprota as arithmetic, and
asserts that no two groups touch the same band row in a tick and that a group is free before its next row starts.
Here its (R, z, fp) and (G, c) are compared bit for bit with the oracle's sequential fp_rotate_row /
fp_rotate_shifted sweeps (core:5691-5717, 5800-5820) on knot sets taken from real fits.  CPU only."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as O

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def sim(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("solo_sim") / "libsolo_sim.so")
    subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fopenmp", "-shared", "-fPIC", "-o", out,
                           os.path.join(HERE, "solo_sim", "solo_sim.c"), "-lm"])
    return C.CDLL(out)


def _dp(a):
    return a.ctypes.data_as(C.c_void_p)


def _cfg(k):
    """SoloCfg<K> of the kernel: groups and tick stride of the data rows, groups of the smoothing rows"""
    k1, k2 = k + 1, k + 2
    gd = min(32 // (k1 + 1), k1)
    return gd, (1 if gd >= k1 else 2), 32 // (k2 + 1)


def test_kernel_configurations_leave_every_group_free_in_time():
    for k in range(1, 6):
        gd, st, gs = _cfg(k)
        assert gd * st >= k + 1 and gd * (k + 2) <= 32 and gs * (k + 3) <= 32 and gs >= 1


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_wavefront_sweeps_equal_the_sequential_sweeps_bit_for_bit(sim, seed):
    rng = np.random.default_rng(seed)
    L = O.lib()
    L.orc_knot_interval.restype = C.c_int32
    for trial in range(60):
        m = int(rng.integers(12, 260))
        k = int(rng.integers(1, 6))
        x = np.sort(rng.uniform(0, 1, m))
        x[0], x[-1] = 0.0, 1.0
        y = np.sin(2 * np.pi * rng.uniform(0.5, 4) * x) + 0.05 * rng.standard_normal(m)
        w = rng.uniform(0.5, 2.0, m)
        if trial % 7 == 0:
            w[rng.integers(0, m)] = 0.0   # a zero weight: zero pivots, skipped rotations
        s = float(m * 0.0025 * rng.choice([0.0, 0.1, 1.0, 10.0]))
        nest = m + k + 1
        o = O.curfit(0, x, y, w, 0.0, 1.0, k, s, nest)
        n = int(o["n"])
        t = np.ascontiguousarray(o["t"][:nest]).copy()
        k1, k2 = k + 1, k + 2
        nk1 = n - k1
        lidx = np.zeros(m, np.int32)
        q = np.zeros((m, k1))
        h = np.zeros(8)
        l = k1
        for it in range(m):
            l = L.orc_knot_interval(_dp(t), C.c_double(x[it]), C.c_int32(l), C.c_int32(nk1))
            lidx[it] = l
            L.orc_fpbspl(_dp(t), C.c_int32(n), C.c_int32(k), C.c_double(x[it]), C.c_int32(l), _dp(h))
            q[it] = h[:k1]
        a0, z0, fp0 = np.zeros(nest * k2), np.zeros(nest), C.c_double()
        sim.sim_ref_lsq_sweep(m, k, nk1, nest, _dp(lidx), _dp(q), _dp(y), _dp(w), _dp(a0), _dp(z0), C.byref(fp0))
        gd, st, gs = _cfg(k)
        a1, z1, fp1, ticks = np.zeros(nest * k2), np.zeros(nest), C.c_double(), C.c_int32()
        sim.sim_lsq_sweep_static(m, k, nk1, nest, _dp(lidx), _dp(q), _dp(y), _dp(w), gd, st, _dp(a1), _dp(z1),
                                 C.byref(fp1), C.byref(ticks))
        assert np.array_equal(a0, a1) and np.array_equal(z0, z1) and fp0.value == fp1.value, (trial, k)
        assert ticks.value <= st * m + nk1          # m + n ticks instead of m (k + 1) rotations
        for G in (1, 3, 8):                          # the dependency-checked form, any number of groups
            a2, z2, fp2 = np.zeros(nest * k2), np.zeros(nest), C.c_double()
            sim.sim_lsq_sweep(m, k, nk1, nest, _dp(lidx), _dp(q), _dp(y), _dp(w), G, _dp(a2), _dp(z2),
                              C.byref(fp2), C.byref(ticks))
            assert np.array_equal(a0, a2) and np.array_equal(z0, z2) and fp0.value == fp2.value, (trial, k, G)
        if n > 2 * k1:
            b = np.zeros(nest * k2)
            sim.sim_fpdisc(_dp(t), n, k2, _dp(b), nest)
            pinv = C.c_double(1.0 / rng.uniform(0.01, 100))
            g0, c0 = a0.copy(), z0.copy()
            sim.sim_ref_smooth_sweep(n, k, nest, _dp(b), pinv, _dp(g0), _dp(c0))
            g1, c1 = a0.copy(), z0.copy()
            sim.sim_smooth_sweep_static(n, k, nest, _dp(b), pinv, gs, _dp(g1), _dp(c1), C.byref(ticks))
            assert np.array_equal(g0, g1) and np.array_equal(c0, c1), (trial, k, "smoothing rows")
            for G in (1, 5):
                g2, c2 = a0.copy(), z0.copy()
                sim.sim_smooth_sweep(n, k, nest, _dp(b), pinv, G, _dp(g2), _dp(c2), C.byref(ticks))
                assert np.array_equal(g0, g2) and np.array_equal(c0, c2), (trial, k, G)
