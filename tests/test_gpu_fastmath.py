"""GPU: the hand-written division / square-root / fpgivs sequences of the fit kernels (fpb_common.cuh: the fast
paths of nvcc's own IEEE expansions with the reciprocal refinement hoisted out) give the bits of the IEEE
operators: 4e9 random operand sets per check, compared on the device."""
import ctypes as C

import pytest

pytestmark = pytest.mark.gpu


def test_fast_division_sqrt_and_givens_are_bit_identical_to_the_ieee_operators():
    import torch
    import fitpack_b200
    from fitpack_b200.batch import Engine
    eng = Engine(torch.device("cuda", 0))
    out = (C.c_int64 * 3)()
    total = 0
    for seed in (1, 2, 3, 4):
        rc = fitpack_b200.lib().fpb_selftest_fastmath(eng.handle, 1_000_000_000, seed, C.addressof(out))
        assert rc == 0
        assert list(out) == [0, 0, 0], (seed, list(out))
        total += 1_000_000_000
    assert total == 4_000_000_000
    eng.close()
