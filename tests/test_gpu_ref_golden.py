"""GPU: the CUDA path against the golden vectors of the Fortran reference itself (tests/golden/ref_*.npz, see
tests/test_ref_golden.py) when they exist, and -- always -- against the oracle on the same replayed case list
and on 65 536 config-2 curves (SURVEY 8d asked for that sample size)."""
import os

import numpy as np
import pytest

import oracle_lib as O
from test_ref_golden import GOLD, MRG, _OracleBackend, compare_c2, compare_curfit

pytestmark = pytest.mark.gpu


class _GpuBackend:
    @staticmethod
    def curfit(*a, **kw):
        import fitpack_b200
        return fitpack_b200.curfit(*a, **kw)

    @staticmethod
    def curfit_batch(iopt, x, y, w, k, s, nest):
        import fitpack_b200
        return fitpack_b200.curfit_batch(iopt, x, y, w, k, s, nest)


def test_case_list_replayed_on_the_gpu_equals_the_oracle():
    got, exp = MRG.run_curfit_cases(_GpuBackend), MRG.run_curfit_cases(_OracleBackend)
    bad = [k for k in exp if not np.array_equal(exp[k], got[k], equal_nan=True)]
    assert not bad, bad[:10]


def test_65536_config2_curves_bit_exact_vs_oracle():
    """the config-2 sample at the size SURVEY 8(d) names: n, ier, fp and hashes of t, c of every curve"""
    got, exp = MRG.run_c2(_GpuBackend, 65536), MRG.run_c2(_OracleBackend, 65536)
    for k in ("n", "ier", "fp", "hash_t", "hash_c"):
        assert np.array_equal(exp[k], got[k]), k
    assert exp["n"].max() > 60 and (exp["ier"] <= 0).all()


@pytest.mark.skipif(not os.path.exists(os.path.join(GOLD, "ref_curfit.npz")),
                    reason="reference goldens not generated: needs gfortran (oracle/ref_recipe/build_ref.sh)")
def test_gpu_matches_the_fortran_reference_on_the_case_list():
    compare_curfit(os.path.join(GOLD, "ref_curfit.npz"), _GpuBackend)


@pytest.mark.skipif(not os.path.exists(os.path.join(GOLD, "ref_c2.npz")),
                    reason="reference goldens not generated: needs gfortran (oracle/ref_recipe/build_ref.sh)")
def test_gpu_matches_the_fortran_reference_on_65536_config2_curves():
    compare_c2(os.path.join(GOLD, "ref_c2.npz"), _GpuBackend)
