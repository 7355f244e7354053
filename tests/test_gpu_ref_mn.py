"""The reference's mnspin and mnspro test programs (test/fitpack_tests.f90:4100-4320) through the C-ABI on the GPU
(fp_insert_c, fp_splint_c, fp_sproot_c, fp_curev_c), and the same intersection counts as the CPU oracle."""
import pytest

import oracle_lib as O
from ref_mn_cases import run_mnspin, run_mnspro

pytestmark = pytest.mark.gpu


def test_mnspin_and_mnspro_on_the_gpu():
    import fitpack_b200 as F
    assert run_mnspin(F) < 1e-13
    assert run_mnspro(F) == run_mnspro(O)
