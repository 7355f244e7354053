"""The reference's mnspin and mnspro test programs (test/fitpack_tests.f90:4100-4320) run against the CPU oracle."""
import oracle_lib as O
from ref_mn_cases import run_mnspin, run_mnspro


def test_mnspin_exact_integrals_of_polynomials():
    assert run_mnspin(O) < 1e-13


def test_mnspro_intersections_of_a_closed_curve_with_straight_lines():
    counts = run_mnspro(O)
    # a closed curve crosses a straight line an even number of times (0 for the last line, which misses it)
    assert len(counts) == 5 and all(m % 2 == 0 for m in counts) and all(m >= 2 for m in counts[:4])
