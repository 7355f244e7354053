"""Golden vectors produced by the reference ITSELF (gfortran build of /root/reference, recipe under
oracle/ref_recipe/): when tests/golden/ref_curfit.npz / ref_c2.npz exist, the oracle must reproduce them bit
for bit -- the pin to the Fortran binary.  This image has no Fortran compiler, so until the recipe has been
run on a box with gfortran the two pin tests SKIP (and the oracle's header says "parity unpinned"); the
plumbing (case list, replay, file format, comparison) is exercised on every run by a self-test that
generates the files with the oracle as the backend."""
import importlib.util
import os
import subprocess
import sys

import numpy as np
import pytest

import oracle_lib as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
_spec = importlib.util.spec_from_file_location("make_ref_golden", os.path.join(ROOT, "oracle", "ref_recipe",
                                                                              "make_ref_golden.py"))
MRG = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(MRG)


class _OracleBackend:
    curfit = staticmethod(O.curfit)

    @staticmethod
    def curfit_batch(iopt, x, y, w, k, s, nest):
        return O.curfit_batch(iopt, x, y, w, k, s, nest)


def compare_curfit(path, backend):
    ref = np.load(path)
    got = MRG.run_curfit_cases(backend)
    assert set(ref.files) == set(got), "case list changed since the file was written"
    bad = [k for k in ref.files if not np.array_equal(ref[k], got[k], equal_nan=True)]
    assert not bad, bad[:10]
    return len(ref.files)


def compare_c2(path, backend):
    ref = np.load(path)
    got = MRG.run_c2(backend, ref["n"].size)
    for k in ("n", "ier", "fp", "hash_t", "hash_c"):
        assert np.array_equal(ref[k], got[k]), k
    return ref["n"].size


def test_reference_golden_plumbing_self_test(tmp_path):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "ref_recipe", "make_ref_golden.py"),
                        "--backend", "oracle", "--out", str(tmp_path), "--c2-curves", "256"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert compare_curfit(str(tmp_path / "ref_curfit.npz"), _OracleBackend) > 1000
    assert compare_c2(str(tmp_path / "ref_c2.npz"), _OracleBackend) == 256


@pytest.mark.skipif(not os.path.exists(os.path.join(GOLD, "ref_curfit.npz")),
                    reason="reference goldens not generated: needs gfortran (oracle/ref_recipe/build_ref.sh)")
def test_oracle_matches_the_fortran_reference_on_the_case_list():
    compare_curfit(os.path.join(GOLD, "ref_curfit.npz"), _OracleBackend)


@pytest.mark.skipif(not os.path.exists(os.path.join(GOLD, "ref_c2.npz")),
                    reason="reference goldens not generated: needs gfortran (oracle/ref_recipe/build_ref.sh)")
def test_oracle_matches_the_fortran_reference_on_65536_config2_curves():
    compare_c2(os.path.join(GOLD, "ref_c2.npz"), _OracleBackend)
