"""CPU: the C-ABI library builds for sm_100a, loads, and exports every symbol that
include/fitpack_b200.h declares.  No compute call is made (no GPU here)."""
import os
import re

import pytest

import fitpack_b200
from fitpack_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "fitpack_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = set(re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\(", src))
    return {n for n in names if n.startswith(("fp_", "fpb_", "fitpack_", "FITPACK_"))}


def test_library_builds_and_loads():
    path = fitpack_b200.ensure_built()
    assert os.path.exists(path)
    L = fitpack_b200.lib()
    assert b"sm_100a" in L.fpb_version()


def test_every_declared_symbol_is_exported_and_bound():
    declared = _declared_symbols()
    assert len(declared) >= 40
    L = fitpack_b200.lib()
    for name in declared:
        assert hasattr(L, name), "missing export " + name
        assert name in _lib.SYMBOLS, "python binding table lacks " + name
    assert set(_lib.SYMBOLS) <= declared | {"fpb_engine"}


def test_messages_and_success_flags():
    assert fitpack_b200.FITPACK_SUCCESS(0) and fitpack_b200.FITPACK_SUCCESS(-2)
    assert not fitpack_b200.FITPACK_SUCCESS(1)
    import oracle_lib as O
    for ier in (-2, -1, 0, 1, 2, 3, 6, 10, 99):
        assert fitpack_b200.FITPACK_MESSAGE(ier) == O.message(ier)


def test_no_cpu_fallback_without_gpu():
    import numpy as np
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(fitpack_b200.FitpackB200Error):
        fitpack_b200.curfit(0, np.arange(10.0), np.arange(10.0), None, 0, 9, 3, 1.0, 20)
    with pytest.raises(fitpack_b200.FitpackB200Error):
        t = np.array([0., 0, 0, 0, 1, 1, 1, 1])
        fitpack_b200.splev(t, 8, np.ones(8), 3, np.array([0.5]))
    # data checks that need no data still answer like the reference (ier=10), no device touched
    assert fitpack_b200.curfit(0, np.arange(10.0), np.arange(10.0), None, 0, 9, 7, 1.0, 20)["ier"] == 10


def test_product_does_not_reference_the_oracle():
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "fitpack_b200")):
        if "_build" in dirpath or "__pycache__" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                if re.search(r"oracle_lib|fitpack_oracle|orc_curfit|orc_splev", txt):
                    bad.append(f)
    assert not bad, bad


REF_INC = "/root/reference/include"


@pytest.mark.skipif(not os.path.isdir(REF_INC), reason="reference headers not mounted")
def test_every_function_of_the_reference_handle_header_is_exported():
    """include/fitpack_curves_c.h of the reference declares 19 fitpack_curve_c_* functions: all must be exported by
    libfitpack_b200.so; of the flat layer (include/fitpack_core_c.h) every 1-D curve routine except clocur and the
    constrained fitters, plus regrid / bispev"""
    import re
    import subprocess
    import fitpack_b200
    lib = fitpack_b200.ensure_built()
    out = subprocess.check_output(["nm", "-D", "--defined-only", lib], text=True)
    exported = {line.split()[-1] for line in out.splitlines() if line.strip()}
    handle = set(re.findall(r"\b(fitpack_curve_c_[a-z_]+)\s*\(", open(os.path.join(REF_INC, "fitpack_curves_c.h")).read()))
    assert len(handle) == 19 and not (handle - exported), sorted(handle - exported)
    flat = set(re.findall(r"\b(fp_[a-z]+_c|FITPACK_SUCCESS_c|fitpack_message_c)\s*\(",
                          open(os.path.join(REF_INC, "fitpack_core_c.h")).read()))
    not_provided = {"fp_bispeu_c", "fp_clocur_c", "fp_cocosp_c", "fp_concon_c", "fp_parder_c", "fp_parsur_c",
                    "fp_pogrid_c", "fp_polar_c", "fp_spgrid_c", "fp_sphere_c", "fp_surfit_c"}
    assert flat - exported == not_provided, sorted(flat - exported)
