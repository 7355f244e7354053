"""CPU: the reference's header-only C++ wrapper include/fpCurve.hpp compiles and LINKS against
libfitpack_b200.so with EVERY method it has (drop-in check of the whole handle C-ABI, include/fitpack_curves_c.h:45-63).  Only runs where
the reference tree is mounted (the build container); it is not executed (no GPU here)."""
import os
import subprocess
import sys
import tempfile

import pytest

import fitpack_b200

REF_INC = "/root/reference/include"

SRC = r"""
#include "fpCurve.hpp"
#include <cstdio>
int main(int argc, char**) {
    if (argc > 100) {               // never executed here: this is a compile + link check
        fpCurve c;
        std::vector<double> x{0,1,2,3,4,5,6,7}, y{0,1,4,9,16,25,36,49};
        FP_FLAG ier = c.new_fit(x, y, 1.0);
        ier = c.fit(0.5);
        ier = c.fit(0.5, 3);
        ier = c.interpolate();
        FP_SIZE e = 0;
        double v = c.eval(2.5, &e);
        std::vector<double> vv = c.eval(x, &e);
        c.set_bc(OUTSIDE_EXTRAPOLATE);
        fpCurve d(c);
        d = c;
        // every method of the wrapper: derivatives, integral, Fourier coefficients
        double d1 = c.ddx(2.5, 1, &e);
        std::vector<double> dall = c.ddx(2.5, &e);
        double area = c.integral(0.0, 7.0);
        std::vector<double> alpha{0.0, 1.0}, A, B;
        ier = c.fourier(alpha, A, B);
        std::printf("%d %g %zu %d %g %g %d %g %zu %g %zu\n", ier, v, vv.size(), c.degree(), c.smoothing(), c.mse(),
                    d.get_bc(), d1, dall.size(), area, A.size() + B.size());
    }
    return FITPACK_SUCCESS_c(0) ? 0 : 1;
}
"""


@pytest.mark.skipif(not os.path.isdir(REF_INC), reason="reference headers not mounted")
def test_reference_fpcurve_header_links_against_our_library():
    lib = fitpack_b200.ensure_built()
    libdir = os.path.dirname(lib)
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "use_fpcurve.cpp")
        exe = os.path.join(d, "use_fpcurve")
        with open(src, "w") as f:
            f.write(SRC)
        cmd = ["g++", "-std=c++17", "-O0", "-I", REF_INC, src, "-o", exe, "-L", libdir,
               "-lfitpack_b200", "-Wl,-rpath," + libdir, "-L/usr/local/cuda/lib64", "-lcudart",
               "-Wl,-rpath,/usr/local/cuda/lib64"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        # runs without touching the device (argc == 1)
        r = subprocess.run([exe], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
