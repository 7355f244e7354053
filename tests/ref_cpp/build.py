"""Build the reference's own C++ test of the curve handle API against libfitpack_b200.so.

`/root/reference/test/test_curve.cpp` is compiled from where it lies (never copied into the repo) with
one section per function; the link keeps only what `test_cpp_sine_fit` reaches (`--gc-sections`), so the
periodic / parametric / constrained wrappers of that file, whose C symbols are outside this repo's
scope, drop out.  The binary lands in tests/_refbin/ (git-ignored, travels to the GPU box with the
snapshot like the built .so files); `tests/test_gpu_ref_cpp.py` RUNS it on the GPU.
"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF_TEST = "/root/reference/test/test_curve.cpp"
OUT_DIR = os.path.join(ROOT, "tests", "_refbin")
EXE = os.path.join(OUT_DIR, "ref_sine_fit")


def build():
    """Returns the path of the test binary, or None when the reference tree is not mounted."""
    if not os.path.exists(REF_TEST):
        return EXE if os.path.exists(EXE) else None
    libdir = os.path.join(ROOT, "fitpack_b200")
    os.makedirs(OUT_DIR, exist_ok=True)
    obj = os.path.join(OUT_DIR, "test_curve.o")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-ffunction-sections", "-fdata-sections", "-c",
                           REF_TEST, "-o", obj])
    subprocess.check_call(["g++", os.path.join(HERE, "run_sine_fit.cpp"), obj, "-Wl,--gc-sections",
                           "-L", libdir, "-lfitpack_b200", "-Wl,-rpath,$ORIGIN/../../fitpack_b200",
                           "-L/usr/local/cuda/lib64", "-lcudart", "-Wl,-rpath,/usr/local/cuda/lib64",
                           "-o", EXE])
    os.remove(obj)
    return EXE


if __name__ == "__main__":
    print(build())
