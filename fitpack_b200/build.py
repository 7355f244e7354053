"""Build libfitpack_b200.so (CUDA kernels + C-ABI) in-tree for sm_100a.

nvcc cross-compiles without a GPU.  The .so lands next to this file so that it
travels with the repo snapshot to the GPU box.
"""
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
BUILD = os.path.join(HERE, "_build")
LIB = os.path.join(HERE, "libfitpack_b200.so")

SOURCES = ["fpb_curfit.cu", "fpb_curfit_d2.cu", "fpb_curfit_d3.cu", "fpb_curfit_d0.cu", "fpb_curfit_solo.cu", "fpb_splev.cu", "fpb_shared.cu", "fpb_util.cu", "fpb_regrid.cu", "fpb_calculus.cu", "fpb_parcur.cu", "fpb_percur.cu", "fpb_percur_solo.cu",
           "fpb_engine.cu",
           "fpb_curve.cpp"]

# -fmad=false: the reference is built without FMA contraction (gfortran, baseline x86-64);
# parity of the discrete outputs (n, knots, ier) needs the same roundings.
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-O2", "-Xptxas", "-v"]


def nvcc_path():
    p = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(p):
        raise RuntimeError("nvcc not found: the CUDA extension cannot be built")
    return p


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    tm = os.path.getmtime(target)
    return any(os.path.getmtime(d) > tm for d in deps)


def build_variant(tag, defines):
    """Experimental side build (tuning only): same sources with extra -D flags, written to
    libfitpack_b200_<tag>.so; select it at run time with FITPACK_B200_LIB=<path>."""
    nvcc = nvcc_path()
    bdir = os.path.join("/tmp", "fitpack_b200_var_" + tag)   # objects out of tree: the GPU snapshot is size-capped
    os.makedirs(bdir, exist_ok=True)
    objs = []
    for src in SOURCES:
        o = os.path.join(bdir, os.path.splitext(src)[0] + ".o")
        cmd = [nvcc] + NVCC_FLAGS + ["-D" + d for d in defines] + ["-x", "cu", "-c",
                                                                   os.path.join(CSRC, src), "-o", o]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(r.stdout + r.stderr)
        objs.append(o)
    out = os.path.join(HERE, "libfitpack_b200_%s.so" % tag)
    subprocess.check_call([nvcc, "-shared", "-o", out] + objs + ["-lcudart", "-ldl"])
    return out


def build(force=False, verbose=False):
    os.makedirs(BUILD, exist_ok=True)
    nvcc = nvcc_path()
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    headers.append(os.path.join(os.path.dirname(HERE), "include", "fitpack_b200.h"))
    headers.append(os.path.abspath(__file__))
    objs, jobs = [], []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(BUILD, os.path.splitext(src)[0] + ".o")
        objs.append(o)
        deps = [s] + headers
        if src.startswith("fpb_curfit_d"):
            deps.append(os.path.join(CSRC, "fpb_curfit.cu"))
        if src == "fpb_percur_solo.cu":
            deps.append(os.path.join(CSRC, "fpb_percur.cu"))
        if force or _stale(o, deps):
            jobs.append((s, o))

    def compile_one(job):
        s, o = job
        cmd = [nvcc] + NVCC_FLAGS + ["-x", "cu", "-c", s, "-o", o]
        r = subprocess.run(cmd, capture_output=True, text=True)
        log = os.path.splitext(o)[0] + ".ptxas.log"
        with open(log, "w") as f:
            f.write(r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s" % (s, r.stdout + r.stderr))
        if verbose:
            print(r.stderr)
        return o

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            list(ex.map(compile_one, jobs))
    if force or jobs or _stale(LIB, objs):
        cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-lcudart", "-ldl"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
