"""A/B timing of the config-2 fit pipeline for two builds of the library.
usage: python profiles/scripts/ab_fit.py [lib.so ...]   (default: the in-tree library)"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
CHILD = r'''
import os, sys
sys.path.insert(0, %r)
import torch
from fitpack_b200 import _lib
import ctypes as C
L = C.CDLL(_lib.LIB_PATH)
_lib.SYMBOLS = {k: v for k, v in _lib.SYMBOLS.items() if hasattr(L, k)}
from fitpack_b200.batch import Engine
sys.path.insert(0, os.path.join(%r, "tests"))
import bench
dev = torch.device("cuda", 0)
eng = Engine(dev)
x, y = bench.synth_device(torch, dev, 0, bench.B_PER_GPU, bench.M)
out = {}
for _ in range(3):
    r = eng.curfit_batch(x, y, None, k=3, s=bench.S, nest=bench.NEST, want_stats=True, out=out)
    out = {kk: r[kk] for kk in ("n", "t", "c", "fp", "ier", "stats")}
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ts = []
for rep in range(5):
    e0.record()
    for _ in range(3):
        eng.curfit_batch(x, y, None, k=3, s=bench.S, nest=bench.NEST, want_stats=True, out=out)
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) / 3)
import hashlib
r = eng.curfit_batch(x, y, None, k=3, s=bench.S, nest=bench.NEST, want_stats=True, out=out)
torch.cuda.synchronize()
hh = hashlib.sha1()
for kk in ("n", "ier", "fp", "stats"):
    hh.update(r[kk].cpu().numpy().tobytes())
# t, c: only the meaningful prefix of every row (the rest of a row is unspecified scratch)
nn = r["n"].cpu().numpy()
tt, cc = r["t"][:65536].cpu().numpy(), r["c"][:65536].cpu().numpy()
for b in range(0, 65536, 7):
    hh.update(tt[b, :nn[b]].tobytes()); hh.update(cc[b, :nn[b] - 4].tobytes())
print(os.path.basename(_lib.LIB_PATH), os.environ.get("FPB_FIT_NO_OVERLAP", ""), " ".join("%%.2f" %% t for t in ts), "ms/step; best %%.3f M fits/s" %% (bench.B_PER_GPU / min(ts) / 1e3), "sha1", hh.hexdigest()[:16], flush=True)
''' % (ROOT, ROOT)

for lib in (sys.argv[1:] or [os.path.join(ROOT, "fitpack_b200", "libfitpack_b200.so")]):
    env = dict(os.environ, FITPACK_B200_LIB=os.path.abspath(lib))
    subprocess.run([sys.executable, "-c", CHILD], env=env, cwd=ROOT, check=False)
