"""a few fp_curfit_c calls on the bench-shaped curve (for ncu captures of curfit_solo_kernel)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
import fitpack_b200 as F
rng = np.random.default_rng(0)
m = 256
x = np.sort(rng.uniform(0, 1, m)); y = np.sin(6 * x) + 0.1 * rng.standard_normal(m)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 4):
    r = F.curfit(0, x, y, None, 0, 1, 3, 2.56, 359)
print(r["n"], r["ier"], r["fp"])
