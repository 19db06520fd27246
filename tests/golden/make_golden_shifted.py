"""Generate tests/golden/ref_shifted.npz from the reference's own shifted_lopbicg_switching (shifted_switching_solver.c:260,
compiled in place into oracle/_ref/libref_strict.so by oracle/Makefile).  Run in the build container only:
    python tests/golden/make_golden_shifted.py
Stores, per case, what the REFERENCE produced: return value, every x_j, the seed residual r and the per-iteration
sqrt(dot_r/dot_zero) history.  Cases include ones where the seed converges first, so the seed-switching branch
(shifted_switching_solver.c:490-527) is exercised."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import mpi_bicgstab_b200 as B
import oracle as O
from helpers import SHIFTED_CASES, global_csr, shifted_problem

out = {}
for name, kind, g, p0, L, scale, seed in SHIFTED_CASES:
    blk, n, ptr, col, val = global_csr(B, kind, g, p0)
    sigma, b = shifted_problem(O, n, ptr, col, val, L, scale, seed)
    r = O.ref_shifted_solve(n, ptr, col, val, b, sigma, seed, tol=1e-12, max_iter=1000)
    out[name + "|ret"] = np.int64(r["ret"])
    out[name + "|x"] = r["x"]
    out[name + "|r"] = r["r"]
    out[name + "|res"] = r["res"]
    print(name, r["ret"], len(r["res"]))
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ref_shifted.npz"), **out)
print("written", len(out), "arrays")
