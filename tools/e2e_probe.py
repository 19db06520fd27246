"""Where does the host-pointer entry point spend its time?  (BICG_VERBOSE=2 breakdown)"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mpi_bicgstab_b200 as B
sys.path.insert(0, ROOT)
from bench import pinned_block, pinned_array
B.set_options(quiet=1, verbose=2, cache=0, tol=1e-8, max_iter=1000)
blk = B.gen_block("stencil15", 117, 14.0)
pblk = pinned_block(B, blk, 0, 1)
n = blk.n
xh, rh = pinned_array(B, n, np.float64), pinned_array(B, n, np.float64)
B.set_options(cache=1)
b = B.spmv_ovlap(pblk, np.ones(n))
B.lib.bicg_matrix_invalidate(pblk.diag)
B.set_options(cache=0)
for rep in range(3):
    xh[:] = 0; rh[:] = b
    t0 = time.perf_counter()
    it = B.bicgstab(pblk, xh, rh)
    print("call", rep, it, "iters", round(1e3 * (time.perf_counter() - t0), 2), "ms", flush=True)
