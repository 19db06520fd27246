"""One short BiCGStab solve of T' through the persistent kernel (target of the ncu --set full capture)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mpi_bicgstab_b200 as B
B.set_options(quiet=1, tol=0.0, max_iter=int(os.environ.get("BICG_MAX_ITER", "40")), mega=1)
blk = B.gen_block("stencil15", 117, 14.0)
dm = B.DeviceMatrix(blk)
b = dm.spmv(np.ones(blk.n)); x = np.zeros(blk.n)
it, st = dm.solve(os.environ.get("QP_METHOD", "bicgstab"), x, b)
print(it, st["loop_ms"], st["kernel_launches"])
dm.destroy()
