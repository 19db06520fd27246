"""Targets for `ncu --set full --profile-from-start off` captures of kernels other than the persistent solver kernel.
usage: ncu_targets.py random   -> the autotuned stand-alone SpMV on the cfg-5 per-GPU block (random 2 M x 32)
       ncu_targets.py shifted  -> a few iterations of shifted_lopbicg_switching with 64 shifts on T' (sh_vec_shift)
The profiled range starts after upload + autotune (cudaProfilerStart), so `-c N` counts from the launches of interest."""
import os, sys
import ctypes as C
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mpi_bicgstab_b200 as B
rt = C.CDLL("libcudart.so.12")
what = sys.argv[1]
B.set_options(quiet=1)
if what == "random":
    blk = B.gen_block("random", 2000000, 32.0)
    dm = B.DeviceMatrix(blk)
    x = np.random.default_rng(1).standard_normal(blk.n)
    y = dm.spmv(x)                                            # warm, autotuned
    rt.cudaProfilerStart()
    y = dm.spmv(x)
    rt.cudaProfilerStop()
    st = B.last_stats()
    print("random block spmv plan:", st["spmv_kind"], st["spmv_lanes"])
else:
    blk = B.gen_block("stencil15", 117, 14.0)
    dm = B.DeviceMatrix(blk)
    L = 64
    n = blk.n
    sigma = (np.arange(L) + 1) * (0.01 / L)
    b = dm.spmv(np.ones(n)) + sigma[0] * np.ones(n)
    B.set_options(shift_tol=1e-8, shift_max_iter=6)
    st = B.bicg_stats()
    x = np.zeros((L, n)); r = b.copy()
    rt.cudaProfilerStart()
    k = B.lib.bicg_shifted_solve(dm.h, x.ctypes.data_as(C.c_void_p), r.ctypes.data_as(C.c_void_p), sigma.ctypes.data_as(C.c_void_p), L, 0, C.byref(st))
    rt.cudaProfilerStop()
    print("shifted:", k, st.kernel_launches)
dm.destroy()
