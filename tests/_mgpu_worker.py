"""Multi-GPU parity worker (one process per GPU, torchrun + NCCL for the bootstrap only).

Every rank owns a row block, solves collectively through the C ABI and checks its slice against the oracle's
P-rank emulation (same partition, same diag-then-offd association, dots summed in rank order)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import mpi_bicgstab_b200 as B
import oracle as O
from helpers import METHODS, RR

TOL = 1e-10


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    B.set_options(device=local, quiet=1)
    rank, world = B.comm_init_torch()
    cases = [("stencil15", 14, 14.0), ("convdiff", 48, 1.5), ("random", 4001, 8), ("laplace5", 41, 0.0)]
    for kind, g, p0 in cases:
        blk = B.gen_block(kind, g, p0, rank=rank, world=world)
        n, nloc, lo = blk.n, blk.n_loc, int(blk.displs[rank])
        g1 = B.gen_block(kind, g, p0)
        ptr, col, val = B.block_to_global_csr(g1)
        xg = np.random.default_rng(11).standard_normal(n)
        dm = B.DeviceMatrix(blk)
        y = dm.spmv(np.ascontiguousarray(xg[lo:lo + nloc]))
        y_ref = O.spmv(n, ptr, col, val, xg, long_double=True)[lo:lo + nloc]
        assert np.abs(y - y_ref).max() <= 1e-13 * np.abs(y_ref).max(), ("spmv", kind, rank)
        b_ref = O.spmv(n, ptr, col, val, np.ones(n), P=world)
        for method, mega in [(m_, g_) for m_ in METHODS for g_ in (1, 0)]:
            kw = RR if method.endswith("rr") else {}
            B.set_options(tol=TOL, max_iter=600, mega=mega)
            b = dm.spmv(np.ones(nloc))
            x = np.zeros(nloc)
            it, st = dm.solve(method, x, b, **kw)
            hist = B.last_history()
            ref = O.solve(method, n, ptr, col, val, b_ref, P=world, tol=TOL, max_iter=600, **kw)
            m = min(10, it, ref["iters"])
            got, want = np.sqrt(hist[1:m + 1]), np.sqrt(ref["hist"][1:m + 1])
            assert np.all(np.abs(got - want) <= 1e-10 * want + 1e-15), (kind, method, rank, got, want)
            assert abs(it - ref["iters"]) <= max(2, int(0.02 * ref["iters"])), (kind, method, it, ref["iters"])
            assert np.abs(x - 1.0).max() <= 1e3 * max(np.abs(ref["x"] - 1.0).max(), TOL), (kind, method)
            # every rank ran the same number of iterations and saw the same scalars
            t = torch.tensor([float(it), float(hist[it])], dtype=torch.float64, device="cuda")
            tmax, tmin = t.clone(), t.clone()
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX); dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
            assert torch.equal(tmax, tmin), "ranks disagree on iteration count / residual"
            if rank == 0:
                print(f"[mgpu {world}] {kind:10s} {method:17s} mega={mega} {it:4d} it (oracle {ref['iters']}), launches {st['kernel_launches']}", flush=True)
        dm.destroy()
    # shifted family: one seed-switching case, every rank checks its slice of every x_j against the oracle's P-rank emulation
    from helpers import SHIFTED_CASES, shifted_problem
    name, kind, g, p0, L, scale, seed = SHIFTED_CASES[1]
    blk = B.gen_block(kind, g, p0, rank=rank, world=world)
    n, nloc, lo = blk.n, blk.n_loc, int(blk.displs[rank])
    ptr, col, val = B.block_to_global_csr(B.gen_block(kind, g, p0))
    sigma = (np.arange(L) + 1) * scale
    bg = O.spmv(n, ptr, col, val, np.ones(n), P=world)
    O.daxpy(sigma[seed], np.ones(n), bg)
    ref = O.shifted_solve(n, ptr, col, val, bg, sigma, seed, P=world, tol=1e-12, max_iter=1000)
    B.set_options(shift_tol=1e-12, shift_max_iter=1000)
    xs = np.zeros((L, nloc)); rs = np.ascontiguousarray(bg[lo:lo + nloc])
    ret = B.shifted_lopbicg_switching(blk, xs, rs, sigma, seed)
    end_seed, stop = B.last_shift_info(L)
    assert abs(ret - ref["ret"]) <= 2 and end_seed == ref["seed"], (ret, ref["ret"], end_seed, ref["seed"])
    for j in range(L):
        assert np.abs(xs[j] - ref["x"][j][lo:lo + nloc]).max() <= 1e-8 * np.abs(ref["x"][j]).max(), ("shifted", j, rank)
    if rank == 0:
        print(f"[mgpu {world}] shifted_lopbicg_switching {name}: ret {ret} (oracle {ref['ret']}), final seed {end_seed}, stops {stop}", flush=True)

    # the reference-facing host-pointer entry point, collectively
    blk = B.gen_block("stencil15", 14, 14.0, rank=rank, world=world)
    B.set_options(tol=TOL, max_iter=600)
    b = B.spmv_ovlap(blk, np.ones(blk.n_loc))
    x = np.zeros(blk.n_loc)
    it = B.bicgstab(blk, x, b)
    assert np.abs(x - 1).max() < 1e-6
    B.comm_finalize()
    dist.barrier()
    if rank == 0:
        print("MGPU_WORKER_OK", world, flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
