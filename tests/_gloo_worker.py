"""world_size > 1 worker for tests/test_gloo_multirank.py (CPU, gloo): the host-side half of the multi-GPU path.

Every rank builds its own blocks, derives its device layout (merged CSR + ghost slots) and halo plan through the
C ABI, exchanges the plans over torch.distributed exactly like matrix_create() does through the registered
allgather callback, and then *emulates* the device data path in numpy: peers push their runs into the ghost
slots, the merged CSR multiplies the extended vector.  The result must equal the oracle's P-rank SpMV.
"""
import ctypes as C
import os
import sys

import numpy as np
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import mpi_bicgstab_b200 as B
import oracle as O


def main():
    dist.init_process_group("gloo")
    rank, world = B.comm_init_torch()
    assert (B.lib.bicg_comm_rank(), B.lib.bicg_comm_world()) == (rank, world)
    assert B.lib.bicg_comm_selftest() == 0                      # C -> Python callback -> gloo -> back

    for kind, g, p0, gap in (("stencil15", 9, 14.0, 64), ("random", 997, 6, 0), ("random", 997, 6, 64), ("convdiff", 21, 1.5, 4)):
        blk = B.gen_block(kind, g, p0, rank=rank, world=world)
        n, nloc = blk.n, blk.n_loc
        lo = int(blk.displs[rank])
        ghost_off = (nloc + 15) // 16 * 16
        nnz = blk.nnz_loc
        mptr = np.zeros(nloc + 1, dtype=np.uint32); mcol = np.zeros(max(nnz, 1), dtype=np.uint32)
        mval = np.zeros(max(nnz, 1)); recv = np.zeros(4 * (nnz + 1), dtype=np.int32); ng = C.c_int()
        nrecv = B.lib.bicg_plan_merge(C.byref(blk.diag), C.byref(blk.offd), C.byref(blk.info), rank, world, gap, ghost_off,
                                      mptr.ctypes.data_as(C.POINTER(C.c_uint)), mcol.ctypes.data_as(C.POINTER(C.c_uint)),
                                      mval.ctypes.data_as(C.POINTER(C.c_double)), recv.ctypes.data_as(C.POINTER(C.c_int)),
                                      recv.size, C.byref(ng))
        assert nrecv >= 0
        recv = recv[:4 * nrecv].reshape(-1, 4)
        n_ghost = ng.value
        assert np.all(recv[:, 2] != rank) and (recv[:, 1].sum() == n_ghost)
        assert np.all(mcol[:nnz] < ghost_off + max(n_ghost, 1))

        # exchange the receive lists (two rounds, like matrix.cu)
        cnts = [None] * world
        dist.all_gather_object(cnts, int(nrecv))
        max_cnt = max(1, max(cnts))
        mine = np.zeros(4 * max_cnt, dtype=np.int32); mine[:4 * nrecv] = recv.ravel()
        allr = [None] * world
        dist.all_gather_object(allr, mine)
        all_recv = np.concatenate(allr).astype(np.int32)
        cnts_c = (C.c_int * world)(*cnts)

        # a global test vector every rank can reproduce; each rank only "owns" its slice
        xg = np.random.default_rng(5).standard_normal(n)
        x_ext = np.full(ghost_off + max(n_ghost, 1) + 16, np.nan)
        x_ext[:nloc] = xg[lo:lo + nloc]

        # what every peer pushes to me: derive with the C planner from the peer's point of view and apply
        displs = blk.displs
        for src in range(world):
            if src == rank:
                continue
            out = np.zeros(3 * max_cnt, dtype=np.int32)
            k = B.lib.bicg_plan_push_runs(all_recv.ctypes.data_as(C.POINTER(C.c_int)), cnts_c, 4 * max_cnt, src, rank,
                                          int(displs[src]), out.ctypes.data_as(C.POINTER(C.c_int)), out.size)
            assert k >= 0
            runs = out[:3 * k].reshape(-1, 3)
            assert np.all(np.diff(runs[:, 0]) > 0)                                  # sorted by source row, disjoint
            for s_, l_, d_ in runs:
                assert 0 <= s_ and s_ + l_ <= blk.recvcounts[src]
                x_ext[ghost_off + d_:ghost_off + d_ + l_] = xg[displs[src] + s_:displs[src] + s_ + l_]
        assert not np.isnan(x_ext[ghost_off:ghost_off + n_ghost]).any()             # every ghost slot was filled

        # emulate the device SpMV on the merged CSR and compare with the oracle's P-rank SpMV
        y = np.zeros(nloc)
        rows = np.repeat(np.arange(nloc), np.diff(mptr.astype(np.int64)))
        np.add.at(y, rows, mval[:nnz] * x_ext[mcol[:nnz]])
        g1 = B.gen_block(kind, g, p0)
        ptr, col, val = B.block_to_global_csr(g1)
        y_ref = O.spmv(n, ptr, col, val, xg, P=world)[lo:lo + nloc]
        assert np.abs(y - y_ref).max() <= 1e-12 * np.abs(y_ref).max(), (kind, rank)
        if gap == 0:                                                                # exact halo: nothing over-fetched
            need = np.unique(np.asarray(blk.offd_arrays()[1])[:int(blk.offd.nz)])
            assert n_ghost == need.size
    B.comm_finalize()
    dist.barrier()
    if rank == 0:
        print("GLOO_WORKER_OK", world)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
