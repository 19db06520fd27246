"""Per-iteration time of each loop implementation on one workload (single or multi rank)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mpi_bicgstab_b200 as B
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
if world > 1:
    import torch, torch.distributed as dist
    os.environ["NCCL_DEBUG"] = "WARN"
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
B.set_options(device=local, quiet=1)
if world > 1:
    B.comm_init_torch()
blk = B.gen_block("stencil15", 117, 14.0, rank=rank, world=world)
dm = B.DeviceMatrix(blk)
nl = blk.n_loc
for method in (sys.argv[1:] or ["bicgstab", "ca_bicgstab", "pipe_bicgstab"]):
    for mode, kw in (("mega", dict(mega=1)), ("graph", dict(mega=0, graph=1))):
        B.set_options(tol=0.0, max_iter=300, **kw)
        for rep in range(2):
            b = dm.spmv(np.ones(nl)); x = np.zeros(nl)
            it, st = dm.solve(method, x, b)
        if rank == 0:
            print(f"[N={world}] {method:14s} {mode:6s} {st['loop_ms'] / it * 1e3:7.1f} us/it  {it / st['loop_ms'] * 1e3:8.0f} it/s", flush=True)
dm.destroy()
if world > 1:
    B.comm_finalize(); dist.destroy_process_group()
