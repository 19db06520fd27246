"""CPU, world_size 2 and 3 over gloo: bootstrap callback, halo plan exchange, device layout (see _gloo_worker.py)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world", [2, 3])
def test_host_side_of_multi_gpu_path(world):
    port = 29600 + world + (os.getpid() % 200)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "_gloo_worker.py")]
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="", OMP_NUM_THREADS="1")
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    assert f"GLOO_WORKER_OK {world}" in p.stdout


@pytest.mark.parametrize("world", [1, 3])
def test_shm_bootstrap_of_the_mpi_shim(world):
    """include/compat/mpi.h path (no torch, no MPI): ranks rendezvous through POSIX shm (csrc/shm_boot.cpp) and the
    registered allgather round-trips.  This is what MPI_Init() of the unchanged main.c does under tools/bicgrun."""
    code = ("import sys, ctypes; sys.path.insert(0, %r); import mpi_bicgstab_b200 as B; "
            "assert B.lib.bicg_shm_bootstrap() == 0; "
            "r, w = B.lib.bicg_comm_rank(), B.lib.bicg_comm_world(); "
            "assert B.lib.bicg_comm_selftest() == 0; "
            "rc = (ctypes.c_int * w)(); ds = (ctypes.c_int * w)(); B.lib.bicg_plan_partition(10, w, rc, ds); "
            "B.lib.bicg_shm_shutdown(); print('SHM_OK', r, w, flush=True)" % ROOT)
    env = dict(os.environ, WORLD_SIZE=str(world), BICG_JOB_ID=f"pytest{os.getpid()}_{world}", CUDA_VISIBLE_DEVICES="")
    procs = [subprocess.Popen([sys.executable, "-c", code], env=dict(env, RANK=str(r)), stdout=subprocess.PIPE,
                              stderr=subprocess.PIPE, text=True) for r in range(world)]
    outs = [p.communicate(timeout=120) for p in procs]
    for r, (p, (o, e)) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, e[-2000:]
        assert f"SHM_OK {r} {world}" in o
