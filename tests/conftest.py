import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def _cuda_device_count():
    """Device count straight from the driver (no torch import, no context): 0 when there is no usable GPU."""
    import ctypes
    try:
        cu = ctypes.CDLL("libcuda.so.1")
        if cu.cuInit(0) != 0:
            return 0
        n = ctypes.c_int(0)
        return n.value if cu.cuDeviceGetCount(ctypes.byref(n)) == 0 else 0
    except OSError:
        return 0


def pytest_collection_modifyitems(config, items):
    """The library has no CPU path -- its compute entry points exit(1) without a GPU (reference error convention) --
    so on a box without one the gpu-marked tests are reported as skipped instead of killing the pytest process."""
    if _cuda_device_count() > 0:
        return
    skip = pytest.mark.skip(reason="no CUDA device visible: GPU tests need a B200 (run under gpurun)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def B():
    import mpi_bicgstab_b200 as pkg
    return pkg


@pytest.fixture(scope="session")
def O():
    import oracle as orc
    if not orc.have_oracle():
        import subprocess
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "oracle"], check=True)
    return orc
