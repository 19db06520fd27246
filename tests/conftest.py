import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


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
