import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

ORACLE_SO = os.path.join(ROOT, "oracle", "libgms_ref.so")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (oracle/libgms_ref.so), built on demand.  Test infrastructure only."""
    from gridmap_slam_robot_b200 import binding

    src = os.path.join(ROOT, "oracle", "gms_ref.c")
    if not os.path.exists(ORACLE_SO) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")])
    return binding.Library(ORACLE_SO)


@pytest.fixture(scope="session")
def cuda():
    """The CUDA product library.  No fallback: a missing .so or device is a test failure."""
    import torch

    from gridmap_slam_robot_b200 import binding

    assert torch.cuda.is_available(), "gpu-marked test without a CUDA device"
    return binding.load()
