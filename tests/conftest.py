import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    path = os.path.join(ROOT, "tests", "golden", "cdr_golden.npz")
    return dict(np.load(path))


@pytest.fixture(scope="session")
def pkg():
    """The product package with libcdrhead.so present (built in-tree if missing)."""
    import fast_3d_human_pose_estimation_b200 as p
    if not os.path.exists(p._lib.LIB_PATH):
        p.build()
    return p


@pytest.fixture(scope="session")
def cuda_pkg(pkg):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return pkg
