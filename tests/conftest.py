import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    import oracle as o  # test infrastructure (oracle/__init__.py)

    o.build()
    return o


@pytest.fixture(scope="session")
def golden_k13():
    import numpy as np

    d = os.path.join(GOLDEN, "msm_k13")
    return {
        "bases": np.fromfile(os.path.join(d, "bases.bin"), dtype=np.uint8),
        "scalars": np.fromfile(os.path.join(d, "scalars.bin"), dtype=np.uint8),
        "result_affine": np.fromfile(os.path.join(d, "result_affine.bin"), dtype=np.uint8),
    }
