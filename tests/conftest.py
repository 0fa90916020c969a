import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "electrocardio-panorama_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


class Cfg:
    """Stand-in for the yacs node the reference's losswrapper reads (losses.py:26-44, nef_net.yml:9)."""

    class SOLVER:
        reg_loss = "l1_loss"
        loss_using = [1, 2, 3]
        loss_factor = [0.5, 0.5, 1]

    class MODEL:
        model = "model_nefnet"
        theta_L = 1
        loss = "v1"

    class DATA:
        lead_num = 3


@pytest.fixture
def cfg():
    return Cfg
