import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session")
def shipped(golden_dir):
    """(hps, checkpoint variables) of the reference's shipped S-Ax4-G-Ax4 model (tests/golden/NoiseFlow)."""
    from noise_flow_b200 import hps_loader, load_checkpoint
    hps = hps_loader(os.path.join(golden_dir, "NoiseFlow", "hps.txt"))
    ck = load_checkpoint(os.path.join(golden_dir, "NoiseFlow", "ckpt", "model.ckpt.best"))
    return hps, ck
