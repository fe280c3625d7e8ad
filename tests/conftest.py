import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    # native pieces the tests need: synthetic generator, C oracle port, C-ABI lib
    from hqp_b200 import build
    build.build_synth()
    build.build_cuda()
    from oracle import portoracle
    portoracle.build()


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")
