import os
import shutil
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

_NO_CUDA_LIB = None  # reason the CUDA library could not be built (CPU-only checkout)


def pytest_configure(config):
    global _NO_CUDA_LIB
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    # native pieces the tests need: synthetic generator, C oracle port, C-ABI lib
    from hqp_b200 import build
    build.build_synth()
    from oracle import portoracle
    portoracle.build()
    # The CUDA library is git-ignored: on a checkout without nvcc and without a
    # prebuilt .so only the tests that need it are skipped (oracle / gloo tests run).
    lib = os.path.join(build.LIB, "libhqpcuda.so")
    if shutil.which("nvcc"):
        build.build_cuda()
        build.build_hl()
        build.build_docp()
    elif not os.path.exists(lib):
        _NO_CUDA_LIB = "nvcc not found and hqp_b200/lib/libhqpcuda.so not prebuilt"


def pytest_collection_modifyitems(config, items):
    if not _NO_CUDA_LIB:
        return
    skip = pytest.mark.skip(reason=_NO_CUDA_LIB)
    for item in items:
        if ("gpu" in item.keywords or "test_abi" in item.nodeid or "test_library_exports" in item.nodeid
                or "test_bad_arguments" in item.nodeid):
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")
