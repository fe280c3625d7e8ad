"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every
symbol include/hqp_ipcuda.h declares, and refuses to run without a GPU (no
silent CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

from hqp_b200 import ipcuda
from hqp_b200.problem import synth_lqdocp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(ROOT, "include", "hqp_ipcuda.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(hqpcu_\w+)\s*\(", text)))


def test_header_declares_the_plugin_entry_points():
    names = declared_functions()
    for required in ("hqpcu_create", "hqpcu_destroy", "hqpcu_update", "hqpcu_factor",
                     "hqpcu_step", "hqpcu_solve", "hqpcu_residuum"):
        assert required in names


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(ipcuda.LIB_PATH)
    for name in declared_functions():
        assert hasattr(lib, name), f"{name} declared in include/hqp_ipcuda.h but not exported"


def test_library_is_sm100a_sass():
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", ipcuda.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    assert "sm_100a" in out.stdout


def test_null_and_bad_dims_are_rejected():
    lib = ipcuda.lib()
    h = ctypes.c_void_p()
    assert lib.hqpcu_create(None, ctypes.byref(h)) == ipcuda.HQPCU_E_NULL
    dims = ipcuda.HqpcuDims(K=0, nx=2, nu=1, batch=1)
    assert lib.hqpcu_create(ctypes.byref(dims), ctypes.byref(h)) == ipcuda.HQPCU_E_SIZES


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    p = synth_lqdocp(2, 1, 4)
    with pytest.raises(RuntimeError):
        ipcuda.IpCuda(p)
