"""ctypes mirror of include/hqp_hlcuda.h (libhqphl.so): the block-diagonal BFGS update of
the Lagrangian Hessian on the GPU (reference:
Hqp_HL_BFGS::update / update_b_Q, hqp/Hqp_HL_BFGS.C:149-243).  No CPU fallback."""
from __future__ import annotations

import ctypes
import os

import numpy as np

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libhqphl.so")
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            from . import build
            build.build_hl()
        _LIB = ctypes.CDLL(LIB_PATH)
        _LIB.hqphl_last_error.restype = ctypes.c_char_p
    return _LIB


def _dp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def bfgs_update(bsize, Q, s, u, alpha, gamma=0.1, eps=1e-8, eigen_control=True, device=0):
    """All diagonal blocks at once.  bsize: block sizes; Q: the blocks packed one after the
    other (row-major, both triangles); s, u: step / gradient difference in block order.
    Returns (Q_new, info) with info = (blocks shifted, blocks skipped, blocks whose
    eigenvalue iteration hit the sweep limit)."""
    bs = np.ascontiguousarray(bsize, dtype=np.int32)
    Qn = np.array(Q, dtype=np.float64, order="C", copy=True).ravel()
    s = np.ascontiguousarray(s, dtype=np.float64)
    u = np.ascontiguousarray(u, dtype=np.float64)
    if Qn.size != int(np.sum(bs.astype(np.int64) ** 2)) or s.size != int(bs.sum()) or u.size != s.size:
        raise ValueError("bfgs_update: sizes do not match the block structure")
    info = np.zeros(3, dtype=np.int32)
    rc = lib().hqphl_bfgs_update(ctypes.c_int(device), ctypes.c_int(bs.size),
                                 bs.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), _dp(Qn), _dp(s), _dp(u),
                                 ctypes.c_double(alpha), ctypes.c_double(gamma), ctypes.c_double(eps),
                                 ctypes.c_int(1 if eigen_control else 0),
                                 info.ctypes.data_as(ctypes.POINTER(ctypes.c_int)))
    if rc:
        raise RuntimeError(f"hqphl_bfgs_update: status {rc}: {lib().hqphl_last_error().decode()}")
    return Qn, tuple(int(v) for v in info)


def bfgs_update_dev(bsize_t, qoff_t, voff_t, Q_t, s_t, u_t, info_t, max_bsize, alpha, gamma=0.1, eps=1e-8,
                    eigen_control=True, stream=0):
    """Device-resident variant on torch CUDA tensors (int32 / int64 / int32 / float64 ...)."""
    rc = lib().hqphl_bfgs_update_dev(ctypes.c_void_p(stream), ctypes.c_int(bsize_t.numel()),
                                     ctypes.c_int(max_bsize), ctypes.c_void_p(bsize_t.data_ptr()),
                                     ctypes.c_void_p(qoff_t.data_ptr()), ctypes.c_void_p(voff_t.data_ptr()),
                                     ctypes.c_void_p(Q_t.data_ptr()), ctypes.c_void_p(s_t.data_ptr()),
                                     ctypes.c_void_p(u_t.data_ptr()), ctypes.c_double(alpha),
                                     ctypes.c_double(gamma), ctypes.c_double(eps),
                                     ctypes.c_int(1 if eigen_control else 0), ctypes.c_void_p(info_t.data_ptr()))
    if rc:
        raise RuntimeError(f"hqphl_bfgs_update_dev: status {rc}: {lib().hqphl_last_error().decode()}")
