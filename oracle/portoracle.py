"""ctypes wrapper around oracle/_build/liblqoracle.so (the plain-C restatement).

TEST INFRASTRUCTURE ONLY -- importable from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline leg, never from hqp_b200/.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class _Problem(ctypes.Structure):
    _fields_ = [("nx", ctypes.c_int), ("nu", ctypes.c_int), ("K", ctypes.c_int),
                ("fixed_x0", ctypes.c_int), ("m", ctypes.c_int),
                ("ineq_stage", ctypes.POINTER(ctypes.c_int)),
                ("ineq_ptr", ctypes.POINTER(ctypes.c_int)),
                ("ineq_lcol", ctypes.POINTER(ctypes.c_int)),
                ("ineq_val", ctypes.POINTER(ctypes.c_double)),
                ("Q", ctypes.POINTER(ctypes.c_double)),
                ("fx", ctypes.POINTER(ctypes.c_double)),
                ("fu", ctypes.POINTER(ctypes.c_double))]


def build():
    subprocess.run(["make", "-C", _HERE, "port"], check=True, capture_output=True)


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "_build", "liblqoracle.so")
        if not os.path.exists(path):
            build()
        _LIB = ctypes.CDLL(path)
        _LIB.lqo_alloc.restype = ctypes.c_void_p
        _LIB.lqo_residuum.restype = ctypes.c_double
        _LIB.lqo_Vxx.restype = ctypes.POINTER(ctypes.c_double)
        _LIB.lqo_Rux.restype = ctypes.POINTER(ctypes.c_double)
    return _LIB


def _dp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def _ip(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int))


class PortOracle:
    """init/update/factor/step/solve/residuum of the restated LQDOCP path."""

    def __init__(self, prob):
        if prob.n_eq:
            raise NotImplementedError("stage equality rows: use the compiled reference")
        self.prob = prob
        stage, lcol = prob.ineq_stage_local()
        self._keep = [stage, lcol, prob.ineq_ptr.astype(np.int32),
                      np.ascontiguousarray(prob.ineq_val, np.float64),
                      np.ascontiguousarray(prob.Q), np.ascontiguousarray(prob.fx),
                      np.ascontiguousarray(prob.fu)]
        k = self._keep
        self._p = _Problem(prob.nx, prob.nu, prob.K, int(prob.fixed_x0), prob.m,
                           _ip(k[0]), _ip(k[2]), _ip(k[1]), _dp(k[3]), _dp(k[4]),
                           _dp(k[5]), _dp(k[6]))
        self.h = ctypes.c_void_p(lib().lqo_alloc(ctypes.byref(self._p)))

    def factor(self, z, w):
        z = np.ascontiguousarray(z, np.float64)
        w = np.ascontiguousarray(w, np.float64)
        return lib().lqo_factor(self.h, _dp(z), _dp(w))

    def _out(self):
        p = self.prob
        return np.zeros(p.N), np.zeros(p.me), np.zeros(max(p.m, 1)), np.zeros(max(p.m, 1))

    def step(self, r1, r2, r3, r4):
        dx, dy, dz, dw = self._out()
        a = [np.ascontiguousarray(v, np.float64) for v in (r1, r2, r3, r4)]
        rc = lib().lqo_step(self.h, *[_dp(v) for v in a], _dp(dx), _dp(dy), _dp(dz), _dp(dw))
        if rc:
            raise ArithmeticError(f"oracle step: error {rc}")
        m = self.prob.m
        return dx, dy, dz[:m], dw[:m]

    def solve(self, r1, r2, r3, r4, eps=1e-10):
        dx, dy, dz, dw = self._out()
        a = [np.ascontiguousarray(v, np.float64) for v in (r1, r2, r3, r4)]
        res, n = ctypes.c_double(0), ctypes.c_int(0)
        rc = lib().lqo_solve(self.h, ctypes.c_double(eps), *[_dp(v) for v in a], _dp(dx),
                             _dp(dy), _dp(dz), _dp(dw), ctypes.byref(res), ctypes.byref(n))
        if rc:
            raise ArithmeticError(f"oracle solve: error {rc}")
        m = self.prob.m
        return dx, dy, dz[:m], dw[:m], res.value, n.value

    def residuum(self, r1, r2, r3, r4, dx, dy, dz, dw):
        a = [np.ascontiguousarray(v, np.float64) for v in (r1, r2, r3, r4, dx, dy, dz, dw)]
        return lib().lqo_residuum(self.h, *[_dp(v) for v in a])

    def Vxx(self):
        p = self.prob
        return np.ctypeslib.as_array(lib().lqo_Vxx(self.h), ((p.K + 1), p.nx, p.nx)).copy()

    def Rux(self):
        p = self.prob
        return np.ctypeslib.as_array(lib().lqo_Rux(self.h), (p.K, p.nu, p.nx)).copy()

    def close(self):
        if self.h:
            lib().lqo_free(self.h)
            self.h = None
