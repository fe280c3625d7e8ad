"""ctypes binding of libhqpcuda.so (include/hqp_ipcuda.h).

Python mirror of the reference's matrix-module interface
``Hqp_IpMatrix::{init,update,factor,step,solve,residuum}``
(hqp/Hqp_IpMatrix.h:63-88) for tests/ and bench.py; the C++ mirror that plugs
into HQP itself is hqp_b200/host/Hqp_IpCuda.C.  There is no fallback: if the
CUDA library is missing or a call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libhqpcuda.so")
_LIB = None

HQPCU_OK, HQPCU_E_SIZES, HQPCU_E_SING, HQPCU_E_NULL = 0, 1, 4, 8
HQPCU_E_CUDA, HQPCU_E_UNSUPPORTED, HQPCU_E_NOTPD = 100, 101, 102

_c_int_p = ctypes.POINTER(ctypes.c_int)
_c_dbl_p = ctypes.POINTER(ctypes.c_double)


class HqpcuDims(ctypes.Structure):
    _fields_ = [("K", ctypes.c_int), ("nx", ctypes.c_int), ("nu", ctypes.c_int),
                ("batch", ctypes.c_int), ("fixed_x0", ctypes.c_int),
                ("n_ineq", ctypes.c_int), ("ineq_stage", _c_int_p),
                ("ineq_ptr", _c_int_p), ("ineq_lcol", _c_int_p),
                ("n_eq", ctypes.c_int), ("eq_stage", _c_int_p), ("eq_ptr", _c_int_p),
                ("eq_lcol", _c_int_p), ("device", ctypes.c_int), ("nseg", ctypes.c_int),
                ("ngpu", ctypes.c_int)]


class SingularError(ArithmeticError):
    """HQPCU_E_SING: what the host module turns into m_error(E_SING, ...)."""


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} missing: build it with "
                               "`python -c 'import __graft_entry__ as g; g.build()'`")
        _LIB = ctypes.CDLL(LIB_PATH)
        _LIB.hqpcu_last_error.restype = ctypes.c_char_p
        _LIB.hqpcu_launch_count.restype = ctypes.c_longlong
        _LIB.hqpcu_launch_count.argtypes = [ctypes.c_void_p]
        _LIB.hqpcu_nseg.argtypes = [ctypes.c_void_p]
    return _LIB


def _check(rc, what):
    if rc == HQPCU_OK:
        return
    msg = lib().hqpcu_last_error().decode()
    if rc == HQPCU_E_SING:
        raise SingularError(f"{what}: singular stage block (E_SING)")
    raise RuntimeError(f"{what}: status {rc} {msg}")


def _hp(a):
    """host double pointer of a numpy array (None -> NULL)"""
    return a.ctypes.data_as(_c_dbl_p) if a is not None else None


def _ip(a):
    return a.ctypes.data_as(_c_int_p)


def _vp(p):
    """device pointer (int) -> void*"""
    return ctypes.c_void_p(int(p))


class IpCuda:
    """B200 KKT engine for one LQ-DOCP structure (optionally a batch of them).

    Host-pointer methods take/return numpy arrays (copies included, like the
    plugin); ``*_dev`` methods take raw CUDA device pointers (ints, e.g.
    ``torch.Tensor.data_ptr()``) and enqueue on the handle's stream.
    """

    def __init__(self, prob, batch=1, device=0, nseg=0, ngpu=0):
        self.prob = prob
        self.batch = batch
        stage, lcol = prob.ineq_stage_local()
        ptr = np.ascontiguousarray(prob.ineq_ptr, np.int32)
        estage, elcol = prob.eq_stage_local()
        eptr = np.ascontiguousarray(prob.eq_ptr, np.int32)
        self._keep = (stage, lcol, ptr, estage, elcol, eptr)
        dims = HqpcuDims(prob.K, prob.nx, prob.nu, batch, int(prob.fixed_x0), prob.m,
                         _ip(stage), _ip(ptr), _ip(lcol), prob.n_eq, _ip(estage),
                         _ip(eptr), _ip(elcol), device, nseg, ngpu)
        self.h = ctypes.c_void_p()
        _check(lib().hqpcu_create(ctypes.byref(dims), ctypes.byref(self.h)), "hqpcu_create")
        self.N, self.me, self.m = prob.N, prob.me, prob.m

    # -- life cycle -------------------------------------------------------
    def close(self):
        if self.h:
            lib().hqpcu_destroy(self.h)
            self.h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- horizon split (hqpcu_comm_*): this handle = one rank's stage range ----
    @staticmethod
    def comm_unique_id() -> bytes:
        buf = (ctypes.c_ubyte * 128)()
        _check(lib().hqpcu_comm_unique_id(buf), "hqpcu_comm_unique_id")
        return bytes(buf)

    def comm_init(self, uid: bytes, rank: int, world: int):
        buf = (ctypes.c_ubyte * 128).from_buffer_copy(uid) if uid is not None else None
        _check(lib().hqpcu_comm_init(self.h, buf, rank, world), "hqpcu_comm_init")

    def comm_info(self):
        r, w, m = ctypes.c_int(0), ctypes.c_int(0), ctypes.c_longlong(0)
        _check(lib().hqpcu_comm_info(self.h, ctypes.byref(r), ctypes.byref(w), ctypes.byref(m)),
               "hqpcu_comm_info")
        return r.value, w.value, m.value

    def set_stream(self, cuda_stream: int):
        _check(lib().hqpcu_set_stream(self.h, ctypes.c_void_p(cuda_stream)), "set_stream")

    @property
    def launches(self):
        return lib().hqpcu_launch_count(self.h)

    @property
    def nseg(self):
        return lib().hqpcu_nseg(self.h)

    def set_nseg(self, nseg):
        _check(lib().hqpcu_set_nseg(self.h, nseg), "set_nseg")

    def sync_status(self):
        return lib().hqpcu_sync_status(self.h)

    def profile(self, on=True):
        _check(lib().hqpcu_profile(self.h, int(on)), "hqpcu_profile")

    def profile_read(self):
        import json
        buf = ctypes.create_string_buffer(8192)
        _check(lib().hqpcu_profile_read(self.h, buf, len(buf)), "hqpcu_profile_read")
        return json.loads(buf.value.decode())

    # -- host-pointer API (what the plugin calls) --------------------------
    def update(self, Q=None, fx=None, fu=None, ineq_val=None, eq_val=None):
        p = self.prob
        arrs = [np.ascontiguousarray(a if a is not None else d, np.float64)
                for a, d in ((Q, p.Q), (fx, p.fx), (fu, p.fu), (ineq_val, p.ineq_val),
                             (eq_val, p.eq_val))]
        _check(lib().hqpcu_update(self.h, *[_hp(a) for a in arrs]), "hqpcu_update")

    def factor(self, z, w):
        z = np.ascontiguousarray(z, np.float64)
        w = np.ascontiguousarray(w, np.float64)
        _check(lib().hqpcu_factor(self.h, _hp(z), _hp(w)), "hqpcu_factor")

    def _outs(self):
        B = self.batch
        return (np.zeros(B * self.N), np.zeros(B * self.me), np.zeros(max(B * self.m, 1)),
                np.zeros(max(B * self.m, 1)))

    def step(self, r1, r2, r3, r4):
        a = [np.ascontiguousarray(v, np.float64) for v in (r1, r2, r3, r4)]
        dx, dy, dz, dw = self._outs()
        _check(lib().hqpcu_step(self.h, *[_hp(v) for v in a], _hp(dx), _hp(dy), _hp(dz),
                                _hp(dw)), "hqpcu_step")
        n = self.batch * self.m
        return dx, dy, dz[:n], dw[:n]

    def solve(self, r1, r2, r3, r4, eps=1e-10):
        a = [np.ascontiguousarray(v, np.float64) for v in (r1, r2, r3, r4)]
        dx, dy, dz, dw = self._outs()
        res, n = ctypes.c_double(0), ctypes.c_int(0)
        _check(lib().hqpcu_solve(self.h, ctypes.c_double(eps), *[_hp(v) for v in a], _hp(dx),
                                 _hp(dy), _hp(dz), _hp(dw), ctypes.byref(res),
                                 ctypes.byref(n)), "hqpcu_solve")
        k = self.batch * self.m
        return dx, dy, dz[:k], dw[:k], res.value, n.value

    def residuum(self, r1, r2, r3, r4, dx, dy, dz, dw):
        a = [np.ascontiguousarray(v, np.float64) for v in (r1, r2, r3, r4, dx, dy, dz, dw)]
        res = ctypes.c_double(0)
        _check(lib().hqpcu_residuum(self.h, *[_hp(v) for v in a], ctypes.byref(res)),
               "hqpcu_residuum")
        return res.value

    # ---- SQP-level vector operations -----------------
    def sqp_grd_L(self, c, y, z):
        """c - A'y - C'z (Hqp_SqpSolver::grd_L, hqp/Hqp_SqpSolver.C:430-445) from the QP
        values on the device."""
        a = [np.ascontiguousarray(v, np.float64) for v in (c, y, z)]
        out = np.zeros(self.N)
        _check(lib().hqpcu_sqp_grd_L(self.h, *[_hp(v) for v in a], _hp(out)), "hqpcu_sqp_grd_L")
        return out

    def sqp_merit(self, f, c, s, b, d, re, r):
        """[phi, phi1, s'Qs, c's, ||b||inf, ||min(d,0)||inf, sum re|As+b|, -sum r min(0,Cs+d)]
        (Hqp_SqpPowell::phi / ::phi1, hqp/Hqp_SqpPowell.C:189-244; Hqp_SqpSolver::norm_inf
        and s'Qs, hqp/Hqp_SqpSolver.C:155-174, 299-301)."""
        a = [np.ascontiguousarray(v, np.float64) for v in (c, s, b, d, re, r)]
        out = np.zeros(8)
        _check(lib().hqpcu_sqp_merit(self.h, ctypes.c_double(f), *[_hp(v) for v in a], _hp(out)),
               "hqpcu_sqp_merit")
        return out

    def sqp_quad(self, x):
        """x'Qx (hqp/Hqp_SqpSolver.C:225-226)."""
        x = np.ascontiguousarray(x, np.float64)
        out = ctypes.c_double(0)
        _check(lib().hqpcu_sqp_quad(self.h, _hp(x), ctypes.byref(out)), "hqpcu_sqp_quad")
        return out.value

    def mehrotra_solve(self, c=None, b=None, d=None, eps=1e-9, max_iters=0, hot=None,
                       max_warm_iters=0):
        """Hqp_IpsMehrotra cold_start + solve on the device (hqpcu_mehrotra_solve);
        hot = (x, y) of the previous solve: hot_start + solve
        (hqpcu_mehrotra_hot_solve, z and w continue on the device)."""
        p = self.prob
        c = np.ascontiguousarray(p.c if c is None else c, np.float64)
        b = np.ascontiguousarray(p.b if b is None else b, np.float64)
        d = np.ascontiguousarray(p.d if d is None else d, np.float64)
        x, y = np.zeros(self.N), np.zeros(self.me)
        z, w = np.zeros(max(self.m, 1)), np.zeros(max(self.m, 1))
        it, res, gap = ctypes.c_int(0), ctypes.c_int(0), ctypes.c_double(0)
        if hot is not None:
            x[:], y[:] = hot[0], hot[1]
            _check(lib().hqpcu_mehrotra_hot_solve(self.h, _hp(c), _hp(b), _hp(d),
                                                  ctypes.c_double(eps), max_iters, max_warm_iters,
                                                  _hp(x), _hp(y), _hp(z), _hp(w), ctypes.byref(it),
                                                  ctypes.byref(res), ctypes.byref(gap)),
                   "hqpcu_mehrotra_hot_solve")
        else:
            _check(lib().hqpcu_mehrotra_solve(self.h, _hp(c), _hp(b), _hp(d), ctypes.c_double(eps),
                                              max_iters, _hp(x), _hp(y), _hp(z), _hp(w),
                                              ctypes.byref(it), ctypes.byref(res), ctypes.byref(gap)),
                   "hqpcu_mehrotra_solve")
        names = ["optimal", "feasible", "infeasible", "suboptimal", "degenerate"]
        return dict(x=x, y=y, z=z[:self.m], w=w[:self.m], iters=it.value,
                    result=names[res.value], gap=gap.value)

    def franke_solve(self, c=None, b=None, d=None, eps=1e-9, max_iters=0, hot=None,
                     max_warm_iters=0, beta=0.0, mu0=0.0):
        """Hqp_IpsFranke cold_start (or hot_start from hot = (x, y, z, w)) + solve on
        the device (hqpcu_franke_solve)."""
        p = self.prob
        c = np.ascontiguousarray(p.c if c is None else c, np.float64)
        b = np.ascontiguousarray(p.b if b is None else b, np.float64)
        d = np.ascontiguousarray(p.d if d is None else d, np.float64)
        x, y = np.zeros(self.N), np.zeros(self.me)
        z, w = np.zeros(max(self.m, 1)), np.zeros(max(self.m, 1))
        if hot is not None:
            x[:], y[:] = hot[0], hot[1]
            z[:self.m], w[:self.m] = hot[2], hot[3]
        it, res, gap = ctypes.c_int(0), ctypes.c_int(0), ctypes.c_double(0)
        _check(lib().hqpcu_franke_solve(self.h, _hp(c), _hp(b), _hp(d), ctypes.c_double(eps),
                                        max_iters, int(hot is not None), max_warm_iters,
                                        ctypes.c_double(beta), ctypes.c_double(mu0), _hp(x), _hp(y),
                                        _hp(z), _hp(w), ctypes.byref(it), ctypes.byref(res),
                                        ctypes.byref(gap)), "hqpcu_franke_solve")
        names = ["optimal", "feasible", "infeasible", "suboptimal", "degenerate"]
        return dict(x=x, y=y, z=z[:self.m], w=w[:self.m], iters=it.value,
                    result=names[res.value], gap=gap.value)

    def set_value_map(self, dst, dst2):
        dst = np.ascontiguousarray(dst, np.int64)
        dst2 = np.ascontiguousarray(dst2, np.int64)
        _check(lib().hqpcu_set_value_map(self.h, ctypes.c_longlong(len(dst)),
                                         dst.ctypes.data_as(ctypes.c_void_p),
                                         dst2.ctypes.data_as(ctypes.c_void_p)),
               "hqpcu_set_value_map")

    def update_values(self, vals, ineq_val=None, eq_val=None):
        p = self.prob
        iv = np.ascontiguousarray(p.ineq_val if ineq_val is None else ineq_val, np.float64)
        ev = np.ascontiguousarray(p.eq_val if eq_val is None else eq_val, np.float64)
        vals = np.ascontiguousarray(vals, np.float64)
        _check(lib().hqpcu_update_values(self.h, _hp(vals), _hp(iv), _hp(ev)), "hqpcu_update_values")

    def solve_stats(self):
        a, b = ctypes.c_longlong(0), ctypes.c_longlong(0)
        _check(lib().hqpcu_solve_stats(self.h, ctypes.byref(a), ctypes.byref(b)), "hqpcu_solve_stats")
        return a.value, b.value

    def get_factor(self):
        p, B = self.prob, self.batch
        V = np.zeros((B, p.K + 1, p.nx, p.nx))
        R = np.zeros((B, p.K, p.nu, p.nx))
        _check(lib().hqpcu_get_factor(self.h, _hp(V), _hp(R)), "hqpcu_get_factor")
        return V, R

    # -- device-pointer API (inputs already resident in HBM) -------------------
    def update_dev(self, Q, fx, fu, ineq_val):
        _check(lib().hqpcu_update_dev(self.h, _vp(Q), _vp(fx), _vp(fu), _vp(ineq_val), None),
               "hqpcu_update_dev")

    def update_stages_dev(self, k0, nq, Q, nf, fx, fu):
        _check(lib().hqpcu_update_stages_dev(self.h, int(k0), int(nq), _vp(Q), int(nf), _vp(fx),
                                             _vp(fu)), "hqpcu_update_stages_dev")

    def update_ineq_dev(self, ineq_val):
        _check(lib().hqpcu_update_ineq_dev(self.h, _vp(ineq_val)), "hqpcu_update_ineq_dev")

    def factor_dev(self, z, w):
        _check(lib().hqpcu_factor_dev(self.h, _vp(z), _vp(w)), "hqpcu_factor_dev")

    def step_dev(self, r1, r2, r3, r4, dx, dy, dz, dw):
        _check(lib().hqpcu_step_dev(self.h, *[_vp(v) for v in (r1, r2, r3, r4, dx, dy, dz, dw)]),
               "hqpcu_step_dev")

    def solve_dev(self, r1, r2, r3, r4, dx, dy, dz, dw, eps=1e-10):
        res, n = ctypes.c_double(0), ctypes.c_int(0)
        _check(lib().hqpcu_solve_dev(self.h, ctypes.c_double(eps),
                                     *[_vp(v) for v in (r1, r2, r3, r4, dx, dy, dz, dw)],
                                     ctypes.byref(res), ctypes.byref(n)), "hqpcu_solve_dev")
        return res.value, n.value
