"""ctypes wrapper around oracle/_ref/libhqpharness.so (the UNMODIFIED reference).

TEST INFRASTRUCTURE ONLY.  Importable from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs -- never from hqp_b200/.
The shared objects are prebuilt in the build container by `make -C oracle ref`
(they travel to the GPU box with the snapshot; /root/reference does not).
"""
from __future__ import annotations

import ctypes
import json
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(_HERE, "_ref")
_LIB = None

HQP_RESULT = ["optimal", "feasible", "infeasible", "suboptimal", "degenerate"]


def available() -> bool:
    return os.path.exists(os.path.join(REF_DIR, "libhqpharness.so"))


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(REF_DIR, "libhqpharness.so")
        if not os.path.exists(path):
            raise RuntimeError(f"{path} missing: run `make -C oracle ref` where "
                               "/root/reference exists")
        _LIB = ctypes.CDLL(path, mode=ctypes.RTLD_GLOBAL)
        _LIB.ref_qp_create.restype = ctypes.c_void_p
        _LIB.ref_mat_create.restype = ctypes.c_void_p
        _LIB.ref_last_error.restype = ctypes.c_char_p
        _LIB.ref_get_real.restype = ctypes.c_double
        _LIB.ref_init()
    return _LIB


def _dp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double)) if a is not None else None


def _ip(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int))


class RefQP:
    """An Hqp_Program built from an hqp_b200.problem.LQProblem."""

    def __init__(self, prob):
        self.prob = prob
        L = lib()
        self._keep = (prob.csr_Q_upper(), prob.csr_A(), prob.csr_C())
        (qp, qj, qv), (ap, aj, av), (cp, cj, cv) = self._keep
        self.n, self.me, self.m = prob.N, prob.me, prob.m
        d = prob.d if prob.m else np.zeros(1)
        self.h = ctypes.c_void_p(L.ref_qp_create(
            self.n, self.me, self.m, _ip(qp), _ip(qj), _dp(qv), _dp(prob.c),
            _ip(ap), _ip(aj), _dp(av), _dp(prob.b), _ip(cp), _ip(cj), _dp(cv), _dp(d)))

    def close(self):
        if self.h:
            lib().ref_qp_free(self.h)
            self.h = None


class RefMatrix:
    """A reference Hqp_IpMatrix plugin ("LQDOCP", "RedSpBKP", "SpBKP", ...)."""

    def __init__(self, name, qp: RefQP):
        L = lib()
        self.qp = qp
        self.h = ctypes.c_void_p(L.ref_mat_create(name.encode()))
        if not self.h:
            raise RuntimeError(f"unknown matrix module {name}")
        self._check(L.ref_mat_init(self.h, qp.h), "init")

    @staticmethod
    def _check(err, what):
        if err:
            raise ArithmeticError(f"reference raised Meschach error {err} in {what}")

    def update(self):
        self._check(lib().ref_mat_update(self.h, self.qp.h), "update")

    def factor(self, z, w):
        z = np.ascontiguousarray(z, np.float64)
        w = np.ascontiguousarray(w, np.float64)
        self._check(lib().ref_mat_factor(self.h, self.qp.h, len(z), _dp(z), _dp(w)), "factor")

    def _apply(self, mode, z, w, r1, r2, r3, r4, sol=None):
        q = self.qp
        if sol is None:
            dx, dy, dz, dw = np.zeros(q.n), np.zeros(q.me), np.zeros(max(q.m, 1)), np.zeros(max(q.m, 1))
        else:
            dx, dy, dz, dw = (np.ascontiguousarray(a, np.float64).copy() for a in sol)
        res = ctypes.c_double(0)
        args = [np.ascontiguousarray(a, np.float64) for a in (z, w, r1, r2, r3, r4)]
        err = lib().ref_mat_apply(self.h, q.h, mode, *[_dp(a) for a in args],
                                  _dp(dx), _dp(dy), _dp(dz), _dp(dw), ctypes.byref(res))
        self._check(err, "step/solve")
        return dx, dy, dz[:q.m], dw[:q.m], res.value

    def step(self, z, w, r1, r2, r3, r4):
        return self._apply(0, z, w, r1, r2, r3, r4)[:4]

    def solve(self, z, w, r1, r2, r3, r4):
        return self._apply(1, z, w, r1, r2, r3, r4)

    def residuum(self, z, w, r1, r2, r3, r4, dx, dy, dz, dw):
        return self._apply(2, z, w, r1, r2, r3, r4, (dx, dy, dz, dw))[4]

    def time(self, z, w, r1, r2, r3, r4, reps=3, nstep=2):
        tf, ts = ctypes.c_double(0), ctypes.c_double(0)
        args = [np.ascontiguousarray(a, np.float64) for a in (z, w, r1, r2, r3, r4)]
        err = lib().ref_mat_time(self.h, self.qp.h, *[_dp(a) for a in args],
                                 reps, nstep, ctypes.byref(tf), ctypes.byref(ts))
        self._check(err, "time")
        return tf.value, ts.value

    def close(self):
        if self.h:
            lib().ref_mat_free(self.h)
            self.h = None


def ips_solve(qp: RefQP, solver="Mehrotra", mat="LQDOCP", eps=1e-9, max_iters=0):
    """Cold-started IP solve; returns dict(x,y,z,iters,result,seconds)."""
    x, y, z = np.zeros(qp.n), np.zeros(qp.me), np.zeros(max(qp.m, 1))
    it, res, sec = ctypes.c_int(0), ctypes.c_int(0), ctypes.c_double(0)
    err = lib().ref_ips_solve(qp.h, solver.encode(), mat.encode(), ctypes.c_double(eps),
                              max_iters, _dp(x), _dp(y), _dp(z), ctypes.byref(it),
                              ctypes.byref(res), ctypes.byref(sec))
    if err:
        raise ArithmeticError(f"reference IP solve failed with code {err}")
    return dict(x=x, y=y, z=z[:qp.m], iters=it.value, result=HQP_RESULT[res.value],
                seconds=sec.value)


def ips_solve_seq(qp: RefQP, cs, bs, ds, solver="Mehrotra", mat="LQDOCP", eps=1e-9, max_iters=0):
    """First solve cold-started, the others hot-started with the linear terms
    c, b, d replaced by the rows of cs, bs, ds; -> list of dict(x,y,z,iters,result)."""
    cs, bs, ds = (np.ascontiguousarray(a, np.float64) for a in (cs, bs, ds))
    ns = cs.shape[0]
    xs, ys = np.zeros((ns, qp.n)), np.zeros((ns, qp.me))
    zs = np.zeros((ns, max(qp.m, 1)))
    it = (ctypes.c_int * ns)()
    res = (ctypes.c_int * ns)()
    err = lib().ref_ips_solve_seq(qp.h, solver.encode(), mat.encode(), ctypes.c_double(eps),
                                  max_iters, ns, _dp(cs), _dp(bs), _dp(ds), _dp(xs), _dp(ys),
                                  _dp(zs), it, res)
    if err:
        raise ArithmeticError(f"reference IP solve sequence failed with code {err}")
    return [dict(x=xs[k], y=ys[k], z=zs[k][:qp.m], iters=it[k], result=HQP_RESULT[res[k]])
            for k in range(ns)]


def load_plugin(path):
    if lib().ref_load_plugin(os.fsencode(path)):
        raise RuntimeError(lib().ref_last_error().decode())


def docp_did(kmax=60, qp_solver="", mat_solver="LQDOCP", plugin="", with_cns=1, env=None):
    """Run the hqp_docp example in a subprocess (global solver state)."""
    exe = os.path.join(REF_DIR, "docp_ref")
    out = subprocess.run([exe, str(kmax), qp_solver, mat_solver, plugin, str(with_cns), "0"],
                         capture_output=True, text=True, timeout=600, env=env)
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    if not lines:
        raise RuntimeError(f"docp_ref failed: rc={out.returncode}\n{out.stdout}\n{out.stderr}")
    return json.loads(lines[-1])


def hl_bfgs_block(Q, s, u, alpha, gamma=0.1, eps=1e-8, eigen_control=True):
    """The UNMODIFIED Hqp_HL_BFGS::update_b_Q (hqp/Hqp_HL_BFGS.C:149-213) on one dense
    block (both triangles); returns the updated block."""
    L = lib()
    Q = np.array(Q, dtype=np.float64, order="C", copy=True)
    n = Q.shape[0]
    s = np.ascontiguousarray(s, dtype=np.float64)
    u = np.ascontiguousarray(u, dtype=np.float64)
    L.ref_hl_bfgs_block.restype = ctypes.c_int
    rc = L.ref_hl_bfgs_block(ctypes.c_int(n), _dp(Q), _dp(s), _dp(u), ctypes.c_double(alpha),
                             ctypes.c_double(gamma), ctypes.c_double(eps), ctypes.c_int(1 if eigen_control else 0))
    if rc:
        raise RuntimeError("ref_hl_bfgs_block failed")
    return Q


def hl_update(qp: "RefQP", hela, s, u, alpha, gamma=0.1, eps=1e-8, eigen_control=True):
    """Hqp_HL::setup + ::update of the module `hela` ("BFGS", or "CudaBFGS" after
    load_plugin) on the QP's Hessian, in place."""
    L = lib()
    s = np.ascontiguousarray(s, dtype=np.float64)
    u = np.ascontiguousarray(u, dtype=np.float64)
    L.ref_hl_update.restype = ctypes.c_int
    rc = L.ref_hl_update(hela.encode(), qp.h, _dp(s), _dp(u), ctypes.c_double(alpha),
                         ctypes.c_double(gamma), ctypes.c_double(eps), ctypes.c_int(1 if eigen_control else 0))
    if rc:
        raise RuntimeError(f"ref_hl_update({hela}) failed: {rc}")


def qp_get_Q_block(qp: "RefQP", offs, size):
    out = np.zeros((size, size))
    lib().ref_qp_get_Q_block(qp.h, ctypes.c_int(offs), ctypes.c_int(size), _dp(out))
    return out


def sqp_eval(qp: "RefQP", f, s, y, z, re, r):
    """Row f3 on the reference: (grd_L, out8) with out8 = [phi, phi1, s'Qs, c's, norm_inf, ...]
    from the UNMODIFIED Hqp_SqpSolver::grd_L / ::norm_inf and Hqp_SqpPowell::phi / ::phi1
    (hqp/Hqp_SqpSolver.C:155-174, 430-445; hqp/Hqp_SqpPowell.C:189-244) on this QP with
    qp->x = s."""
    L = lib()
    a = [np.ascontiguousarray(v, dtype=np.float64) for v in (s, y, z, re, r)]
    grd = np.zeros(qp.n)
    out8 = np.zeros(8)
    L.ref_sqp_eval.restype = ctypes.c_int
    rc = L.ref_sqp_eval(qp.h, ctypes.c_double(f), *[_dp(v) if v.size else None for v in a], _dp(grd), _dp(out8))
    if rc:
        raise RuntimeError("ref_sqp_eval failed")
    return grd, out8


class RefDocp:
    """Row f4: the UNMODIFIED Hqp_Docp (setup / update / update_fbd / update_grds /
    update_bounds) driven for the reference's Prg_DID (model 0) or for the synthetic model
    restated as an Hqp_Docp subclass in oracle/prg_synthnl.cpp (model 1).  prob: an
    hqp_b200.docpcuda.DocpProblem (dimensions, parameters, start values)."""

    def __init__(self, prob, cuda=False):
        """cuda: the same program with its stage loop on the GPU -- Prg_DIDCuda from the plugin
        library / Prg_SynthNLCuda (the host mix-in hqp_b200/host/Hqp_DocpCuda.h under test)."""
        L = lib()
        L.ref_docp_create.restype = ctypes.c_void_p
        if cuda and int(prob.model) == 0:
            from hqp_b200 import build
            path = os.path.join(build.LIB, "libhqp_ipcuda_plugin.so")
            if L.ref_load_plugin(path.encode()):
                raise RuntimeError(f"cannot load {path}: {L.ref_last_error().decode()}")
        par = np.ascontiguousarray(prob.par, dtype=np.float64)
        spar = np.ascontiguousarray(prob.spar, dtype=np.float64)
        xin = np.ascontiguousarray(prob.x_init, dtype=np.float64)
        nspar = spar.shape[1] if spar.ndim == 2 else 0
        self.h = ctypes.c_void_p(L.ref_docp_create(
            int(prob.model) + (2 if cuda else 0), prob.K, prob.nx, prob.nu, prob.nc, prob.ncK, _dp(par), par.size,
            _dp(spar) if spar.size else None, nspar, _dp(xin)))
        if not self.h:
            raise RuntimeError("ref_docp_create failed")
        n, me, m = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        L.ref_docp_sizes(self.h, ctypes.byref(n), ctypes.byref(me), ctypes.byref(m))
        self.N, self.me, self.m = n.value, me.value, m.value

    def x(self):
        out = np.empty(self.N)
        lib().ref_docp_get_x(self.h, _dp(out))
        return out

    def solve(self, sqp_eps=1e-6, simulate=True, qp_solver="", mat_solver=""):
        """hqp_solve on this program (it must be the one created last): the calls of
        hqp_docp/Docp_Main.C:68-76.  Returns dict result, objective, sqp_iters, qp_iters, x."""
        obj = ctypes.c_double()
        it, qit = ctypes.c_int(), ctypes.c_int()
        res = ctypes.create_string_buffer(256)
        rc = lib().ref_prg_solve(ctypes.c_double(sqp_eps), int(bool(simulate)), qp_solver.encode(),
                                 mat_solver.encode(), ctypes.byref(obj), ctypes.byref(it), ctypes.byref(qit),
                                 res, 256)
        if rc:
            raise RuntimeError(f"ref_prg_solve: {rc}")
        return dict(result=res.value.decode(), objective=obj.value, sqp_iters=it.value, qp_iters=qit.value,
                    x=self.x())

    def update(self, x=None, fbd_only=False, matrices=True):
        """Returns dict f, b, d, c (qp->c) and -- unless fbd_only -- dense A [me,N], C [m,N]."""
        f = ctypes.c_double()
        b, d, c = np.empty(self.me), np.empty(max(1, self.m)), np.empty(self.N)
        want = matrices and not fbd_only
        A = np.empty((self.me, self.N)) if want else None
        C = np.empty((max(1, self.m), self.N)) if want else None
        xa = np.ascontiguousarray(x, dtype=np.float64) if x is not None else None
        rc = lib().ref_docp_update(self.h, _dp(xa), int(bool(fbd_only)), ctypes.byref(f), _dp(b), _dp(d),
                                   _dp(c), _dp(A), _dp(C))
        if rc:
            raise RuntimeError(f"ref_docp_update: {rc}")
        return dict(f=f.value, b=b, d=d[:self.m], c=c, A=A, C=C[:self.m] if C is not None else None)

    def close(self):
        if self.h:
            lib().ref_docp_free(self.h)
            self.h = None
