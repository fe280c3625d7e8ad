"""ctypes mirror of include/hqp_docpcuda.h (libhqpdocp.so): the stage loop of
Hqp_Docp::update / ::update_fbd on the GPU for device-resident stage models (SURVEY.md
section 8, row f4; reference: hqp/Hqp_Docp.C:831-891, 944-1075, 1097-1180, 893-940).
No CPU fallback.

Also here, because a host module needs it to fill hqpdocp_dims: DocpProblem, the host-side
description of a discrete-time optimal control program with uniform stages -- bounds per stage
parsed into the six association tables the way Hqp_Docp::setup_x / parse_constr do
(hqp/Hqp_Docp.C:370-397, 440-583)."""
from __future__ import annotations

import ctypes
import os
from dataclasses import dataclass, field

import numpy as np

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libhqpdocp.so")
_LIB = None

MODEL_DID, MODEL_SYNTHNL = 0, 1
GRAD_FD, GRAD_AD = 0, 1
INF = float("inf")


class _Assoc(ctypes.Structure):
    _fields_ = [("dim", ctypes.c_int), ("idxs", ctypes.POINTER(ctypes.c_int)),
                ("vals", ctypes.POINTER(ctypes.c_double))]


class _Dims(ctypes.Structure):
    _fields_ = [("K", ctypes.c_int), ("nx", ctypes.c_int), ("nu", ctypes.c_int), ("nc", ctypes.c_int),
                ("ncK", ctypes.c_int), ("model", ctypes.c_int), ("npar", ctypes.c_int),
                ("par", ctypes.POINTER(ctypes.c_double)), ("nspar", ctypes.c_int),
                ("spar", ctypes.POINTER(ctypes.c_double)),
                ("xu_eq", _Assoc), ("xu_lb", _Assoc), ("xu_ub", _Assoc),
                ("cns_eq", _Assoc), ("cns_lb", _Assoc), ("cns_ub", _Assoc), ("device", ctypes.c_int),
                ("k_first", ctypes.c_int), ("K_total", ctypes.c_int)]


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            from . import build
            build.build_docp()
        _LIB = ctypes.CDLL(LIB_PATH)
        _LIB.hqpdocp_last_error.restype = ctypes.c_char_p
        _LIB.hqpdocp_launch_count.restype = ctypes.c_longlong
        _LIB.hqpdocp_launch_count.argtypes = [ctypes.c_void_p]
    return _LIB


def _dp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def _check(rc, what):
    if rc:
        raise RuntimeError(f"{what}: status {rc}: {lib().hqpdocp_last_error().decode()}")


@dataclass
class Assoc:
    """Hqp_DocpAssoc (hqp/Hqp_Docp.C:36-112): bound values and the index each belongs to."""
    idxs: list = field(default_factory=list)
    vals: list = field(default_factory=list)

    def arrays(self):
        return (np.ascontiguousarray(self.idxs, dtype=np.int32),
                np.ascontiguousarray(self.vals, dtype=np.float64))


def parse_constr(cmin, cmax, idx, eq: Assoc, lb: Assoc, ub: Assoc):
    """Hqp_Docp::parse_constr, hqp/Hqp_Docp.C:370-397."""
    for lo, hi in zip(cmin, cmax):
        if lo == hi:
            if not np.isfinite(lo):
                raise ValueError("parse_constr: equal infinite bounds (a Periodical state) are not supported")
            eq.idxs.append(idx)
            eq.vals.append(lo)
        else:
            if lo > -INF:
                lb.idxs.append(idx)
                lb.vals.append(lo)
            if hi < INF:
                ub.idxs.append(idx)
                ub.vals.append(hi)
        idx += 1


@dataclass
class DocpProblem:
    """Uniform-stage DOCP: K stages with (nx, nu, nc), final stage with (nx, 0, ncK)."""
    model: int
    K: int
    nx: int
    nu: int
    nc: int
    ncK: int
    par: np.ndarray
    spar: np.ndarray            # [(K+1), nspar]
    x_init: np.ndarray          # [N]
    xu_eq: Assoc = field(default_factory=Assoc)
    xu_lb: Assoc = field(default_factory=Assoc)
    xu_ub: Assoc = field(default_factory=Assoc)
    cns_eq: Assoc = field(default_factory=Assoc)
    cns_lb: Assoc = field(default_factory=Assoc)
    cns_ub: Assoc = field(default_factory=Assoc)
    k_first: int = 0            # a stage range of a longer horizon (hqpdocp_dims.k_first / K_total)
    K_total: int = 0            # 0: the whole horizon

    @property
    def owns_final(self):
        """False: local stage K belongs to the next range, only its x is read (a halo)."""
        return self.K_total == 0 or self.k_first + self.K == self.K_total

    @property
    def nd(self):
        return self.nx + self.nu

    @property
    def N(self):
        return self.K * self.nd + self.nx

    @property
    def ncns(self):
        return self.K * self.nc + (self.ncK if self.owns_final else 0)

    @property
    def me(self):
        return self.K * self.nx + len(self.xu_eq.idxs) + len(self.cns_eq.idxs)

    @property
    def m(self):
        return len(self.xu_lb.idxs) + len(self.xu_ub.idxs) + len(self.cns_lb.idxs) + len(self.cns_ub.idxs)

    def set_bounds(self, stage_bounds):
        """stage_bounds(k) -> (x_min, x_max, u_min, u_max, c_min, c_max); the walk of
        Hqp_Docp::setup_x, hqp/Hqp_Docp.C:465-541 (x, then u, then the constraints of a stage)."""
        for t in (self.xu_eq, self.xu_lb, self.xu_ub, self.cns_eq, self.cns_lb, self.cns_ub):
            t.idxs.clear()
            t.vals.clear()
        for k in range(self.K + 1 if self.owns_final else self.K):
            x_min, x_max, u_min, u_max, c_min, c_max = stage_bounds(self.k_first + k)
            assert len(x_min) == self.nx and len(u_min) == (self.nu if k < self.K else 0)
            assert len(c_min) == (self.nc if k < self.K else self.ncK)
            # (indices below are LOCAL to the range)
            parse_constr(x_min, x_max, k * self.nd, self.xu_eq, self.xu_lb, self.xu_ub)
            parse_constr(u_min, u_max, k * self.nd + self.nx, self.xu_eq, self.xu_lb, self.xu_ub)
            parse_constr(c_min, c_max, k * self.nc, self.cns_eq, self.cns_lb, self.cns_ub)
        self._stage_bounds = stage_bounds
        return self

    def shard(self, rank, world):
        """The contiguous stage range of rank `rank` out of `world` (the split hqp_b200/dist.py
        uses for the KKT horizon): a DocpProblem with local indices, x_init = the range's states
        and controls plus the next state (halo).  Results of the ranges concatenate group by
        group -- b: dynamics rows | x/u fixings | constraint equalities; d: the four bound groups
        -- and the objective adds up (assemble_shards)."""
        assert self.K_total == 0 and 0 <= rank < world <= self.K
        base, rem = divmod(self.K, world)  # (hqp_b200/dist.py:stage_ranges)
        k0 = rank * base + min(rank, rem)
        k1 = k0 + base + (1 if rank < rem else 0)
        q = DocpProblem(self.model, k1 - k0, self.nx, self.nu, self.nc, self.ncK, self.par,
                        np.ascontiguousarray(self.spar[k0:k1 + 1]), self.x_slice(self.x_init, k0, k1),
                        k_first=k0, K_total=self.K)
        return q.set_bounds(self._stage_bounds)

    def x_slice(self, x, k0, k1):
        """[x_k0 u_k0 ... x_k1] of a full-horizon vector."""
        return np.ascontiguousarray(x[k0 * self.nd:k1 * self.nd + self.nx])


def assemble_shards(full: "DocpProblem", shards, outs):
    """Full-horizon results from the per-range results (dicts as DocpCuda.update returns, or
    (f, b, d) tuples of update_fbd) of shards = [full.shard(r, world) ...]."""
    if isinstance(outs[0], tuple):
        outs = [dict(f=o[0], b=o[1], d=o[2]) for o in outs]
    res = {"f": float(sum(o["f"] for o in outs))}
    nb = [(q.K * q.nx, len(q.xu_eq.idxs), len(q.cns_eq.idxs)) for q in shards]
    nd_ = [(len(q.xu_lb.idxs), len(q.xu_ub.idxs), len(q.cns_lb.idxs), len(q.cns_ub.idxs)) for q in shards]

    def groups(key, sizes):
        cols = []
        for g in range(len(sizes[0])):
            for o, sz in zip(outs, sizes):
                off = sum(sz[:g])
                cols.append(o[key][off:off + sz[g]])
        return np.concatenate(cols) if cols else np.zeros(0)

    res["b"], res["d"] = groups("b", nb), groups("d", nd_)
    if "g" in outs[0]:
        gs = []
        for q, o in zip(shards, outs):  # the halo state's gradient entries belong to the next range
            gs.append(o["g"] if q.owns_final else o["g"][:q.K * q.nd])
        res["g"] = np.concatenate(gs)
        for key in ("fx", "fu", "cu"):
            res[key] = np.concatenate([o[key] for o in outs])
        res["cx"] = np.concatenate([o["cx"][:q.ncns] for q, o in zip(shards, outs)])
    return res


def did_problem(kmax=60, with_cns=True) -> DocpProblem:
    """The reference's example program: bounds and start values of Prg_DID::setup_vars
    (hqp_docp/Prg_DID.C:33-74), dt = 1/kmax."""
    K = kmax
    x_init = np.zeros(K * 3 + 2)
    x_init[0] = 1.0
    x_init[2:K * 3:3] = -2.0
    p = DocpProblem(MODEL_DID, K, 2, 1, 1 if with_cns else 0, 0, np.array([1.0 / kmax]),
                    np.zeros((K + 1, 0)), x_init)

    def bounds(k):
        x_min, x_max = [-INF, -INF], [INF, INF]
        if k == 0:
            x_min, x_max = [1.0, 0.0], [1.0, 0.0]
        elif k < K:
            x_max[1] = 0.01
        else:
            x_min, x_max = [-1.0, 0.0], [-1.0, 0.0]
        u = ([-INF], [INF]) if k < K else ([], [])
        c = ([-INF], [0.01]) if (with_cns and k < K) else ([], [])
        return x_min, x_max, u[0], u[1], c[0], c[1]

    return p.set_bounds(bounds)


def synthnl_problem(K, nx, nu, nc=1, ncK=0, eps=0.1, seed=1234) -> DocpProblem:
    """The synthetic SQP-driven workload (docp_models.cuh, ModelSynthNL): dynamics matrix
    I + 0.1 U/sqrt(nx) (spectral radius ~ 1), input matrix U(-1,1),
    positive weights, a slowly varying tracking reference; fixed initial state, box bounds on u,
    two state bounds per interior stage, constraint bounds cycling through ub / lb+ub / eq."""
    rng = np.random.default_rng(seed)
    A = np.eye(nx) + 0.1 * rng.uniform(-1, 1, (nx, nx)) / np.sqrt(nx)
    B = rng.uniform(-1, 1, (nx, nu))
    qw = rng.uniform(0.5, 1.5, nx)
    rw = rng.uniform(0.5, 1.5, nu)
    par = np.concatenate([[eps], A.ravel(), B.ravel(), qw, rw])
    t = np.arange(K + 1)[:, None] / max(1, K)
    spar = 0.5 * np.sin(2 * np.pi * (t + rng.uniform(0, 1, (1, nx))))
    nd = nx + nu
    x_init = rng.uniform(-1, 1, K * nd + nx)
    p = DocpProblem(MODEL_SYNTHNL, K, nx, nu, nc, ncK, par, np.ascontiguousarray(spar), x_init)

    def bounds(k):
        x_min, x_max = [-INF] * nx, [INF] * nx
        if k == 0:
            x_min = x_max = list(x_init[:nx])
        elif k < K:
            x_max[1 if nx > 1 else 0] = 2.0
            x_min[0] = -3.0
        if k < K:
            u_min, u_max = [-1.0] * nu, [1.0] * nu
            c_min, c_max = [-INF] * nc, [INF] * nc
            for i in range(nc):
                if i % 3 == 0:
                    c_max[i] = 2.0
                elif i % 3 == 1:
                    c_min[i], c_max[i] = -1.0, 1.0
                else:
                    c_min[i] = c_max[i] = 0.1
        else:
            u_min, u_max = [], []
            c_min, c_max = [-1.0] * ncK, [INF] * ncK
        return x_min, x_max, u_min, u_max, c_min, c_max

    return p.set_bounds(bounds)


class DocpCuda:
    """One handle of libhqpdocp.so for a DocpProblem."""

    def __init__(self, prob: DocpProblem, device=0):
        self.p = prob
        self._keep = []

        def assoc(t):
            idxs, vals = t.arrays()
            self._keep += [idxs, vals]
            return _Assoc(len(idxs), idxs.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), _dp(vals))

        par = np.ascontiguousarray(prob.par, dtype=np.float64)
        spar = np.ascontiguousarray(prob.spar, dtype=np.float64)
        self._keep += [par, spar]
        d = _Dims(prob.K, prob.nx, prob.nu, prob.nc, prob.ncK, prob.model, par.size, _dp(par),
                  spar.shape[1] if spar.ndim == 2 else 0, _dp(spar) if spar.size else None,
                  assoc(prob.xu_eq), assoc(prob.xu_lb), assoc(prob.xu_ub),
                  assoc(prob.cns_eq), assoc(prob.cns_lb), assoc(prob.cns_ub), device,
                  prob.k_first, prob.K_total)
        self.h = ctypes.c_void_p()
        _check(lib().hqpdocp_create(ctypes.byref(d), ctypes.byref(self.h)), "hqpdocp_create")

    def close(self):
        if self.h:
            lib().hqpdocp_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def launches(self):
        return int(lib().hqpdocp_launch_count(self.h))

    def set_stream(self, stream):
        _check(lib().hqpdocp_set_stream(self.h, ctypes.c_void_p(stream)), "hqpdocp_set_stream")

    def update_fbd(self, x):
        """Hqp_Docp::update_fbd: returns (f, b, d)."""
        p = self.p
        x = np.ascontiguousarray(x, dtype=np.float64)
        assert x.size == p.N
        f = ctypes.c_double(0.0)
        b, d = np.empty(p.me), np.empty(max(1, p.m))
        _check(lib().hqpdocp_update_fbd(self.h, _dp(x), ctypes.byref(f), _dp(b), _dp(d)), "hqpdocp_update_fbd")
        return f.value, b, d[:p.m]

    def update(self, x, grad_mode=GRAD_FD):
        """Hqp_Docp::update: returns a dict f, b, d, g, fx [K,nx,nx], fu [K,nx,nu],
        cx [ncns,nx], cu [K*nc,nu]."""
        p = self.p
        x = np.ascontiguousarray(x, dtype=np.float64)
        assert x.size == p.N
        f = ctypes.c_double(0.0)
        b, d, g = np.empty(p.me), np.empty(max(1, p.m)), np.empty(p.N)
        fx, fu = np.empty((p.K, p.nx, p.nx)), np.empty((p.K, p.nx, max(1, p.nu)))
        cx, cu = np.empty((max(1, p.ncns), p.nx)), np.empty((max(1, p.K * p.nc), max(1, p.nu)))
        _check(lib().hqpdocp_update(self.h, ctypes.c_int(grad_mode), _dp(x), ctypes.byref(f), _dp(b), _dp(d),
                                    _dp(g), _dp(fx), _dp(fu), _dp(cx), _dp(cu)), "hqpdocp_update")
        fu = fu.ravel()[:p.K * p.nx * p.nu].reshape(p.K, p.nx, p.nu)
        cx = cx.ravel()[:p.ncns * p.nx].reshape(p.ncns, p.nx)
        cu = cu.ravel()[:p.K * p.nc * p.nu].reshape(p.K * p.nc, p.nu)
        return dict(f=f.value, b=b, d=d[:p.m], g=g, fx=fx, fu=fu, cx=cx, cu=cu)

    def update_dev(self, x, f, b, d, g, fx, fu, cx, cu, grad_mode=GRAD_AD):
        """Device-resident variant on torch CUDA float64 tensors, asynchronous on the stream."""
        ptr = lambda t: ctypes.c_void_p(t.data_ptr() if t is not None else 0)
        _check(lib().hqpdocp_update_dev(self.h, ctypes.c_int(grad_mode), ptr(x), ptr(f), ptr(b), ptr(d), ptr(g),
                                        ptr(fx), ptr(fu), ptr(cx), ptr(cu)), "hqpdocp_update_dev")

    def update_fbd_dev(self, x, f, b, d):
        ptr = lambda t: ctypes.c_void_p(t.data_ptr() if t is not None else 0)
        _check(lib().hqpdocp_update_fbd_dev(self.h, ptr(x), ptr(f), ptr(b), ptr(d)), "hqpdocp_update_fbd_dev")
