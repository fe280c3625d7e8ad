"""Stage-slab description of an LQ-DOCP QP, as the C ABI consumes it.

The reference never passes the stage structure explicitly: Hqp_IpLQDOCP::init
re-derives it from the sparsity of ``A`` (hqp/Hqp_IpLQDOCP.C:201-407).  The
``Hqp_IpCuda`` host module does the same and hands the result to the C ABI in
the layout below (include/hqp_ipcuda.h).  This Python class is the same layout
for the ctypes path used by tests/ and bench.py.

Variable order (hqp/Hqp_Docp.C:465-536): ``[x0,u0,x1,u1,...,xK]``,
``N = K*(nx+nu)+nx``.
Equality rows: ``K*nx`` dynamics rows ``[fx fu -I]`` (``A x + b = 0``), then
``nx`` rows ``+1`` fixing x0 when ``fixed_x0`` (Hqp_IpLQDOCP.C:344-351), then
``n_eq`` general stage-local equality rows.
Inequality rows ``C x + d >= 0`` in any order; each row is local to one stage
(Hqp_IpLQDOCP.C:323-341).
"""
from __future__ import annotations

import ctypes
import dataclasses
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SYNTH = None


def _synth_lib():
    global _SYNTH
    if _SYNTH is None:
        path = os.path.join(_HERE, "lib", "libhqpsynth.so")
        if not os.path.exists(path):
            raise RuntimeError(
                f"{path} missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
        _SYNTH = ctypes.CDLL(path)
    return _SYNTH


def _dp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


@dataclasses.dataclass
class LQProblem:
    nx: int
    nu: int
    K: int
    Q: np.ndarray          # (K+1, nm, nm) full symmetric, block K uses [:nx,:nx]
    c: np.ndarray          # (N,)
    fx: np.ndarray         # (K, nx, nx)
    fu: np.ndarray         # (K, nx, nu)
    b: np.ndarray          # (me,)
    fixed_x0: bool = True
    # inequality rows, CSR over GLOBAL columns
    ineq_ptr: np.ndarray = None
    ineq_col: np.ndarray = None
    ineq_val: np.ndarray = None
    d: np.ndarray = None
    # extra equality rows (after dynamics and x0 rows), CSR over GLOBAL columns
    eq_ptr: np.ndarray = None
    eq_col: np.ndarray = None
    eq_val: np.ndarray = None

    def __post_init__(self):
        if self.ineq_ptr is None:
            self.ineq_ptr = np.zeros(1, np.int32)
            self.ineq_col = np.zeros(0, np.int32)
            self.ineq_val = np.zeros(0)
            self.d = np.zeros(0)
        if self.eq_ptr is None:
            self.eq_ptr = np.zeros(1, np.int32)
            self.eq_col = np.zeros(0, np.int32)
            self.eq_val = np.zeros(0)

    # ---- sizes ---------------------------------------------------------
    @property
    def nm(self):
        return self.nx + self.nu

    @property
    def N(self):
        return self.K * self.nm + self.nx

    @property
    def n_eq(self):
        return len(self.eq_ptr) - 1

    @property
    def me(self):
        return self.K * self.nx + (self.nx if self.fixed_x0 else 0) + self.n_eq

    @property
    def m(self):
        return len(self.ineq_ptr) - 1

    # ---- stage maps for the C ABI -----------------------------------------
    def _stage_local(self, ptr, col):
        """(stage per row, local col per nonzero); rows must be stage-local."""
        nrows = len(ptr) - 1
        stage_nz = np.minimum(col // self.nm, self.K)
        local = col - stage_nz * self.nm
        stage = np.zeros(nrows, np.int32)
        if nrows:
            first = ptr[:-1]
            nonempty = ptr[1:] > first
            stage[nonempty] = stage_nz[first[nonempty]]
            rows = np.repeat(np.arange(nrows), np.diff(ptr))
            if not np.all(stage_nz == stage[rows]):
                raise ValueError("constraint row couples different stages")
        return stage.astype(np.int32), local.astype(np.int32)

    def ineq_stage_local(self):
        return self._stage_local(self.ineq_ptr, self.ineq_col)

    def eq_stage_local(self):
        return self._stage_local(self.eq_ptr, self.eq_col)

    # ---- global CSR (for the reference harness / dense checks) ------------
    def csr_Q_upper(self):
        nm, nx, K = self.nm, self.nx, self.K
        iu, ju = np.triu_indices(nm)
        base = (np.arange(K) * nm)[:, None]
        rows = (base + iu[None, :]).ravel()
        cols = (base + ju[None, :]).ravel()
        vals = self.Q[:K][:, iu, ju].ravel()
        it, jt = np.triu_indices(nx)
        rows = np.concatenate([rows, K * nm + it])
        cols = np.concatenate([cols, K * nm + jt])
        vals = np.concatenate([vals, self.Q[K][it, jt]])
        return _coo_to_csr(self.N, rows, cols, vals)

    def csr_A(self):
        nm, nx, nu, K = self.nm, self.nx, self.nu, self.K
        r = (np.arange(K)[:, None, None] * nx + np.arange(nx)[None, :, None])
        cx = np.arange(K)[:, None, None] * nm + np.arange(nx)[None, None, :]
        cu = np.arange(K)[:, None, None] * nm + nx + np.arange(nu)[None, None, :]
        rows = [np.broadcast_to(r, (K, nx, nx)).ravel(),
                np.broadcast_to(r, (K, nx, nu)).ravel(),
                (np.arange(K)[:, None] * nx + np.arange(nx)[None, :]).ravel()]
        cols = [np.broadcast_to(cx, (K, nx, nx)).ravel(),
                np.broadcast_to(cu, (K, nx, nu)).ravel(),
                ((np.arange(K)[:, None] + 1) * nm + np.arange(nx)[None, :]).ravel()]
        vals = [self.fx.ravel(), self.fu.ravel(), -np.ones(K * nx)]
        row0 = K * nx
        if self.fixed_x0:
            rows.append(row0 + np.arange(nx))
            cols.append(np.arange(nx))
            vals.append(np.ones(nx))
            row0 += nx
        if self.n_eq:
            rows.append(row0 + np.repeat(np.arange(self.n_eq), np.diff(self.eq_ptr)))
            cols.append(self.eq_col)
            vals.append(self.eq_val)
        return _coo_to_csr(self.me, np.concatenate(rows), np.concatenate(cols),
                           np.concatenate(vals))

    def value_map(self):
        """(vals, dst, dst2) for hqpcu_set_value_map /
        hqpcu_update_values, built from the CSR patterns of Q (upper triangle) and
        of A exactly as a host that only holds the sparse matrices would: value i
        of [Q entries in CSR order | entries of the K*nx dynamics rows of A, the
        -1 of x_{k+1} left out] goes to slab position dst[i] (and dst2[i], the
        mirrored entry of the symmetric Q block, or -1)."""
        nm, nx, nu, K = self.nm, self.nx, self.nu, self.K
        szQ, szX = (K + 1) * nm * nm, K * nx * nx
        qp, qj, qv = self.csr_Q_upper()
        qi = np.repeat(np.arange(self.N), np.diff(qp))
        k = np.minimum(qi // nm, K)
        li, lj = qi - k * nm, qj.astype(np.int64) - k * nm
        d1 = k * nm * nm + li * nm + lj
        d2 = np.where(li != lj, k * nm * nm + lj * nm + li, -1)
        ap, aj, av = self.csr_A()
        ai = np.repeat(np.arange(self.me), np.diff(ap))
        dyn = ai < K * nx
        ai, aj, av = ai[dyn], aj[dyn].astype(np.int64), av[dyn]
        ka = ai // nx
        lc = aj - ka * nm
        keep = lc < nm                      # drops the -1 entry in the columns of stage k+1
        ai, ka, lc, av = ai[keep], ka[keep], lc[keep], av[keep]
        r = ai - ka * nx
        da = np.where(lc < nx, szQ + (ka * nx + r) * nx + lc,
                      szQ + szX + (ka * nx + r) * nu + (lc - nx))
        vals = np.concatenate([qv, av]).astype(np.float64)
        dst = np.concatenate([d1, da]).astype(np.int64)
        dst2 = np.concatenate([d2, -np.ones(len(da), np.int64)]).astype(np.int64)
        return np.ascontiguousarray(vals), np.ascontiguousarray(dst), np.ascontiguousarray(dst2)

    def csr_C(self):
        return (self.ineq_ptr.astype(np.int32), self.ineq_col.astype(np.int32),
                self.ineq_val.astype(np.float64))

    def dense_kkt_blocks(self):
        """Dense Q (symmetric), A, C.  Only for small problems (tests)."""
        Qd = np.zeros((self.N, self.N))
        nm = self.nm
        for k in range(self.K + 1):
            dk = nm if k < self.K else self.nx
            o = k * nm
            Qd[o:o + dk, o:o + dk] = self.Q[k][:dk, :dk]
        p, j, v = self.csr_A()
        A = np.zeros((self.me, self.N))
        A[np.repeat(np.arange(self.me), np.diff(p)), j] = v
        C = np.zeros((self.m, self.N))
        if self.m:
            C[np.repeat(np.arange(self.m), np.diff(self.ineq_ptr)), self.ineq_col] = self.ineq_val
        return Qd, A, C


def _coo_to_csr(nrows, rows, cols, vals):
    order = np.lexsort((cols, rows))
    rows, cols, vals = rows[order], cols[order], vals[order]
    ptr = np.zeros(nrows + 1, np.int64)
    np.add.at(ptr, rows + 1, 1)
    ptr = np.cumsum(ptr)
    return ptr.astype(np.int32), cols.astype(np.int32), vals.astype(np.float64)


def box_bounds_on_u(nx, nu, K):
    """-1 <= u <= 1 as rows 2*(k*nu+j) (+u_j+1>=0) and +1 (-u_j+1>=0)
    (SURVEY.md App. B.7)."""
    m = 2 * nu * K
    ptr = np.arange(m + 1, dtype=np.int32)
    ucol = (np.arange(K)[:, None] * (nx + nu) + nx + np.arange(nu)[None, :]).ravel()
    col = np.repeat(ucol, 2).astype(np.int32)
    val = np.tile(np.array([1.0, -1.0]), nu * K)
    return ptr, col, val, np.ones(m)


def synth_lqdocp(nx, nu, K, seed=1234, bounds=True) -> LQProblem:
    """Seeded synthetic LQ-DOCP of SURVEY.md App. B.7 / section 8(d)."""
    lib = _synth_lib()
    nm = nx + nu
    Q = np.zeros((K + 1, nm, nm))
    c = np.zeros(K * nm + nx)
    fx = np.zeros((K, nx, nx))
    fu = np.zeros((K, nx, nu))
    b = np.zeros(K * nx + nx)
    lib.hqp_synth_lqdocp(ctypes.c_int(nx), ctypes.c_int(nu), ctypes.c_int(K),
                         ctypes.c_ulonglong(seed), _dp(Q), _dp(c), _dp(fx),
                         _dp(fu), _dp(b))
    prob = LQProblem(nx, nu, K, Q, c, fx, fu, b, fixed_x0=True)
    if bounds:
        prob.ineq_ptr, prob.ineq_col, prob.ineq_val, prob.d = box_bounds_on_u(nx, nu, K)
    return prob


def synth_rhs(prob: LQProblem, seed=4321):
    """z, w, r1..r4 of SURVEY.md App. B.7 (direct-plugin timing RHS)."""
    lib = _synth_lib()
    N, me, m = prob.N, prob.me, prob.m
    z, w, r3, r4 = (np.zeros(m) for _ in range(4))
    r1, r2 = np.zeros(N), np.zeros(me)
    lib.hqp_synth_rhs(ctypes.c_int(N), ctypes.c_int(me), ctypes.c_int(m),
                      ctypes.c_ulonglong(seed), _dp(z), _dp(w), _dp(r1),
                      _dp(r2), _dp(r3), _dp(r4))
    return z, w, r1, r2, r3, r4


def add_random_stage_ineq(prob: LQProblem, rows_per_stage=2, nnz_per_row=3, seed=7,
                          include_terminal=True):
    """Append general stage-local inequality rows  c'[x_k;u_k] + d >= 0  (mixed
    state/control constraints like Prg_DID's x2 + 0.5 dt x1 <= 0.01,
    hqp_docp/Prg_DID.C) to the problem's C; used to exercise the dense
    C'(z/w)C path of Hqp_IpLQDOCP::factor (hqp/Hqp_IpLQDOCP.C:68-103)."""
    rng = np.random.default_rng(seed)
    ptr = list(prob.ineq_ptr)
    col = list(prob.ineq_col)
    val = list(prob.ineq_val)
    d = list(prob.d)
    nm = prob.nm
    for k in range(prob.K + (1 if include_terminal else 0)):
        dk = nm if k < prob.K else prob.nx
        for _ in range(rows_per_stage):
            nz = min(nnz_per_row, dk)
            cols = np.sort(rng.choice(dk, size=nz, replace=False))
            col.extend((k * nm + cols).tolist())
            val.extend(rng.uniform(-1, 1, nz).tolist())
            d.append(float(rng.uniform(0.5, 2.0)))
            ptr.append(len(col))
    prob.ineq_ptr = np.asarray(ptr, np.int32)
    prob.ineq_col = np.asarray(col, np.int32)
    prob.ineq_val = np.asarray(val, np.float64)
    prob.d = np.asarray(d, np.float64)
    return prob


def rhs_for(prob: LQProblem, seed=4321):
    """like synth_rhs but numpy-seeded (works for any row count)"""
    rng = np.random.default_rng(seed)
    m = prob.m
    return (rng.uniform(0.5, 1.5, m), rng.uniform(0.5, 1.5, m), rng.uniform(-1, 1, prob.N),
            rng.uniform(-1, 1, prob.me), rng.uniform(-1, 1, m), rng.uniform(-1, 1, m))


def add_stage_equalities(prob: LQProblem, stages, rows_per_stage=1, nnz_per_row=None, seed=21):
    """Append general stage-local equality rows  e'[x_k;u_k] + b = 0  (e.g. a
    terminal state constraint like Prg_DID's x_K = (-1, 0), hqp_docp/Prg_DID.C)
    after the dynamics / x0 rows of A.  Rows of one stage are linearly
    independent with probability one."""
    rng = np.random.default_rng(seed)
    ptr, col, val = list(prob.eq_ptr), list(prob.eq_col), list(prob.eq_val)
    b = list(prob.b)
    nm = prob.nm
    for k in stages:
        dk = nm if k < prob.K else prob.nx
        for _ in range(rows_per_stage):
            nz = dk if nnz_per_row is None else min(nnz_per_row, dk)
            cols = np.sort(rng.choice(dk, size=nz, replace=False))
            col.extend((k * nm + cols).tolist())
            val.extend(rng.uniform(-1, 1, nz).tolist())
            b.append(float(rng.uniform(-0.1, 0.1)))
            ptr.append(len(col))
    prob.eq_ptr = np.asarray(ptr, np.int32)
    prob.eq_col = np.asarray(col, np.int32)
    prob.eq_val = np.asarray(val, np.float64)
    prob.b = np.asarray(b, np.float64)
    return prob
