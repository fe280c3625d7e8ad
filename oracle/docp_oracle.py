"""TEST INFRASTRUCTURE, not product code: plain-Python restatement of the reference's DOCP
update.  Only tests/, smoke() and bench.py's cpu_baseline leg
may import it.  Scalar Python loops: small cases only.

  update_fbd    Hqp_Docp::update_fbd     hqp/Hqp_Docp.C:831-891   f, b, d for the iterate x
  update        Hqp_Docp::update         hqp/Hqp_Docp.C:944-1075  + g (qp->c), fx, fu, cx, cu
  grds_fd       Hqp_Docp::update_grds    hqp/Hqp_Docp.C:1097-1180 forward differences, dv = 1e-4|v|+1e-6
  bounds        Hqp_Docp::update_bounds  hqp/Hqp_Docp.C:893-940
  assemble_AC   Hqp_Docp::setup_qp + the derivative part of ::update (:597-741, 1045-1064):
                dense A, C in the reference's row order
  vals_did      Prg_DID::update_vals     hqp_docp/Prg_DID.C:78-98 ; grds_did: Prg_DID::setup_struct /
                ::update_stage (:101-165): constant Jacobians, f0u = 2 u dt
  vals_synthnl  the synthetic model (hqp_b200/csrc/docp_models.cuh) in the same operation order;
                grds_synthnl_exact: its derivatives in closed form (numpy)
A DocpProblem may be a stage range of a longer horizon (DocpProblem.shard): stage indices are
local, the range's last state is a halo that is only read.
Pinned against the compiled reference (oracle/_ref: ref_docp_update drives the UNMODIFIED
Hqp_Docp for Prg_DID and for oracle/prg_synthnl.cpp) in tests/test_docp_update.py, and against
the golden vectors tests/golden/docp_update_*.npz made from it by tests/golden/make_docp_update_golden.py.
"""
from __future__ import annotations

import numpy as np

MODEL_DID, MODEL_SYNTHNL = 0, 1


def vals_did(p, k, x, u):
    dt = float(p.par[0])
    f, c, f0 = [0.0, 0.0], [0.0] * (p.nc if k < p.K else 0), 0.0
    if k < p.K:
        f[0] = x[0] + u[0] * dt
        f[1] = x[0] * dt + x[1] + u[0] * 0.5 * dt * dt
        f0 = u[0] * u[0] * dt
        if p.nc:
            c[0] = x[1] + 0.5 * dt * x[0]
    return f, f0, c


def vals_synthnl(p, k, x, u):
    nx, nu = p.nx, p.nu
    par = [float(v) for v in p.par]
    eps = par[0]
    A = par[1:1 + nx * nx]
    B = par[1 + nx * nx:1 + nx * nx + nx * nu]
    qw = par[1 + nx * nx + nx * nu:1 + nx * nx + nx * nu + nx]
    rw = par[1 + nx * nx + nx * nu + nx:]
    r = [float(v) for v in p.spar[k]]
    s = 0.0
    for i in range(nx):
        e = x[i] - r[i]
        s = s + qw[i] * e * e
    f = [0.0] * nx
    if k < p.K:
        for i in range(nx):
            acc = 0.0
            for j in range(nx):
                acc = acc + A[i * nx + j] * x[j]
            for j in range(nu):
                acc = acc + B[i * nu + j] * u[j]
            f[i] = acc + eps * (x[i] / (1.0 + x[i] * x[i]))
        for j in range(nu):
            s = s + rw[j] * u[j] * u[j]
        s = 0.5 * s
        for j in range(min(nx, nu)):
            s = s + eps * x[j] * u[j]
        f0 = s
        c = [0.0] * p.nc
        if p.nc > 0:
            q = 0.0
            for i in range(nx):
                q = q + x[i] * x[i]
            c[0] = q / float(nx) + eps * u[0] * x[0]
            for i in range(1, p.nc):
                c[i] = x[i] * u[i % nu]
    else:
        f0 = 0.5 * s
        c = [x[i] * x[i] for i in range(p.ncK)]
    return f, f0, c


def _vals(p, k, x, u):
    return (vals_did if p.model == MODEL_DID else vals_synthnl)(p, k, x, u)


def _stage(p, xv, k):
    nd = p.nx + p.nu
    x = [float(v) for v in xv[k * nd:k * nd + p.nx]]
    u = [float(v) for v in xv[k * nd + p.nx:(k + 1) * nd]] if k < p.K else []
    return x, u


def bounds(p, xv, cval, b, d):
    """update_bounds plus the constraint rows of update_fbd: the association tables -> b, d."""
    o = p.K * p.nx
    for t, src in ((p.xu_eq, xv), (p.cns_eq, cval)):
        for i, v in zip(t.idxs, t.vals):
            b[o] = src[i] - v
            o += 1
    o = 0
    for t, src, sign in ((p.xu_lb, xv, 1), (p.xu_ub, xv, -1), (p.cns_lb, cval, 1), (p.cns_ub, cval, -1)):
        for i, v in zip(t.idxs, t.vals):
            d[o] = src[i] - v if sign > 0 else v - src[i]
            o += 1


def update_fbd(p, xv):
    xv = np.asarray(xv, dtype=np.float64)
    b, d, cval = np.zeros(p.me), np.zeros(p.m), np.zeros(p.ncns)
    nd = p.nx + p.nu
    fsum = 0.0
    for k in range(p.K + 1 if p.owns_final else p.K):  # (a stage range: the halo stage is not ours)
        x, u = _stage(p, xv, k)
        f, f0, c = _vals(p, k, x, u)
        fsum += f0
        if k < p.K:
            for i in range(p.nx):
                b[k * p.nx + i] = f[i] - float(xv[(k + 1) * nd + i])
        for i, v in enumerate(c):
            cval[k * p.nc + i] = v
    bounds(p, xv, cval, b, d)
    return fsum, b, d, cval


def grds_fd(p, k, x, u):
    """Hqp_Docp::update_grds: returns fx, fu, f0x, f0u, cx, cu of stage k (lists of rows)."""
    nx, nu = len(x), len(u)
    f, f0, c = _vals(p, k, x, u)
    nf, nc = (p.nx if k < p.K else 0), len(c)
    fx = np.zeros((nf, nx)); fu = np.zeros((nf, nu)); cx = np.zeros((nc, nx)); cu = np.zeros((nc, nu))
    f0x = np.zeros(nx); f0u = np.zeros(nu)
    for vec, J, g0, Jc in ((x, fx, f0x, cx), (u, fu, f0u, cu)):
        for j in range(len(vec)):
            bak = vec[j]
            dv = 1e-4 * abs(bak) + 1e-6
            vec[j] = vec[j] + dv
            df, df0, dc = _vals(p, k, x, u)
            for i in range(nf):
                J[i, j] = (df[i] - f[i]) / dv
            g0[j] = (df0 - f0) / dv
            for i in range(nc):
                Jc[i, j] = (dc[i] - c[i]) / dv
            vec[j] = bak
    return fx, fu, f0x, f0u, cx, cu


def grds_did(p, k, x, u):
    dt = float(p.par[0])
    if k < p.K:
        fx = np.array([[1.0, 0.0], [dt, 1.0]]); fu = np.array([[dt], [0.5 * dt * dt]])
        cx = np.array([[0.5 * dt, 1.0]]) if p.nc else np.zeros((0, 2))
        cu = np.zeros((p.nc, 1))
        return fx, fu, np.zeros(2), np.array([2.0 * u[0] * dt]), cx, cu
    return np.zeros((0, 2)), np.zeros((0, 0)), np.zeros(2), np.zeros(0), np.zeros((0, 2)), np.zeros((0, 0))


def grds_synthnl_exact(p, k, x, u):
    nx, nu = p.nx, p.nu
    par = np.asarray(p.par, dtype=np.float64)
    eps = par[0]
    A = par[1:1 + nx * nx].reshape(nx, nx)
    B = par[1 + nx * nx:1 + nx * nx + nx * nu].reshape(nx, nu)
    qw = par[1 + nx * nx + nx * nu:1 + nx * nx + nx * nu + nx]
    rw = par[1 + nx * nx + nx * nu + nx:]
    x = np.asarray(x); u = np.asarray(u)
    f0x = qw * (x - p.spar[k])
    if k == p.K:
        cx = np.zeros((p.ncK, nx))
        for i in range(p.ncK):
            cx[i, i] = 2 * x[i]
        return np.zeros((0, nx)), np.zeros((0, 0)), f0x, np.zeros(0), cx, np.zeros((p.ncK, 0))
    fx = A + eps * np.diag((1 - x * x) / (1 + x * x) ** 2)
    f0u = rw * u
    nm = min(nx, nu)
    f0x[:nm] += eps * u[:nm]
    f0u[:nm] += eps * x[:nm]
    cx = np.zeros((p.nc, nx)); cu = np.zeros((p.nc, nu))
    if p.nc:
        cx[0] = 2 * x / nx
        cx[0, 0] += eps * u[0]
        cu[0, 0] = eps * x[0]
        for i in range(1, p.nc):
            cx[i, i] += u[i % nu]
            cu[i, i % nu] += x[i]
    return fx, B.copy(), f0x, f0u, cx, cu


def update(p, xv, grads="fd"):
    """grads: "fd" (the reference's default update_grds), "did" (Prg_DID's own update_stage) or
    "exact" (closed-form derivatives of the synthetic model).  Returns the dict DocpCuda.update
    returns."""
    xv = np.asarray(xv, dtype=np.float64)
    f, b, d, _ = update_fbd(p, xv)
    nd = p.nx + p.nu
    g = np.zeros(p.N)
    fx = np.zeros((p.K, p.nx, p.nx)); fu = np.zeros((p.K, p.nx, p.nu))
    cx = np.zeros((p.ncns, p.nx)); cu = np.zeros((p.K * p.nc, p.nu))
    fn = {"fd": grds_fd, "did": grds_did, "exact": grds_synthnl_exact}[grads]
    for k in range(p.K + 1 if p.owns_final else p.K):
        x, u = _stage(p, xv, k)
        jfx, jfu, f0x, f0u, jcx, jcu = fn(p, k, x, u)
        g[k * nd:k * nd + p.nx] = f0x
        if k < p.K:
            g[k * nd + p.nx:(k + 1) * nd] = f0u
            fx[k], fu[k] = jfx, jfu
            cx[k * p.nc:(k + 1) * p.nc] = jcx
            cu[k * p.nc:(k + 1) * p.nc] = jcu
        else:
            cx[p.K * p.nc:] = jcx
    return dict(f=f, b=b, d=d, g=g, fx=fx, fu=fu, cx=cx, cu=cu)


def assemble_AC(p, fx, fu, cx, cu):
    """Dense A [me, N] and C [m, N] as Hqp_Docp::setup_qp lays the rows out and ::update fills them:
    A = [dynamics rows [fx fu -I]; unit rows of the fixed x/u; constraint equalities],
    C = [+e_i lower bounds; -e_i upper bounds; +[cx cu] rows; -[cx cu] rows]."""
    nd = p.nx + p.nu
    A = np.zeros((p.me, p.N)); C = np.zeros((p.m, p.N))
    for k in range(p.K):
        r = slice(k * p.nx, (k + 1) * p.nx)
        A[r, k * nd:k * nd + p.nx] = fx[k]
        A[r, k * nd + p.nx:(k + 1) * nd] = fu[k]
        A[r, (k + 1) * nd:(k + 1) * nd + p.nx] -= np.eye(p.nx)

    def crow(ci):
        k = min(ci // p.nc, p.K) if p.nc else p.K
        row = np.zeros(p.N)
        row[k * nd:k * nd + p.nx] = cx[ci]
        if k < p.K:
            row[k * nd + p.nx:(k + 1) * nd] = cu[ci]
        return row

    o = p.K * p.nx
    for i in p.xu_eq.idxs:
        A[o, i] = 1.0
        o += 1
    for ci in p.cns_eq.idxs:
        A[o] = crow(ci)
        o += 1
    o = 0
    for i in p.xu_lb.idxs:
        C[o, i] = 1.0
        o += 1
    for i in p.xu_ub.idxs:
        C[o, i] = -1.0
        o += 1
    for ci in p.cns_lb.idxs:
        C[o] = crow(ci)
        o += 1
    for ci in p.cns_ub.idxs:
        C[o] = -crow(ci)
        o += 1
    return A, C
