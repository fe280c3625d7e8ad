"""numpy implementation of the range-engine interface of hqp_b200/dist.py.

TEST INFRASTRUCTURE: lets the horizon-split exchange protocol (RangeSolver) run
on CPU tensors over gloo.  Same mathematics as the CUDA kernels (lq_factor.cuh
K1/K2b/K3, lq_solve.cuh, lq_range.cuh), written densely and sequentially."""
import numpy as np
import torch


class NumpyRangeEngine:
    def __init__(self, lp, rm):
        self.p, self.rm = lp, rm
        self.stage, self.lcol = lp.ineq_stage_local()

    # ---- helpers -----------------------------------------------------------
    def _H(self, z, w):
        p = self.p
        H = p.Q.copy()
        for r in range(p.m):
            k = self.stage[r]
            e0, e1 = p.ineq_ptr[r], p.ineq_ptr[r + 1]
            cols, vals = self.lcol[e0:e1], p.ineq_val[e0:e1]
            H[k][np.ix_(cols, cols)] += (z[r] / w[r]) * np.outer(vals, vals)
        return H

    @staticmethod
    def _stage(Hk, fx, fu, V, nx):
        F = np.hstack([fx, fu])
        G = Hk + F.T @ V @ F
        Guu, Gux = G[nx:, nx:], G[nx:, :nx]
        Rux = np.linalg.solve(Guu, Gux)
        Vn = G[:nx, :nx] - Gux.T @ Rux
        return 0.5 * (Vn + Vn.T), Rux, fx - fu @ Rux, Guu

    # ---- factor --------------------------------------------------------------
    def factor_begin(self, z, w):
        p, nx = self.p, self.p.nx
        self.z, self.w = z.numpy().copy(), w.numpy().copy()
        self.H = self._H(self.z, self.w)
        J, A, C = np.zeros((nx, nx)), np.eye(nx), np.zeros((nx, nx))
        for k in range(p.K - 1, -1, -1):
            J, Rux, Phi, Guu = self._stage(self.H[k], p.fx[k], p.fu[k], J, nx)
            W = A @ p.fu[k]
            C = C + W @ np.linalg.solve(Guu, W.T)
            A = A @ Phi
        self.Vterm = self.H[p.K][:nx, :nx].copy()
        return torch.from_numpy(np.stack([A, 0.5 * (C + C.T), J, self.Vterm]).ravel().copy())

    def factor_finish(self, gathered, rank, world):
        p, nx = self.p, self.p.nx
        g = gathered.numpy().reshape(world, 4, nx, nx)
        S = g[world - 1][3].copy()
        for r in range(world - 1, rank, -1):
            A, C, J = g[r][0], g[r][1], g[r][2]
            S = J + A.T @ np.linalg.solve(np.eye(nx) + S @ C, S @ A)
            S = 0.5 * (S + S.T)
        V = self.Vterm + (0.0 if rank == world - 1 else S)
        self.V = [None] * (p.K + 1)
        self.V[p.K] = V
        self.Rux, self.Phi, self.Guu = [None] * p.K, [None] * p.K, [None] * p.K
        Psi = np.eye(nx)
        for k in range(p.K - 1, -1, -1):
            V, self.Rux[k], self.Phi[k], self.Guu[k] = self._stage(self.H[k], p.fx[k], p.fu[k], V, nx)
            self.V[k] = V
            Psi = Psi @ self.Phi[k]
        self.Psi = Psi
        return torch.from_numpy(Psi.ravel().copy())

    # ---- solve ---------------------------------------------------------------
    def step_begin(self, r1, r2, r3, r4):
        p, nx, nm = self.p, self.p.nx, self.p.nm
        self.r = [t.numpy().copy() for t in (r1, r2, r3, r4)]
        r1, r2, r3, r4 = self.r
        g = -r1.copy()
        for r in range(p.m):
            k = self.stage[r]
            e0, e1 = p.ineq_ptr[r], p.ineq_ptr[r + 1]
            g[k * nm + self.lcol[e0:e1]] += p.ineq_val[e0:e1] * (self.z[r] * r3[r] + r4[r]) / self.w[r]
        self.g = g
        self.f = [r2[k * nx:(k + 1) * nx] for k in range(p.K)]
        self.wv = [g[k * nm:k * nm + nx] - self.Rux[k].T @ g[k * nm + nx:(k + 1) * nm] for k in range(p.K)]
        self.q = [self.V[k + 1] @ self.f[k] for k in range(p.K)]
        t = np.zeros(nx)
        for k in range(p.K - 1, -1, -1):
            t = self.wv[k] + self.Phi[k].T @ (t + self.q[k])
        self.gK = g[p.K * nm:p.K * nm + nx].copy()
        return torch.from_numpy(np.concatenate([t, self.gK]))

    def step_mid(self, gv, gpsi, rank, world):
        p, nx, nm = self.p, self.p.nx, self.p.nm
        gv = gv.numpy().reshape(world, 2, nx)
        gp = gpsi.numpy().reshape(world, nx, nx)
        t = gv[world - 1][1].copy()
        for r in range(world - 1, rank, -1):
            t = gv[r][0] + gp[r].T @ t
        self.v = [None] * (p.K + 1)
        self.v[p.K] = self.gK + (0.0 if rank == world - 1 else t)
        self.Ru, self.c = [None] * p.K, [None] * p.K
        for k in range(p.K - 1, -1, -1):
            tt = self.v[k + 1] + self.q[k]
            self.v[k] = self.wv[k] + self.Phi[k].T @ tt
            gu = self.g[k * nm + nx:(k + 1) * nm]
            self.Ru[k] = np.linalg.solve(self.Guu[k], gu + p.fu[k].T @ tt)
            self.c[k] = self.f[k] - p.fu[k] @ self.Ru[k]
        x = np.zeros(nx)
        for k in range(p.K):
            x = self.Phi[k] @ x + self.c[k]
        if rank > 0:
            xs = np.zeros(nx)
        elif p.fixed_x0:
            xs = -self.r[1][p.K * nx:p.K * nx + nx]
        else:
            xs = -np.linalg.solve(self.V[0], self.v[0])
        return torch.from_numpy(np.concatenate([x, xs]))

    def step_finish(self, gx, gpsi, rank, world):
        p, nx, nu, nm = self.p, self.p.nx, self.p.nu, self.p.nm
        gx = gx.numpy().reshape(world, 2, nx)
        gp = gpsi.numpy().reshape(world, nx, nx)
        t = gx[0][1].copy()
        for r in range(rank):
            t = gx[r][0] + gp[r] @ t
        r1, r2, r3, r4 = self.r
        dx, dy = np.zeros(p.N), np.zeros(p.me)
        x = t
        x0 = x.copy()
        for k in range(p.K):
            u = -(self.Rux[k] @ x + self.Ru[k])
            dx[k * nm:k * nm + nx] = -x
            dx[k * nm + nx:(k + 1) * nm] = -u
            x = self.Phi[k] @ x + self.c[k]
            dy[k * nx:(k + 1) * nx] = self.V[k + 1] @ x + self.v[k + 1]
        dx[p.K * nm:] = -x
        if p.fixed_x0:
            dy[p.K * nx:] = -(self.v[0] + self.V[0] @ x0)
        dz, dw = np.zeros(p.m), np.zeros(p.m)
        for r in range(p.m):
            k = self.stage[r]
            e0, e1 = p.ineq_ptr[r], p.ineq_ptr[r + 1]
            dw[r] = p.ineq_val[e0:e1] @ dx[k * nm + self.lcol[e0:e1]] - r3[r]
            dz[r] = (r4[r] - self.z[r] * dw[r]) / self.w[r]
        return [torch.from_numpy(a) for a in (dx, dy, dz, dw)]
