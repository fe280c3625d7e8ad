import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, GOLDEN)

from hqp_b200.problem import (synth_lqdocp, add_random_stage_ineq, rhs_for,  # noqa: E402
                              add_stage_equalities)


def make_problem(nx, nu, K, bounds, gen, fixed):
    """must stay identical to tests/golden/make_golden.py:make_problem"""
    p = synth_lqdocp(nx, nu, K, bounds=bool(bounds))
    if gen:
        add_random_stage_ineq(p, rows_per_stage=int(gen), nnz_per_row=3, seed=11)
    if not fixed:
        p.fixed_x0 = False
        p.b = p.b[:K * nx].copy()
    return p


def make_eq_problem(nx, nu, K, bounds, gen, fixed, stages, rows):
    """must stay identical to tests/golden/make_golden.py:make_eq_problem"""
    p = make_problem(nx, nu, K, bounds, gen, fixed)
    return add_stage_equalities(p, stages, rows, seed=33)


def eq_cases():
    import make_golden
    return make_golden.EQ_CASES


def relerr(a, b):
    a, b = np.asarray(a), np.asarray(b)
    if a.size == 0:
        return 0.0
    return float(np.max(np.abs(a - b)) / max(1e-300, np.max(np.abs(b))))


def golden_step_cases():
    return sorted(f[:-4] for f in os.listdir(GOLDEN) if f.startswith("step_") and f.endswith(".npz"))


def load_golden(name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    cfg = [int(v) for v in g["cfg"]]
    return g, cfg


def dense_kkt_solve(p, z, w, r1, r2, r3, r4):
    """Independent dense solve of the 4x4 block system (hqp/Hqp_IpsMehrotra.C:27-31)."""
    Q, A, C = p.dense_kkt_blocks()
    N, me, m = p.N, p.me, p.m
    Kmat = np.zeros((N + me + 2 * m, N + me + 2 * m))
    Kmat[:N, :N] = -Q
    Kmat[:N, N:N + me] = A.T
    Kmat[:N, N + me:N + me + m] = C.T
    Kmat[N:N + me, :N] = A
    Kmat[N + me:N + me + m, :N] = C
    Kmat[N + me:N + me + m, N + me + m:] = -np.eye(m)
    Kmat[N + me + m:, N + me:N + me + m] = np.diag(w)
    Kmat[N + me + m:, N + me + m:] = np.diag(z)
    sol = np.linalg.solve(Kmat, np.concatenate([r1, r2, r3, r4]))
    return sol[:N], sol[N:N + me], sol[N + me:N + me + m], sol[N + me + m:]


class VarDimQP:
    """A DOCP-structured QP with stage-dependent nx_k / nu_k in the layout of
    hqp/Hqp_Docp.C, as CSR for oracle/refharness.RefQP: what
    Hqp_IpLQDOCP::Get_Dim handles with _nk[k], _mk[k] (hqp/Hqp_IpLQDOCP.C:201-287)."""

    def __init__(self, nxs, nus, seed=0):
        rng = np.random.default_rng(seed)
        K = len(nus)
        assert len(nxs) == K + 1 and nxs[0] == nxs[1] and min(nus) >= 1
        xoff = np.concatenate([[0], np.cumsum([nxs[k] + nus[k] for k in range(K)])]).astype(int)
        self.N = int(xoff[K] + nxs[K])
        ndyn = int(sum(nxs[1:]))
        self.me = ndyn + nxs[0]
        qr, qc, qv = [], [], []
        ar, ac, av = [], [], []
        cr, cc, cv, d = [], [], [], []
        self.c = rng.uniform(-1, 1, self.N)
        b = []
        row = 0
        for k in range(K + 1):
            dk = nxs[k] + (nus[k] if k < K else 0)
            M = rng.uniform(-1, 1, (dk, dk))
            H = M.T @ M / dk + 0.1 * np.eye(dk)
            for i in range(dk):
                for j in range(i, dk):
                    qr.append(xoff[k] + i); qc.append(xoff[k] + j); qv.append(H[i, j])
            if k < K:
                fx = np.zeros((nxs[k + 1], nxs[k]))
                mdim = min(nxs[k + 1], nxs[k])
                fx[:mdim, :mdim] = np.eye(mdim)
                fx += 0.1 * rng.uniform(-1, 1, fx.shape) / np.sqrt(max(nxs[k], 1))
                fu = rng.uniform(-1, 1, (nxs[k + 1], nus[k]))
                for i in range(nxs[k + 1]):
                    for j in range(nxs[k]):
                        ar.append(row); ac.append(xoff[k] + j); av.append(fx[i, j])
                    for j in range(nus[k]):
                        ar.append(row); ac.append(xoff[k] + nxs[k] + j); av.append(fu[i, j])
                    ar.append(row); ac.append(xoff[k + 1] + i); av.append(-1.0)
                    b.append(0.01 * rng.uniform(-1, 1))
                    row += 1
                for j in range(nus[k]):          # -1 <= u <= 1
                    col = xoff[k] + nxs[k] + j
                    cr.append(len(d)); cc.append(col); cv.append(1.0); d.append(1.0)
                    cr.append(len(d)); cc.append(col); cv.append(-1.0); d.append(1.0)
        for i in range(nxs[0]):                  # fixed x0, rows after the dynamics rows
            ar.append(row); ac.append(i); av.append(1.0); b.append(-rng.uniform(-1, 1)); row += 1
        self.b = np.asarray(b)
        self.d = np.asarray(d)
        self.m = len(d)
        from hqp_b200.problem import _coo_to_csr
        self._Q = _coo_to_csr(self.N, np.asarray(qr), np.asarray(qc), np.asarray(qv))
        self._A = _coo_to_csr(self.me, np.asarray(ar), np.asarray(ac), np.asarray(av))
        self._C = _coo_to_csr(self.m, np.asarray(cr), np.asarray(cc), np.asarray(cv))

    def csr_Q_upper(self):
        return self._Q

    def csr_A(self):
        return self._A

    def csr_C(self):
        return self._C
