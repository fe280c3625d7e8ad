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
