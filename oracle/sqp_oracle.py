"""TEST INFRASTRUCTURE, not product code: numpy restatement of the reference's SQP-level
vector operations.  Only tests/, smoke() and bench.py's
cpu_baseline leg may import it.

  grd_L   Hqp_SqpSolver::grd_L      hqp/Hqp_SqpSolver.C:430-445   c - A'y - C'z
  norm    Hqp_SqpSolver::norm_inf   :155-174                      max(||b||inf, -min d)
  quad    x'Qx, s'Qs                :225-226, 258-259, 299-301    sp_mv_symmlt + in_prod
  phi     Hqp_SqpPowell::phi        hqp/Hqp_SqpPowell.C:189-210   f + sum re|b| - sum r min(0,d)
  phi1    Hqp_SqpPowell::phi1       :213-244                      f + c's + sum re|As+b| - sum r min(0,Cs+d)
Pinned against the compiled reference (oracle/_ref, ref_sqp_eval) in tests/test_sqp_ops.py.
"""
from __future__ import annotations

import numpy as np


def _mv(csr, x, nrows):
    ptr, col, val = csr
    out = np.zeros(nrows)
    rows = np.repeat(np.arange(nrows), np.diff(ptr))
    np.add.at(out, rows, val * x[col])
    return out


def _vm(csr, y, ncols):
    ptr, col, val = csr
    rows = np.repeat(np.arange(len(ptr) - 1), np.diff(ptr))
    out = np.zeros(ncols)
    np.add.at(out, col, val * y[rows])
    return out


def grd_L(prob, c, y, z):
    out = _vm(prob.csr_A(), y, prob.N)
    if prob.m:
        out += _vm(prob.csr_C(), z, prob.N)
    return c - out


def quad(prob, x):
    """x'Qx with Q given by its upper triangle (symsp)."""
    up = prob.csr_Q_upper()
    ptr, col, val = up
    rows = np.repeat(np.arange(prob.N), np.diff(ptr))
    off = rows != col
    return float(np.sum(val * x[rows] * x[col]) + np.sum(val[off] * x[rows[off]] * x[col[off]]))


def merit(prob, f, c, s, b, d, re, r):
    """[phi, phi1, s'Qs, c's, norm_inf]"""
    As = _mv(prob.csr_A(), s, prob.me)
    phi = f + float(np.sum(re * np.abs(b)))
    phi1 = f + float(c @ s) + float(np.sum(re * np.abs(As + b)))
    nrm = float(np.max(np.abs(b))) if prob.me else 0.0
    if prob.m:
        Cs = _mv(prob.csr_C(), s, prob.m)
        phi -= float(np.sum(r * np.minimum(0.0, d)))
        phi1 -= float(np.sum(r * np.minimum(0.0, Cs + d)))
        nrm = max(nrm, float(-np.min(d)))
    return np.array([phi, phi1, quad(prob, s), float(c @ s), nrm])
