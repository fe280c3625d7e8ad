"""Golden vectors of the reference's block BFGS update (row f2): inputs and outputs of the
UNMODIFIED Hqp_HL_BFGS::update_b_Q (hqp/Hqp_HL_BFGS.C:149-213) compiled into oracle/_ref.
Run in the build container (needs /root/reference):  python tests/golden/make_hl_bfgs_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import refharness as R  # noqa: E402


def cases():
    rng = np.random.default_rng(20261017)
    out = []
    for t, n in enumerate([2, 3, 5, 8, 12, 16, 20, 30, 30, 30, 33, 50, 64]):
        M = rng.uniform(-1, 1, (n, n))
        Q = M.T @ M / n + 0.1 * np.eye(n)
        if t % 4 == 0:
            Q = np.eye(n)                      # first SQP iteration: scaled identity
        s, u = rng.uniform(-1, 1, n), rng.uniform(-1, 1, n)
        if t % 3 == 1:
            u = Q @ s + 0.05 * rng.uniform(-1, 1, n)   # consistent curvature: no damping
        if t % 5 == 2:
            u = -u                             # negative curvature: Powell's damping and a shift
        if t == 9:
            s = np.zeros(n)                    # s'u = s'Qs = 0: block left alone (:186-187)
        for ec in (1, 0):
            for gamma, alpha in ((0.1, 1.0), (-0.2, 0.35)):
                out.append((Q, s, u, alpha, gamma, 1e-8, ec))
    return out


def main():
    d = {}
    cs = cases()
    for i, (Q, s, u, alpha, gamma, eps, ec) in enumerate(cs):
        Qn = R.hl_bfgs_block(Q, s, u, alpha, gamma, eps, bool(ec))
        d[f"Q{i}"], d[f"s{i}"], d[f"u{i}"], d[f"out{i}"] = Q, s, u, Qn
        d[f"par{i}"] = np.array([alpha, gamma, eps, ec])
    d["n"] = np.array(len(cs))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "hl_bfgs_blocks.npz"), **d)
    print("wrote", len(cs), "cases")


if __name__ == "__main__":
    main()
