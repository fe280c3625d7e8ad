"""TEST INFRASTRUCTURE, not product code: numpy restatement of the reference's
block-diagonal BFGS update.  Only tests/, smoke() and
bench.py's cpu_baseline leg may import it.

Follows Hqp_HL_BFGS::update_b_Q, hqp/Hqp_HL_BFGS.C:149-213, statement by statement;
the eigenvalues of the reference's symmeig (meschach/symmeig.c:174) come from
numpy.linalg.eigvalsh here.  Pinned against the compiled reference
(oracle/_ref, ref_hl_bfgs_block) in tests/test_oracle.py and against
tests/golden/hl_bfgs_blocks.npz (generated from the reference by
tests/golden/make_hl_bfgs_golden.py).
"""
from __future__ import annotations

import numpy as np


def update_block(Q, s, u, alpha, gamma=0.1, eps=1e-8, eigen_control=True):
    """One block, in place on a copy; Q: (n, n) with both triangles.  Returns
    (Q_new, shifted, skipped)."""
    Q = np.array(Q, dtype=np.float64, copy=True)
    s = np.asarray(s, dtype=np.float64)
    u = np.asarray(u, dtype=np.float64)
    sv = float(s @ u)                      # :156
    sQ = s @ Q                             # :158  vm_mlt
    Qs = Q @ s                             # :159  mv_mlt
    sQs = float(sQ @ s)                    # :160
    if gamma >= 0.0:                       # :162-172
        g = gamma
    else:
        g = -gamma
        g = g + (1.0 - g) * (1.0 - alpha)
    if sv < g * sQs:                       # :175-181 Powell's modification
        theta = (1.0 - g) * sQs / (sQs - sv)
        v = theta * u + (1.0 - theta) * Qs
        sv = float(s @ v)
    else:
        v = u.copy()
    if not (sv != 0.0) or not (sQs != 0.0):  # :186-187
        return Q, False, True
    n = Q.shape[0]
    for i in range(n):                     # :194-202 upper triangle, mirrored with eigen control
        Q[i, i:] -= Qs[i] * sQ[i:] / sQs
        Q[i, i:] += v[i] * v[i:] / sv
        if eigen_control:
            Q[i:, i] = Q[i, i:]
    shifted = False
    if eigen_control:                      # :204-212
        theta = eps * eps
        if sQs < theta and sQs >= 0.0:
            theta = sQs
        lmin = float(np.min(np.linalg.eigvalsh(Q))) - theta
        if lmin < 0.0:
            Q[np.diag_indices(n)] -= lmin
            shifted = True
    return Q, shifted, False


def update(bsize, Qpacked, s, u, alpha, gamma=0.1, eps=1e-8, eigen_control=True):
    """All blocks (Hqp_HL_BFGS::update, :216-243): Qpacked = the blocks one after the
    other, row-major, both triangles; s, u in block order."""
    out = np.array(Qpacked, dtype=np.float64, copy=True)
    qo = vo = 0
    nshift = nskip = 0
    for n in bsize:
        blk, sh, sk = update_block(out[qo:qo + n * n].reshape(n, n), s[vo:vo + n], u[vo:vo + n], alpha,
                                   gamma, eps, eigen_control)
        out[qo:qo + n * n] = blk.ravel()
        nshift += sh
        nskip += sk
        qo += n * n
        vo += n
    return out, nshift, nskip
