"""Generate the committed golden fixtures by RUNNING THE UNMODIFIED REFERENCE
(oracle/_ref, built from /root/reference by `make -C oracle ref`).

    python tests/golden/make_golden.py

Fixtures (tests/golden/*.npz) hold, for seeded small problems, the inputs that
cannot be regenerated from the seed alone plus the reference's outputs:
  step_*   Hqp_IpLQDOCP::factor + ::step / Hqp_IpMatrix::solve on a random RHS
  ips_*    a full cold-started Hqp_IpsMehrotra solve (x, y, z, iteration count)
  docp.json  the hqp_docp Prg_DID example under several solver combinations
The reference cannot travel to the GPU box; these files can.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from hqp_b200.problem import (synth_lqdocp, add_random_stage_ineq, rhs_for,  # noqa: E402
                              add_stage_equalities)
from oracle import refharness as R  # noqa: E402

STEP_CASES = {
    # name: (nx, nu, K, bounds, general ineq rows/stage, fixed_x0)
    "step_n2m1K8": (2, 1, 8, True, 0, True),
    "step_n5m3K40": (5, 3, 40, True, 0, True),
    "step_n5m3K41_free": (5, 3, 41, True, 0, False),
    "step_n12m4K50": (12, 4, 50, True, 0, True),
    "step_n20m10K200": (20, 10, 200, True, 0, True),
    "step_n7m9K64_noineq": (7, 9, 64, False, 0, True),
    "step_n6m2K30_gen": (6, 2, 30, True, 2, True),
    "step_n4m4K25_gen_free": (4, 4, 25, False, 3, False),
}


# general stage equality rows (terminal constraints etc.): name ->
# (nx, nu, K, bounds, general ineq rows/stage, fixed_x0, stages with equality rows, rows per stage)
EQ_CASES = {
    "eqstep_terminal_n4m2K12": (4, 2, 12, True, 0, True, [12], 2),
    "eqstep_terminal_full_n5m2K20": (5, 2, 20, True, 0, True, [20], 5),
    "eqstep_mid_and_terminal_n3m2K16": (3, 2, 16, True, 1, True, [16, 7], 1),
    "eqstep_free_x0_n4m3K9": (4, 3, 9, False, 0, False, [9, 0], 2),
    "eqstep_n20m10K60": (20, 10, 60, True, 0, True, [60], 6),
    # 56 rows: above the 48 KB default shared-memory limit of the Schur inverse
    "eqstep_56rows_n20m10K60": (20, 10, 60, True, 0, True, [60, 52, 44, 36, 28, 20, 12], 8),
}


def make_eq_problem(nx, nu, K, bounds, gen, fixed, stages, rows):
    p = make_problem(nx, nu, K, bounds, gen, fixed)
    return add_stage_equalities(p, stages, rows, seed=33)


def make_problem(nx, nu, K, bounds, gen, fixed):
    p = synth_lqdocp(nx, nu, K, bounds=bounds)
    if gen:
        add_random_stage_ineq(p, rows_per_stage=gen, nnz_per_row=3, seed=11)
    if not fixed:
        p.fixed_x0 = False
        p.b = p.b[:K * nx].copy()
    return p


def main():
    only = sys.argv[1:]  # optional: regenerate just the named fixtures
    for name, cfg in STEP_CASES.items():
        if only and name not in only:
            continue
        p = make_problem(*cfg)
        z, w, r1, r2, r3, r4 = rhs_for(p, seed=99)
        qp = R.RefQP(p)
        M = R.RefMatrix("LQDOCP", qp)
        M.factor(z, w)
        dx, dy, dz, dw = M.step(z, w, r1, r2, r3, r4)
        sx, sy, sz, sw, res = M.solve(z, w, r1, r2, r3, r4)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), cfg=np.array(cfg, dtype=np.int64),
                            z=z, w=w, r1=r1, r2=r2, r3=r3, r4=r4, dx=dx, dy=dy, dz=dz, dw=dw,
                            sx=sx, sy=sy, sz=sz, sw=sw, res=res)
        print(name, "res", res)
        M.close()
        qp.close()
    for name, cfg in EQ_CASES.items():
        if only and name not in only:
            continue
        p = make_eq_problem(*cfg)
        z, w, r1, r2, r3, r4 = rhs_for(p, seed=77)
        qp = R.RefQP(p)
        M = R.RefMatrix("LQDOCP", qp)
        M.factor(z, w)
        sx, sy, sz, sw, res = M.solve(z, w, r1, r2, r3, r4)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), z=z, w=w, r1=r1, r2=r2, r3=r3,
                            r4=r4, sx=sx, sy=sy, sz=sz, sw=sw, res=res)
        print(name, "res", res)
        M.close()
        qp.close()
    if only:
        return
    for name, (nx, nu, K) in {"ips_n5m3K40": (5, 3, 40), "ips_n20m10K200": (20, 10, 200)}.items():
        p = synth_lqdocp(nx, nu, K)
        qp = R.RefQP(p)
        out = {}
        for mat in ("LQDOCP", "RedSpBKP"):
            r = R.ips_solve(qp, "Mehrotra", mat, 1e-9)
            out[mat] = r
            print(name, mat, r["iters"], r["result"], np.linalg.norm(r["x"]))
        r = out["LQDOCP"]
        np.savez_compressed(os.path.join(HERE, name + ".npz"), cfg=np.array([nx, nu, K]),
                            x=r["x"], y=r["y"], z=r["z"], iters=r["iters"],
                            iters_redspbkp=out["RedSpBKP"]["iters"],
                            result=np.array(r["result"]))
        qp.close()
    # Mehrotra on a problem with general inequality rows and a terminal equality
    p = make_eq_problem(6, 2, 30, True, 1, True, [30], 3)
    qp = R.RefQP(p)
    r = R.ips_solve(qp, "Mehrotra", "LQDOCP", 1e-9)
    print("ips_eq_n6m2K30", r["iters"], r["result"], np.linalg.norm(r["x"]))
    np.savez_compressed(os.path.join(HERE, "ips_eq_n6m2K30.npz"), x=r["x"], y=r["y"], z=r["z"],
                        iters=r["iters"], result=np.array(r["result"]))
    qp.close()
    docp = {}
    for key, (kmax, qps, mat) in {
            # "" keeps the as-shipped Hqp_IpsFranke instance (qp_eps 1e-9 set by
            # Hqp_SqpSolver's constructor, hqp/Hqp_SqpSolver.C:80)
            "K60_asShipped_Franke_LQDOCP": (60, "", "LQDOCP"),
            "K60_asShipped_Franke_RedSpBKP": (60, "", "RedSpBKP"),
            "K60_Mehrotra_LQDOCP": (60, "Mehrotra", "LQDOCP"),
            "K60_Mehrotra_RedSpBKP": (60, "Mehrotra", "RedSpBKP"),
            "K200_Mehrotra_LQDOCP": (200, "Mehrotra", "LQDOCP"),
            "K1000_Mehrotra_LQDOCP": (1000, "Mehrotra", "LQDOCP")}.items():
        docp[key] = R.docp_did(kmax, qps, mat)
        print(key, docp[key])
    with open(os.path.join(HERE, "docp.json"), "w") as f:
        json.dump(docp, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
