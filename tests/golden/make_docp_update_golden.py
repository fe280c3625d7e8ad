"""Golden vectors for row f4 (the DOCP update), made by the UNMODIFIED reference: Hqp_Docp's
setup / update / update_fbd driven through oracle/_ref for Prg_DID and for the synthetic model
(oracle/prg_synthnl.cpp).  Run in the build container (needs oracle/_ref):
    python tests/golden/make_docp_update_golden.py
Writes tests/golden/docp_update_<case>.npz: the iterate x and f, b, d, c (qp->c), dense A, C
after update(), and f, b, d after update_fbd() at a second iterate."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from hqp_b200 import docpcuda as dc  # noqa: E402
from oracle import refharness as rh  # noqa: E402

CASES = {
    "did_k12": lambda: dc.did_problem(12, True),
    "did_k60_nocns": lambda: dc.did_problem(60, False),
    "synthnl_n4m2c1K6": lambda: dc.synthnl_problem(6, 4, 2, 1, 0),
    "synthnl_n6m3c3K5": lambda: dc.synthnl_problem(5, 6, 3, 3, 2),
    "synthnl_n20m10c1K8": lambda: dc.synthnl_problem(8, 20, 10, 1, 1),
}


def case_iterates(p):
    rng = np.random.default_rng(5)
    return p.x_init + 0.1 * rng.uniform(-1, 1, p.N), p.x_init + 0.3 * rng.uniform(-1, 1, p.N)


if __name__ == "__main__":
    out = os.path.dirname(os.path.abspath(__file__))
    for name, make in CASES.items():
        p = make()
        r = rh.RefDocp(p)
        x1, x2 = case_iterates(p)
        u = r.update(x1)
        v = r.update(x2, fbd_only=True)
        np.savez_compressed(os.path.join(out, f"docp_update_{name}.npz"), x=x1, f=u["f"], b=u["b"], d=u["d"],
                            c=u["c"], A=u["A"], C=u["C"], x2=x2, f2=v["f"], b2=v["b"], d2=v["d"])
        r.close()
        print(name, p.N, p.me, p.m)
