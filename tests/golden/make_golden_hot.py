"""Golden fixtures for the hot-started Mehrotra sequence (ipshot_*.npz): generated
by the UNMODIFIED reference (oracle/_ref) -- a cold-started solve followed by
hot-started solves (Hqp_IpsMehrotra::hot_start, hqp/Hqp_IpsMehrotra.C:330-352,
with the restart logic of solve(), :696-733) after the linear terms changed the
way consecutive SQP iterations change them.  Run here (needs /root/reference):
    python tests/golden/make_golden_hot.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from hqp_b200.problem import synth_lqdocp  # noqa: E402
from oracle import refharness as R  # noqa: E402


def sequence(p, scales, seed):
    rng = np.random.default_rng(seed)
    dc, db = rng.standard_normal(p.N), rng.standard_normal(p.me)
    db[p.K * p.nx:] *= 0.1
    cs = np.stack([p.c + s * dc for s in scales])
    bs = np.stack([p.b + s * db for s in scales])
    ds = np.stack([p.d for _ in scales])
    return cs, bs, ds


if __name__ == "__main__":
    for name, (nx, nu, K), scales in (("ipshot_n5m3K40", (5, 3, 40), (0.0, 0.02, 0.05, 2.0, 2.01)),
                                      ("ipshot_n20m10K100", (20, 10, 100), (0.0, 0.01, 0.03, 1.0))):
        p = synth_lqdocp(nx, nu, K)
        cs, bs, ds = sequence(p, scales, 11)
        qp = R.RefQP(p)
        out = R.ips_solve_seq(qp, cs, bs, ds)
        print(name, [(o["iters"], o["result"]) for o in out])
        np.savez_compressed(os.path.join(HERE, name + ".npz"), cfg=np.array([nx, nu, K]),
                            scales=np.array(scales), seed=11,
                            x=np.stack([o["x"] for o in out]), y=np.stack([o["y"] for o in out]),
                            z=np.stack([o["z"] for o in out]),
                            iters=np.array([o["iters"] for o in out]),
                            result=np.array([o["result"] for o in out]))
        qp.close()
