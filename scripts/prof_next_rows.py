"""Scratch: one BFGS update of config 2's Hessian and one grd_L / merit evaluation (for ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from hqp_b200 import hlcuda
from hqp_b200.ipcuda import IpCuda
from hqp_b200.problem import synth_lqdocp
K, n = 10000, 30
rng = np.random.default_rng(1)
M = rng.uniform(-1, 1, (K, n, n))
Q = (np.einsum("kij,kil->kjl", M, M) / n + 0.05 * np.eye(n)).ravel()
s, u = rng.uniform(-1, 1, K * n), rng.uniform(-1, 1, K * n)
for _ in range(2):
    hlcuda.bfgs_update([n] * K, Q, s, u, 1.0)
u2 = u.copy(); u2[: K * n // 2] *= -30.0   # strong negative curvature on half of the blocks: shifts needed
got, info = hlcuda.bfgs_update([n] * K, Q, s, u2, 1.0)
print("bfgs info", info)
p = synth_lqdocp(20, 10, 10000)
e = IpCuda(p); e.update()
y, z = rng.uniform(-1, 1, p.me), rng.uniform(0, 1, p.m)
for _ in range(2):
    e.sqp_grd_L(p.c, y, z)
    e.sqp_merit(0.0, p.c, rng.uniform(-1, 1, p.N), p.b, p.d, np.ones(p.me), np.ones(p.m))
e.close()
print("done")
