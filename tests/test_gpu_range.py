"""GPU test of the horizon-split kernels (hqpcu_range_*, lq_range.cuh): the stage
ranges of world = 2..4 ranks are run in lock-step in ONE process on one GPU (the
all-gathers become torch.stack), and the assembled solution must match the
full-horizon CPU oracle.  The multi-process NCCL version of the same protocol is
hqp_b200/dist.py:RangeSolver (CPU-tested over gloo in tests/test_dist_gloo.py,
timed by `bench.py --gpus N`)."""
import numpy as np
import pytest
import torch

from common import relerr
from hqp_b200.dist import CudaRangeEngine, local_vectors, scatter_solution, split_problem
from hqp_b200.problem import add_random_stage_ineq, rhs_for, synth_lqdocp
from oracle.portoracle import PortOracle

pytestmark = pytest.mark.gpu


def run_lockstep(p, world, vecs, nseg=0):
    dev = torch.device("cuda", 0)
    parts = split_problem(p, world)
    engines = [CudaRangeEngine(lp, rm, device=0, nseg=nseg) for lp, rm in parts]
    lv = [[torch.from_numpy(a.copy()).to(dev) if a.size else torch.zeros(1, dtype=torch.float64, device=dev)
           for a in local_vectors(p, rm, *vecs)] for _, rm in parts]
    g = torch.stack([e.factor_begin(v[0], v[1]).clone() for e, v in zip(engines, lv)])
    gpsi = torch.stack([e.factor_finish(g, r, world).clone() for r, e in enumerate(engines)])
    for e in engines:
        assert e.status() == 0
    gv = torch.stack([e.step_begin(*v[2:]).clone() for e, v in zip(engines, lv)])
    gx = torch.stack([e.step_mid(gv, gpsi, r, world).clone() for r, e in enumerate(engines)])
    dx, dy, dz, dw = np.zeros(p.N), np.zeros(p.me), np.zeros(p.m), np.zeros(p.m)
    for r, (e, (lp, rm)) in enumerate(zip(engines, parts)):
        out = e.step_finish(gx, gpsi, r, world)
        torch.cuda.synchronize()
        loc = [t.cpu().numpy() for t in out]
        loc[2], loc[3] = loc[2][:lp.m], loc[3][:lp.m]
        scatter_solution(p, rm, loc, dx, dy, dz, dw)
    for e in engines:
        e.close()
    return dx, dy, dz, dw


@pytest.mark.parametrize("dims,world,fixed,nseg", [
    ((4, 2, 23), 2, True, 0), ((4, 2, 23), 3, False, 1), ((20, 10, 200), 2, True, 0),
    ((20, 10, 403), 4, True, 0), ((12, 4, 64), 4, True, 1), ((5, 3, 37), 3, True, 4),
])
def test_split_horizon_matches_full_oracle(dims, world, fixed, nseg):
    p = synth_lqdocp(*dims)
    add_random_stage_ineq(p, rows_per_stage=1, nnz_per_row=3, seed=3)
    if not fixed:
        p.fixed_x0 = False
        p.b = p.b[:p.K * p.nx].copy()
    vecs = rhs_for(p, seed=9)
    o = PortOracle(p)
    o.factor(vecs[0], vecs[1])
    ref = o.step(*vecs[2:])
    o.close()
    mine = run_lockstep(p, world, vecs, nseg)
    for a, b in zip(mine, ref):
        assert relerr(a, b) < 1e-10
