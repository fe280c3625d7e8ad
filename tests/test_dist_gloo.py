"""CPU test of the N>1 path: the horizon-split exchange protocol
(hqp_b200/dist.py: partition, all-gathers, replicated boundary chains) on
world_size 2 and 3 over gloo, with the numpy range engine, against the
full-horizon CPU oracle."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from common import relerr  # noqa: E402
from hqp_b200.dist import (RangeSolver, local_vectors, scatter_solution, split_problem,  # noqa: E402
                           stage_ranges)
from hqp_b200.problem import add_random_stage_ineq, rhs_for, synth_lqdocp  # noqa: E402
from numpy_range_engine import NumpyRangeEngine  # noqa: E402
from oracle.portoracle import PortOracle  # noqa: E402


def build(fixed):
    p = synth_lqdocp(4, 2, 23)
    add_random_stage_ineq(p, rows_per_stage=1, nnz_per_row=3, seed=3)
    if not fixed:
        p.fixed_x0 = False
        p.b = p.b[:p.K * p.nx].copy()
    return p


def worker(rank, world, port, fixed, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    p = build(fixed)
    vecs = rhs_for(p, seed=9)
    lp, rm = split_problem(p, world)[rank]
    lv = [torch.from_numpy(a.copy()) for a in local_vectors(p, rm, *vecs)]
    solver = RangeSolver(NumpyRangeEngine(lp, rm), rank, world)
    solver.factor(lv[0], lv[1])
    out = solver.step(*lv[2:])
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), *[t.numpy() for t in out])
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,fixed", [(2, True), (3, True), (2, False)])
def test_horizon_split_matches_full_horizon_oracle(world, fixed, tmp_path):
    port = 29500 + os.getpid() % 2000 + world
    mp.spawn(worker, args=(world, port, fixed, str(tmp_path)), nprocs=world, join=True)
    p = build(fixed)
    z, w, r1, r2, r3, r4 = rhs_for(p, seed=9)
    o = PortOracle(p)
    o.factor(z, w)
    ref = o.step(r1, r2, r3, r4)
    o.close()
    dx, dy, dz, dw = np.zeros(p.N), np.zeros(p.me), np.zeros(p.m), np.zeros(p.m)
    for rank, (lp, rm) in enumerate(split_problem(p, world)):
        f = np.load(os.path.join(str(tmp_path), f"r{rank}.npz"))
        scatter_solution(p, rm, [f[k] for k in f.files], dx, dy, dz, dw)
    for a, b in zip((dx, dy, dz, dw), ref):
        assert relerr(a, b) < 1e-10


def test_stage_ranges_cover_the_horizon():
    for K, world in [(10, 3), (7, 7), (10000, 8), (5, 2)]:
        r = stage_ranges(K, world)
        assert r[0][0] == 0 and r[-1][1] == K
        assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
        assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1
