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


# ---- row f4: the DOCP update over stage ranges (one hqpdocp handle per GPU).  The ranges need no
# exchange -- only a halo of one state -- so the N > 1 logic is the partition itself: local
# index maps of the bound tables, the halo stage that is read but not evaluated, and the
# group-wise concatenation of b and d.  Engine here: the plain-Python restatement.

def docp_worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from hqp_b200 import docpcuda as dc
    from oracle import docp_oracle
    p = dc.synthnl_problem(11, 5, 2, 3, 2, seed=9)
    x = p.x_init + 0.1 * np.random.default_rng(0).uniform(-1, 1, p.N)
    q = p.shard(rank, world)
    out = docp_oracle.update(q, p.x_slice(x, q.k_first, q.k_first + q.K), "fd")
    # the objective is the one quantity that needs a reduction
    f = torch.tensor([out["f"]], dtype=torch.float64)
    dist.all_reduce(f)
    gathered = [None] * world
    dist.all_gather_object(gathered, out)
    if rank == 0:
        full = dc.assemble_shards(p, [p.shard(r, world) for r in range(world)], gathered)
        np.savez(os.path.join(out_dir, "docp.npz"), f_allreduce=f.numpy(), **full)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_docp_update_over_stage_ranges_matches_full_horizon(world, tmp_path):
    from hqp_b200 import docpcuda as dc
    from hqp_b200.dist import stage_ranges
    from oracle import docp_oracle
    port = 31500 + os.getpid() % 2000 + world
    mp.spawn(docp_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    p = dc.synthnl_problem(11, 5, 2, 3, 2, seed=9)
    x = p.x_init + 0.1 * np.random.default_rng(0).uniform(-1, 1, p.N)
    want = docp_oracle.update(p, x, "fd")
    got = np.load(os.path.join(str(tmp_path), "docp.npz"))
    for key in ("b", "d", "g", "fx", "fu", "cx", "cu"):
        assert np.array_equal(got[key], want[key]), key
    assert abs(float(got["f_allreduce"][0]) - want["f"]) <= 1e-13 * abs(want["f"])
    # the ranges are the ones the KKT horizon split uses
    assert [(q.k_first, q.k_first + q.K) for q in (p.shard(r, world) for r in range(world))] == stage_ranges(p.K, world)
