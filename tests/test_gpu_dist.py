"""Horizon split over several GPUs, one PROCESS per GPU, NCCL called from inside
libhqpcuda.so (hqpcu_comm_init): the assembled factor / step / refined solve /
residuum / device-resident Mehrotra solve of 2 and 4 ranks against the
full-horizon CPU oracle and the single-GPU engine.  Needs >= 2 CUDA devices
(`gpurun --gpus 2|4 -- python -m pytest tests/test_gpu_dist.py`); skipped on a
one-GPU box."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, cfg, graphs):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["HQPCU_GRAPHS"] = "1" if graphs else "0"
    import torch
    import torch.distributed as dist
    from hqp_b200.dist import (DistIpCuda, local_vectors, scatter_solution, split_problem)
    from hqp_b200.ipcuda import IpCuda
    from hqp_b200.problem import add_random_stage_ineq, rhs_for, synth_lqdocp
    from oracle.portoracle import PortOracle

    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank,
                            world_size=world, device_id=torch.device("cuda", rank))
    nx, nu, K, gen, fixed = cfg
    p = synth_lqdocp(nx, nu, K, seed=77)
    if gen:
        add_random_stage_ineq(p, rows_per_stage=gen, nnz_per_row=3, seed=5)
    if not fixed:
        p.fixed_x0 = False
        p.b = p.b[:K * nx].copy()
    z, w, r1, r2, r3, r4 = rhs_for(p, seed=9)
    lp, rm = split_problem(p, world)[rank]
    e = DistIpCuda(lp, rank, world, device=rank)
    assert e.comm_info()[:2] == (rank, world) and e.comm_info()[2] == p.m
    e.update()
    lz, lw, l1, l2, l3, l4 = local_vectors(p, rm, z, w, r1, r2, r3, r4)

    def relerr(a, b):
        return float(np.max(np.abs(a - b)) / max(1e-300, np.max(np.abs(b)))) if a.size else 0.0

    def assemble(local):
        """all ranks' local solutions -> global vectors (on every rank)"""
        parts = [None] * world
        dist.all_gather_object(parts, [np.asarray(v) for v in local])
        dx, dy, dz, dw = np.zeros(p.N), np.zeros(p.me), np.zeros(p.m), np.zeros(p.m)
        for r in range(world):
            _, rmr = split_problem(p, world)[r]
            scatter_solution(p, rmr, parts[r], dx, dy, dz, dw)
        return dx, dy, dz, dw

    o = PortOracle(p)
    o.factor(z, w)
    ref = o.step(r1, r2, r3, r4)
    for rep in range(2):  # second round replays the captured graphs
        e.factor(lz, lw)
        mine = assemble(e.step(l1, l2, l3, l4))
        for a, b, key in zip(mine, ref, ("dx", "dy", "dz", "dw")):
            assert relerr(a, b) < 1e-10, (key, relerr(a, b), rank, rep)
    # refined solve: same control flow on every rank, residual below mat_eps
    sx, sy, sz, sw, res, nsteps = e.solve(l1, l2, l3, l4)
    got = assemble((sx, sy, sz, sw))
    for a, b in zip(got, ref):
        assert relerr(a, b) < 1e-10
    assert res <= 1e-10 and nsteps == 1
    # residuum of arbitrary (perturbed) vectors equals the oracle's on the global ones
    rng = np.random.default_rng(3)
    pert = [v + 1e-3 * rng.standard_normal(v.shape) for v in ref]
    lpert = local_vectors(p, rm, z, w, pert[0], pert[1], pert[2], pert[3])[2:]
    # (local_vectors cuts x / dyn-row / ineq-row vectors the same way for solutions)
    rg = e.residuum(l1, l2, l3, l4, *lpert)
    ro = o.residuum(r1, r2, r3, r4, *pert)
    assert abs(rg - ro) <= 1e-11 * max(1.0, ro), (rg, ro)
    # device-resident Mehrotra: iteration count and solution of the single-GPU engine
    r = e.mehrotra_solve(c=lp.c, b=lp.b, d=lp.d)
    one = None
    if rank == 0:
        s = IpCuda(p, device=0)
        s.update()
        one = s.mehrotra_solve()
        s.close()
    box = [one]
    dist.broadcast_object_list(box, src=0)
    one = box[0]
    assert r["iters"] == one["iters"] and r["result"] == one["result"] == "optimal", (r["iters"], one["iters"])
    parts = [None] * world
    dist.all_gather_object(parts, [r["x"], r["y"], r["z"], r["w"]])
    gx, gy, gz, gw = np.zeros(p.N), np.zeros(p.me), np.zeros(p.m), np.zeros(p.m)
    for rr in range(world):
        _, rmr = split_problem(p, world)[rr]
        scatter_solution(p, rmr, parts[rr], gx, gy, gz, gw)
    assert relerr(gx, one["x"]) < 1e-9 and relerr(gz, one["z"]) < 1e-8
    o.close()
    e.close()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(_ngpu() < 2, reason="needs >= 2 CUDA devices")
@pytest.mark.parametrize("graphs", [True, False])
@pytest.mark.parametrize("cfg", [(20, 10, 400, 0, 1), (12, 4, 230, 2, 0), (40, 10, 260, 0, 1)])
@pytest.mark.parametrize("world", [2, 4])
def test_split_horizon_matches_oracle(world, cfg, graphs):
    if _ngpu() < world:
        pytest.skip(f"needs {world} CUDA devices")
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(world, _free_port(), cfg, graphs), nprocs=world, join=True)


# ---- several GPUs behind ONE handle of one process (hqpcu_dims::ngpu, mat_ngpu) ------
@pytest.mark.skipif(_ngpu() < 2, reason="needs >= 2 CUDA devices")
@pytest.mark.parametrize("cfg", [(20, 10, 400, 0, 1), (12, 4, 230, 2, 0)])
@pytest.mark.parametrize("ngpu", [2, 4])
def test_single_process_dispatcher_matches_oracle(ngpu, cfg):
    if _ngpu() < ngpu:
        pytest.skip(f"needs {ngpu} CUDA devices")
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from common import relerr
    from hqp_b200.ipcuda import IpCuda
    from hqp_b200.problem import add_random_stage_ineq, rhs_for, synth_lqdocp
    from oracle.portoracle import PortOracle
    nx, nu, K, gen, fixed = cfg
    p = synth_lqdocp(nx, nu, K, seed=77)
    if gen:
        add_random_stage_ineq(p, rows_per_stage=gen, nnz_per_row=3, seed=5)
    if not fixed:
        p.fixed_x0 = False
        p.b = p.b[:K * nx].copy()
    z, w, r1, r2, r3, r4 = rhs_for(p, seed=9)
    o = PortOracle(p)
    o.factor(z, w)
    ref = o.step(r1, r2, r3, r4)
    e = IpCuda(p, ngpu=ngpu)         # full-length vectors in and out
    e.update()
    for _ in range(2):
        e.factor(z, w)
        for a, b in zip(e.step(r1, r2, r3, r4), ref):
            assert relerr(a, b) < 1e-10
    sx, sy, sz, sw, res, nsteps = e.solve(r1, r2, r3, r4)
    assert res <= 1e-10 and nsteps == 1 and relerr(sx, ref[0]) < 1e-10
    rng = np.random.default_rng(3)
    pert = [v + 1e-3 * rng.standard_normal(v.shape) for v in ref]
    assert abs(e.residuum(r1, r2, r3, r4, *pert) - o.residuum(r1, r2, r3, r4, *pert)) <= 1e-11
    one = IpCuda(p)
    one.update()
    for solver in ("mehrotra_solve", "franke_solve"):
        a, b = getattr(e, solver)(), getattr(one, solver)()
        assert a["iters"] == b["iters"] and a["result"] == b["result"] == "optimal", solver
        assert relerr(a["x"], b["x"]) < 1e-9
    one.close(); e.close(); o.close()


def _plugin_worker(ngpu, q):
    """the unmodified Hqp_IpsMehrotra of the reference driving Hqp_IpCuda (plugin) on a
    synthetic QP, in a fresh process (mat_ngpu is read from the environment by the
    test harness when the solver module is created)"""
    sys.path.insert(0, ROOT)
    if ngpu > 1:
        os.environ["HQP_MAT_NGPU"] = str(ngpu)
    from hqp_b200.problem import synth_lqdocp
    from oracle import refharness
    refharness.load_plugin(os.path.join(ROOT, "hqp_b200", "lib", "libhqp_ipcuda_plugin.so"))
    p = synth_lqdocp(20, 10, 300)
    r = refharness.ips_solve(refharness.RefQP(p), "Mehrotra", "Cuda", 1e-9)
    q.put((r["iters"], r["result"], r["x"]))


@pytest.mark.skipif(_ngpu() < 2, reason="needs >= 2 CUDA devices")
def test_reference_ip_solver_through_the_plugin_with_mat_ngpu():
    """qp_mat_solver Cuda + mat_ngpu 2 under the UNMODIFIED Hqp_IpsMehrotra: same
    iteration count and solution as with one GPU and as with Hqp_IpLQDOCP.  (The
    docp example itself fixes its terminal state with general equality rows, which a
    split horizon does not carry.)"""
    import multiprocessing as mp
    from oracle import refharness
    if not refharness.available():
        pytest.skip("compiled reference not present")
    ctx = mp.get_context("spawn")
    out = []
    for ngpu in (1, 2):
        q = ctx.Queue()
        pr = ctx.Process(target=_plugin_worker, args=(ngpu, q))
        pr.start()
        out.append(q.get(timeout=300))
        pr.join()
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from common import relerr
    from hqp_b200.problem import synth_lqdocp
    ref = refharness.ips_solve(refharness.RefQP(synth_lqdocp(20, 10, 300)), "Mehrotra", "LQDOCP", 1e-9)
    for it, res, x in out:
        assert it == ref["iters"] and res == ref["result"] == "optimal"
        assert relerr(x, ref["x"]) < 1e-8
