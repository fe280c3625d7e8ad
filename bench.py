#!/usr/bin/env python
"""KKT factor+solve throughput of the Hqp_IpCuda engine (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                  [--workload c2|c3|c4|c5|c5s] [--no-extra]

One "step" = one unit of interior-point KKT work on one synthetic LQ-DOCP
horizon: 1 factor + 2 step (Mehrotra predictor + corrector).
`value` = stages/s = stages of the whole job / step time, device-timed with
inputs resident in HBM; `e2e` = the same through the host-pointer C-ABI calls
the Hqp_IpCuda plugin makes (H2D/D2H inside the timed region; pinned buffers for
the contract's number, pageable ones -- what Meschach VECs are -- next to it).

Workloads (BASELINE.json configs):
  c2   nx=20 nu=10 K=10,000, one instance  (the configuration the metric is quoted on)
       N > 1: weak scaling, ONE horizon of N*10,000 stages split into N ranges
  c3   4096 instances nx=12 nu=4 K=50, one CTA per instance (N > 1: instances sharded)
  c4   nx=200 nu=50 K=2,000: stage blocks in a global workspace, FP64 DMMA block
       products staged through shared memory (the flop-bound configuration)
  c5   nx=40 nu=10 K=1,000,000 (BASELINE configs[4]) generated on the device;
       N > 1: strong scaling, the horizon split into N contiguous stage ranges
  c5s  nx=40 nu=10 K=100,000 host-generated slice of the same shape
The default line (workload c2) carries the other configurations as sub-objects
(`c5`, and at N = 1 `c4`, `c3`, and the next rows of SURVEY section 8: `hl_bfgs` = the
block-diagonal BFGS update, `sqp_ops` = grd_L / merit functions, `docp_update` = the stage loop of Hqp_Docp::update) unless --no-extra is given.

Multi-GPU: one process per GPU (torchrun); the horizon split lives in
libhqpcuda.so (hqpcu_comm_init): NCCL all-gathers of the boundary elements /
vectors are graph nodes between the kernels.  Every run ends with a refined
solve whose KKT residual (max over ranks) is asserted before the line is printed.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    "c2": dict(nx=20, nu=10, K=10000, batch=1, scaling="weak",
               name="synthetic LQ-DOCP nx=20 nu=10 K=10000 (BASELINE configs[1])"),
    "c3": dict(nx=12, nu=4, K=50, batch=4096, scaling="weak",
               name="batched MPC QPs 4096 x (nx=12 nu=4 K=50) (BASELINE configs[2])"),
    "c5": dict(nx=40, nu=10, K=1000000, batch=1, scaling="strong", device_generated=True,
               name="long-horizon LQ-DOCP nx=40 nu=10 K=1000000 (BASELINE configs[4]), "
                    "horizon split over the GPUs"),
    "c4": dict(nx=200, nu=50, K=2000, batch=1, scaling="weak", bound="fp64",
               name="large-block LQ-DOCP nx=200 nu=50 K=2000 (BASELINE configs[3])"),
    "c5s": dict(nx=40, nu=10, K=100000, batch=1, scaling="weak",
                name="long-horizon LQ-DOCP nx=40 nu=10 K=100000 (slice of BASELINE configs[4])"),
}
# dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of seg_riccati_kernel
# from the committed `ncu --set full` capture; null where no capture exists.
NCU_TRAFFIC = {"c2": (124.024832e6 + 50.681344e6, "profiles/r01_ncu_full_k1k3_v7.md (K3 section)")}
METRIC = "LQ-DOCP KKT factor+solve stages/s"
UNIT = "stages/s"
UNIT_OF_WORK = "1 factor + 2 step"
# FP64 DMMA peak measured with scripts/mb/mb_dmma.cu on this pool's B200s
# (profiles/r02_mb_dmma.md); MEASURED_PEAKS.json carries HBM and bf16 only
FP64_TFLOPS = 37.2


def algorithmic_bytes(nx, nu, mc):
    """bytes per stage of one factor / one step."""
    bf = 8 * ((nx + nu) * (nx + nu + 1) // 2 + nx * (nx + nu) + 2 * mc) + \
        8 * (nx * nx + 2 * nx * nu + nu * nu)
    bs = 8 * (2 * nx * nx + 3 * nx * nu + nu * nu) + 8 * (4 * nx + 3 * nu + 6 * mc)
    return bf, bs


def algorithmic_flops(nx, nu):
    ff = 4 * nx ** 3 + 6 * nx * nx * nu + 3 * nx * nu * nu + nu ** 3 / 3.0
    fs = 8 * nx * nx + 8 * nx * nu + 2 * nu * nu
    return ff, fs


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f).get("hbm_gbs", 6650.0), "measured"
    return 6650.0, "fallback"


def config_of(wl, world):
    """identical in both arms (the driver compares it)"""
    K, batch = wl["K"], wl["batch"]
    per_gpu = K * batch // world if wl["scaling"] == "strong" else K * batch
    return {"workload": wl["name"], "unit_of_work": UNIT_OF_WORK, "stages_per_gpu": per_gpu,
            "scaling": wl["scaling"]}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx,
                "reasons": sorted(reasons), "samples": len(sm)}


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


# ------------------------------------------------------------------ reference --
def time_reference(wl, steps, warmup, max_stages=None):
    """Reference CPU path on the host cores: the UNMODIFIED Hqp_IpLQDOCP
    (oracle/_ref) when it travelled with the snapshot, else the C port."""
    from hqp_b200.problem import synth_lqdocp, synth_rhs
    nx, nu, K, batch = wl["nx"], wl["nu"], wl["K"], wl["batch"]
    Ks = K if max_stages is None else min(K, max_stages)
    p = synth_lqdocp(nx, nu, Ks)
    z, w, r1, r2, r3, r4 = synth_rhs(p)
    from oracle import refharness
    times = []
    if refharness.available():
        kind = "reference"
        qp = refharness.RefQP(p)
        M = refharness.RefMatrix("LQDOCP", qp)
        for i in range(warmup + steps):
            # timed inside the harness with clock_gettime(CLOCK_MONOTONIC) around
            # Hqp_IpLQDOCP::factor and 2 x ::step (oracle/ref_harness.cpp:ref_mat_time)
            tf, ts = M.time(z, w, r1, r2, r3, r4, reps=1, nstep=2)
            if i >= warmup:
                times.append(tf + 2 * ts)
    else:
        kind = "port"
        from oracle.portoracle import PortOracle
        o = PortOracle(p)
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            o.factor(z, w)
            o.step(r1, r2, r3, r4)
            o.step(r1, r2, r3, r4)
            t = time.perf_counter() - t0
            if i >= warmup:
                times.append(t)
    t_step = float(np.mean(times))
    sample = (f"{Ks} of {K * batch} stages (one instance), {steps} x ({UNIT_OF_WORK}), "
              f"single thread: the reference path is not thread-safe")
    return dict(value=Ks / t_step, unit=UNIT, cores=1, kind=kind, sample=sample,
                ms_per_unit=1e3 * t_step, stages=Ks)


def _c3_worker(args):
    """one host process of the batched CPU baseline: its share of instances, one after the other"""
    nx, nu, K, n_inst, seed = args
    sys.path.insert(0, ROOT)
    from hqp_b200.problem import synth_lqdocp, synth_rhs
    from oracle import refharness
    p = synth_lqdocp(nx, nu, K, seed=seed)
    rhs = synth_rhs(p)
    if refharness.available():
        qp = refharness.RefQP(p)
        M = refharness.RefMatrix("LQDOCP", qp)
        M.time(*rhs, reps=1, nstep=2)
        t0 = time.perf_counter()
        for _ in range(n_inst):
            M.time(*rhs, reps=1, nstep=2)
        return time.perf_counter() - t0
    from oracle.portoracle import PortOracle
    o = PortOracle(p)
    t0 = time.perf_counter()
    for _ in range(n_inst):
        o.factor(rhs[0], rhs[1])
        o.step(*rhs[2:])
        o.step(*rhs[2:])
    return time.perf_counter() - t0


def time_reference_batched(wl, per_proc=256):
    """one worker process per host core, each looping over its share
    of the independent instances with the single-threaded reference path."""
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    ctx = mp.get_context("spawn")
    with ctx.Pool(cores) as pool:
        t0 = time.perf_counter()
        pool.map(_c3_worker, [(wl["nx"], wl["nu"], wl["K"], per_proc, 100 + i) for i in range(cores)])
        wall = time.perf_counter() - t0
    # (wall includes process start and problem set-up: a lower bound on the baseline's speed)
    return dict(value=cores * per_proc * wl["K"] / wall, unit=UNIT, cores=cores,
                kind="reference", sample=f"{cores} processes x {per_proc} instances of "
                f"{wl['batch']} (wall clock incl. process start-up)")


def run_reference(args, wl):
    rank, world, _ = dist_env()
    if rank != 0:
        return
    cap = (100 if wl["nx"] > 64 else 10000) if wl["batch"] == 1 else wl["K"]
    cb = time_reference(wl, args.steps, args.warmup, max_stages=cap)
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT,
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": cb["ms_per_unit"], "higher_is_better": True, "scaling": wl["scaling"],
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_of(wl, max(args.gpus, 1)),
            "details": {"timed_stages": cb["stages"]},
            "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------- ours --
class Structure:
    """what IpCuda needs of a problem whose matrices are generated on the device"""

    def __init__(self, nx, nu, K, fixed_x0, last):
        from hqp_b200.problem import LQProblem, box_bounds_on_u
        ptr, col, val, d = box_bounds_on_u(nx, nu, K)
        self.prob = LQProblem(nx, nu, K, None, None, None, None, None, fixed_x0=fixed_x0,
                              ineq_ptr=ptr, ineq_col=col, ineq_val=val, d=d)
        self.last = last


def device_generated_engine(wl, rank, world, local, nseg, group):
    """C5: the rank's stage range of a K-stage horizon, matrices drawn on the device
    (torch Philox, seeded per rank) chunk by chunk and handed to the library with
    hqpcu_update_stages_dev -- no host copy, no second full device copy."""
    import torch
    from hqp_b200.dist import DistIpCuda, stage_ranges
    nx, nu, K = wl["nx"], wl["nu"], wl["K"]
    nm = nx + nu
    k0, k1 = stage_ranges(K, world)[rank]
    Kl = k1 - k0
    st = Structure(nx, nu, Kl, fixed_x0=(rank == 0), last=(rank == world - 1))
    t0 = time.perf_counter()
    eng = DistIpCuda(st.prob, rank, world, device=local, nseg=nseg, group=group)
    dev = torch.device("cuda", local)
    g = torch.Generator(device=dev)
    g.manual_seed(1234 + rank)
    f64 = dict(dtype=torch.float64, device=dev)
    chunk = 50000
    eye_m = torch.eye(nm, **f64)
    eye_x = torch.eye(nx, **f64)
    for c0 in range(0, Kl + 1, chunk):
        nq = min(chunk, Kl + 1 - c0)
        nf = min(chunk, Kl - c0) if c0 < Kl else 0
        M = torch.rand(nq, nm, nm, generator=g, **f64) * 2 - 1
        Q = torch.bmm(M.transpose(1, 2), M) / nm + 0.1 * eye_m      # SPD, full Hxu block
        if c0 + nq == Kl + 1 and not st.last:
            Q[-1].zero_()       # the trailing state block belongs to the next range
        fx = fu = None
        if nf:
            fx = eye_x + 0.1 * (torch.rand(nf, nx, nx, generator=g, **f64) * 2 - 1) / nx ** 0.5
            fu = torch.rand(nf, nx, nu, generator=g, **f64) * 2 - 1
        torch.cuda.synchronize()
        eng.update_stages_dev(c0, nq, Q.data_ptr(), nf, fx.data_ptr() if nf else 0,
                              fu.data_ptr() if nf else 0)
        torch.cuda.synchronize()
        del M, Q, fx, fu
    cv = torch.from_numpy(st.prob.ineq_val).to(dev)
    eng.update_ineq_dev(cv.data_ptr())
    torch.cuda.synchronize()
    p = st.prob
    dvec = [torch.rand(p.m, generator=g, **f64) + 0.5, torch.rand(p.m, generator=g, **f64) + 0.5,
            torch.rand(p.N, generator=g, **f64) * 2 - 1, torch.rand(p.me, generator=g, **f64) * 2 - 1,
            torch.rand(p.m, generator=g, **f64) * 2 - 1, torch.rand(p.m, generator=g, **f64) * 2 - 1]
    if not st.last:
        dvec[2][Kl * nm:].zero_()   # r1 of the trailing state block: owned by the next range
    setup_ms = 1e3 * (time.perf_counter() - t0)
    return eng, p, dvec, setup_ms


def run_workload(args, key, steps, warmup, dist_group, extras=True):
    """-> the JSON line (dict) on rank 0, None elsewhere"""
    import torch
    import torch.distributed as dist
    from hqp_b200 import ipcuda as _ic
    from hqp_b200.ipcuda import IpCuda
    from hqp_b200.problem import synth_lqdocp, synth_rhs

    wl = WORKLOADS[key]
    rank, world, local = dist_env()
    dev = torch.device("cuda", local)
    nx, nu, K, batch = wl["nx"], wl["nu"], wl["K"], wl["batch"]
    update_ms = update_sparse_ms = None
    sharded_instances = world > 1 and batch > 1
    f64 = dict(dtype=torch.float64, device=dev)

    if wl.get("device_generated"):
        eng, lp, dvec, setup_ms = device_generated_engine(wl, rank, world, local, args.nseg,
                                                          dist_group)
        p = lp
        Kg = K  # stages of the whole job
        host = None
        parallelism = (f"horizon split, {world} contiguous stage ranges of {K // world} stages; "
                       "NCCL all-gathers inside the library's CUDA graphs (2 per factor, 2 per step)"
                       if world > 1 else "single")
    elif world == 1 or sharded_instances:
        # one horizon (or one batch of instances) on the GPU; with N > 1 GPUs a batch
        # workload shards its independent instances: `batch` per rank, no collective
        # in the KKT path
        p = synth_lqdocp(nx, nu, K, seed=1234)
        z, w, r1, r2, r3, r4 = synth_rhs(p, seed=4321)
        eng = IpCuda(p, batch=batch, device=local, nseg=args.nseg)

        def rep(a):
            return np.ascontiguousarray(np.tile(a, batch)) if batch > 1 else a

        if batch > 1:
            eng.update(Q=np.broadcast_to(p.Q, (batch,) + p.Q.shape),
                       fx=np.broadcast_to(p.fx, (batch,) + p.fx.shape),
                       fu=np.broadcast_to(p.fu, (batch,) + p.fu.shape),
                       ineq_val=np.broadcast_to(p.ineq_val, (batch,) + p.ineq_val.shape))
        else:
            t0 = time.perf_counter()
            eng.update()
            update_ms = 1e3 * (time.perf_counter() - t0)
            # the same update from the sparse values (map registered
            # once, values uploaded + scattered on the device per SQP iteration)
            vals, dst, dst2 = p.value_map()
            eng.set_value_map(dst, dst2)
            eng.update_values(vals)
            t0 = time.perf_counter()
            eng.update_values(vals)
            update_sparse_ms = 1e3 * (time.perf_counter() - t0)
            del dst, dst2
        host = [rep(a) for a in (z, w, r1, r2, r3, r4)]
        lp = p
        Kg = K * batch * world
        parallelism = (f"{world} x {batch} independent instances, sharded over the ranks, no collective"
                       if sharded_instances else "single")
    else:
        # horizon split, weak: ONE horizon of world*K stages, contiguous
        # stage ranges per rank, exchanges inside the library
        from hqp_b200.dist import DistIpCuda, local_vectors, split_problem
        pg = synth_lqdocp(nx, nu, K * world, seed=1234)
        gvec = synth_rhs(pg, seed=4321)
        lp, rm = split_problem(pg, world)[rank]
        host = [np.ascontiguousarray(a) for a in local_vectors(pg, rm, *gvec)]
        del pg, gvec
        t0 = time.perf_counter()
        eng = DistIpCuda(lp, rank, world, device=local, nseg=args.nseg, group=dist_group)
        eng.update()
        update_ms = 1e3 * (time.perf_counter() - t0)
        p = lp
        Kg = K * world
        parallelism = (f"horizon split, {world} contiguous stage ranges of {K} stages; NCCL "
                       "all-gathers inside the library's CUDA graphs (2 per factor, 2 per step)")
    stream = torch.cuda.current_stream()
    eng.set_stream(stream.cuda_stream)
    if host is not None:
        pinned = [torch.from_numpy(np.array(a)).pin_memory() for a in host]
        dvec = [t.to(dev) for t in pinned]
    N, me, m = lp.N * batch, lp.me * batch, lp.m * batch
    outs = [torch.empty(n, **f64) for n in (N, me, max(m, 1), max(m, 1))]
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)  # > 126 MB L2
    L = _ic.lib()

    def unit_dev():
        eng.factor_dev(dvec[0].data_ptr(), dvec[1].data_ptr())
        for _ in range(2):
            eng.step_dev(*[t.data_ptr() for t in dvec[2:]], *[t.data_ptr() for t in outs])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up ---------------------------------------------------------
    # (clocks are sampled by rank 0 only, from before the warm-up until after the e2e
    #  loops: the timed region of a 0.7 ms unit is shorter than nvidia-smi's period)
    sampler = ClockSampler(local) if rank == 0 and not os.environ.get("HQP_BENCH_NOCLOCKS") else None
    if sampler:
        sampler.start()
    # a split horizon gets a few extra untimed units: the first replays of graphs
    # that contain NCCL nodes still set up channels
    for _ in range(warmup + (5 if world > 1 else 0)):
        flush.zero_()
        unit_dev()
    st = eng.sync_status()
    if st != 0:
        raise RuntimeError(f"factor status {st}")

    # ---- timed region: device-resident inputs --------------------------------
    barrier()
    launches0 = eng.launches
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
          for _ in range(steps)]
    for e0, e1 in ev:
        flush.zero_()            # L2 flush between timed iterations (not timed)
        e0.record(stream)
        unit_dev()
        e1.record(stream)
    barrier()
    gpu_launches = eng.launches - launches0
    ms = [e0.elapsed_time(e1) for e0, e1 in ev]
    ms_step = float(np.mean(ms))
    ms_spread = {"min": float(np.min(ms)), "median": float(np.median(ms)), "max": float(np.max(ms))}

    # ---- correctness of what was timed: refined solve, KKT residual (max over the
    # ranks, all-reduced inside the library) -- asserted before anything is printed
    res, nsolve = eng.solve_dev(*[t.data_ptr() for t in dvec[2:]], *[t.data_ptr() for t in outs])
    torch.cuda.synchronize()
    if not (res <= 1e-9):
        raise RuntimeError(f"KKT residual {res} after the refined solve (rank {rank})")

    # ---- e2e: host buffers through the plugin-facing C ABI ----------------------
    e2e = {}
    if host is not None:
        hout = [torch.empty(n, dtype=torch.float64).pin_memory() for n in (N, me, max(m, 1), max(m, 1))]
        for kind in ("pinned", "pageable"):
            if kind == "pinned":
                hp, ho = [t.numpy() for t in pinned], [t.numpy() for t in hout]
            else:   # plain malloc'ed memory: what the plugin passes (VEC::ve)
                hp = [np.array(t.numpy(), copy=True) for t in pinned]
                ho = [np.empty(n) for n in (N, me, max(m, 1), max(m, 1))]

            def unit_host():
                _ic._check(L.hqpcu_factor(eng.h, _ic._hp(hp[0]), _ic._hp(hp[1])), "factor")
                for _ in range(2):
                    _ic._check(L.hqpcu_step(eng.h, *[_ic._hp(a) for a in hp[2:]],
                                            *[_ic._hp(a) for a in ho]), "step")
            unit_host()
            barrier()
            t0 = time.perf_counter()
            for _ in range(steps):
                unit_host()
            torch.cuda.synchronize()
            e2e[kind] = (time.perf_counter() - t0) / steps
            barrier()
    else:
        # device-generated workload: the host side of the same calls is the QP's own
        # vectors; time the C-ABI host entry points on pinned copies of the inputs
        hp = [t.cpu().pin_memory().numpy() for t in dvec]
        ho = [torch.empty(n, dtype=torch.float64).pin_memory().numpy()
              for n in (N, me, max(m, 1), max(m, 1))]

        def unit_host():
            _ic._check(L.hqpcu_factor(eng.h, _ic._hp(hp[0]), _ic._hp(hp[1])), "factor")
            for _ in range(2):
                _ic._check(L.hqpcu_step(eng.h, *[_ic._hp(a) for a in hp[2:]],
                                        *[_ic._hp(a) for a in ho]), "step")
        unit_host()
        barrier()
        t0 = time.perf_counter()
        for _ in range(max(2, steps // 2)):
            unit_host()
        torch.cuda.synchronize()
        e2e["pinned"] = (time.perf_counter() - t0) / max(2, steps // 2)
        barrier()
        del hp, ho

    clocks = sampler.stop() if sampler else None
    # ---- per-kernel CUDA-event times of the same unit (roofline section) ------
    eng.profile(True)
    nprof = min(steps, 5)
    for _ in range(nprof):
        flush.zero_()
        unit_dev()
    prof = eng.profile_read()
    eng.profile(False)

    # ---- whole QP solve on the device (Hqp_IpsMehrotra restated, SURVEY 8d:
    # "full-IP-iteration time including vector kernels") -- reported, not the metric
    ip_solve = None
    if batch == 1 and extras:
        try:
            if wl.get("device_generated"):
                rng = np.random.default_rng(100 + rank)
                cc = rng.uniform(-1, 1, lp.N)
                if rank < world - 1:
                    cc[(lp.K) * (nx + nu):] = 0.0
                bb = 0.01 * rng.uniform(-1, 1, lp.me)
                kw = dict(c=cc, b=bb, d=np.ones(lp.m))
            elif world > 1:
                kw = dict(c=lp.c, b=lp.b, d=lp.d)
            else:
                kw = {}
            if not wl.get("device_generated"):
                eng.mehrotra_solve(**kw)                      # warm (graphs captured)
            s0 = eng.solve_stats()
            barrier()
            t0 = time.perf_counter()
            r = eng.mehrotra_solve(**kw)
            dt = time.perf_counter() - t0
            s1 = eng.solve_stats()
            ip_solve = {"result": r["result"], "iterations": r["iters"], "ms_total": 1e3 * dt,
                        "ms_per_iteration": 1e3 * dt / max(r["iters"], 1),
                        "stages_per_s": Kg * r["iters"] / dt, "gap": r["gap"],
                        "mean_kkt_steps_per_refined_solve": (s1[1] - s0[1]) / max(s1[0] - s0[0], 1),
                        "note": "hqpcu_mehrotra_solve: cold start + IP iterations, host c/b/d in, "
                                "x/y/z/w out; each iteration = 1 factor + 2 refined solves + vector "
                                "kernels" + ("; IP scalars all-reduced over the ranks" if world > 1 else "")}
            # the same QP through the unmodified reference on the host
            # (Hqp_IpsMehrotra + Hqp_IpLQDOCP, oracle/_ref), bounded to <= 10^4 stages
            if rank == 0 and world == 1 and K <= 10000 and not wl.get("device_generated"):
                from oracle import refharness
                if refharness.available():
                    rr = refharness.ips_solve(refharness.RefQP(p), "Mehrotra", "LQDOCP", 1e-9)
                    ip_solve["reference_cpu"] = {
                        "iterations": rr["iters"], "result": rr["result"],
                        "ms_total": 1e3 * rr["seconds"],
                        "x_relative_difference": float(np.max(np.abs(rr["x"] - r["x"])) /
                                                       max(1e-300, np.max(np.abs(rr["x"]))))}
        except Exception as ex:  # reported, never fatal for the metric
            ip_solve = {"error": str(ex)}
            if world > 1:
                raise

    tl = [ms_step] + [e2e.get(k, 0.0) for k in ("pinned", "pageable")]
    tmax = torch.tensor(tl, dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms_step = float(tmax[0])
    e2e_max = {"pinned": float(tmax[1]), "pageable": float(tmax[2])}
    value = Kg / (ms_step * 1e-3)
    n_in = [lp.m * batch, lp.m * batch, N, me, m, m]
    bytes_in = 8 * (n_in[0] + n_in[1]) + 2 * 8 * sum(n_in[2:])
    bytes_out = 2 * 8 * (N + me + 2 * m)
    nseg = eng.nseg
    eng.close()
    if rank != 0:
        return None

    mc = lp.m / max(lp.K, 1)
    bf, bs = algorithmic_bytes(nx, nu, mc)
    ff, fs = algorithmic_flops(nx, nu)
    peak, how = measured_peaks()
    Kloc = lp.K * batch          # stages this rank's kernels process per launch
    per = {k.strip("()"): v["ms"] / nprof for k, v in prof.items()}
    solve_names = ("solve_", "range_scan_vec", "range_export_vec")
    t_factor = sum(v for k, v in per.items() if not k.startswith(solve_names))
    t_solve = sum(v for k, v in per.items() if k.startswith(solve_names)) / 2.0
    dom = max(per, key=per.get)
    kname = [k for k in per if k.startswith("seg_riccati_kernel")][0]
    kms = per[kname]
    k1n = [k for k in per if k.startswith("seg_element_kernel")]
    bf_read = 8 * ((nx + nu) * (nx + nu + 1) // 2 + nx * (nx + nu) + 2 * mc)
    # headline fraction: the algorithmic bytes of the WHOLE unit over the unit's time
    unit_bytes = Kg * (bf + 2 * bs)
    ach_unit = unit_bytes / (ms_step * 1e-3) / 1e9 / world      # per GPU
    k3_ach = Kloc * bf / (kms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": "whole unit (1 factor + 2 step), per GPU",
                "achieved": ach_unit, "peak": peak, "unit": "GB/s", "frac": ach_unit / peak,
                "traffic": NCU_TRAFFIC.get(key, (None, None))[0],
                "traffic_source": NCU_TRAFFIC.get(key, (None, None))[1],
                "traffic_kernel": "seg_riccati_kernel (K3), one launch",
                "algorithmic_bytes_per_unit": unit_bytes, "peak_source": how,
                "algorithmic_bytes_per_stage": {"factor": bf, "step": bs},
                "algorithmic_flops_per_stage": {"factor": ff, "step": fs},
                "fp64": {"factor_tflops": Kloc * ff / (t_factor * 1e-3) / 1e12 if t_factor else None,
                         "peak_tflops": FP64_TFLOPS, "peak_source": "scripts/mb/mb_dmma.cu (measured)",
                         "arithmetic_intensity_factor": ff / bf},
                # K3 = the kernel that does the algorithmic factor work of every stage
                "k3": {"kernel": kname, "ms": kms, "achieved": k3_ach, "frac": k3_ach / peak,
                       "algorithmic_bytes_per_launch": Kloc * bf},
                # K1 (seg_element) is the parallel-in-time condensation: it reads the same
                # Q, fx, fu, z/w bytes as K3 and writes only P boundary elements; overhead
                # of the algorithm, reported next to K3
                "k1_condensation": ({"kernel": k1n[0], "ms": per[k1n[0]],
                                     "read_bytes_per_launch": Kloc * bf_read,
                                     "frac_of_hbm": Kloc * bf_read / (per[k1n[0]] * 1e-3) / 1e9 / peak}
                                    if k1n else None),
                "kernel_ms_per_unit": per, "dominant_kernel": dom,
                "kernel_ms_note": "plain launches bracketed by CUDA events (graphs off): shares, "
                                  "not absolutes; their sum exceeds ms_per_step",
                "factor_ms": t_factor, "step_ms": t_solve,
                "factor_frac_of_hbm": Kloc * bf / (t_factor * 1e-3) / 1e9 / peak if t_factor else None,
                "step_frac_of_hbm": Kloc * bs / (t_solve * 1e-3) / 1e9 / peak if t_solve else None}
    if wl.get("bound") == "fp64":
        # AI of the factor (39.5 flop/B) is above the ridge (37.2 TF / 6.5 TB/s = 5.7):
        # the factor's algorithmic flops over the factor's time against the measured
        # FP64 DMMA peak (the steps stay HBM-bound and are reported above)
        ach = Kloc * ff / (t_factor * 1e-3) / 1e12
        roofline.update({"bound": "fp64", "kernel": "factor (K1 + tree + K3), plain-launch event times",
                         "achieved": ach, "peak": FP64_TFLOPS, "unit": "TFLOP/s",
                         "frac": ach / FP64_TFLOPS, "peak_source": "scripts/mb/mb_dmma.cu (measured)",
                         "hbm_unit_frac": ach_unit / peak})
    cfg = config_of(wl, world)
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world,
            "steps": steps, "warmup": warmup, "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": wl["scaling"], "vs_baseline": None,
            "dtype": "f64", "data": ("synthetic (drawn on the device, torch Philox, seed 1234 + rank)"
                                     if wl.get("device_generated") else "synthetic"),
            "config": cfg,
            "details": {"segments_per_instance": nseg, "parallelism": parallelism,
                        "l2": "flushed (256 MiB write) between timed iterations",
                        "ms_per_step_spread_rank0": ms_spread,
                        "kkt_residual_after_refined_solve": res, "kkt_steps_in_that_solve": nsolve},
            "roofline": roofline,
            "e2e": {"value": Kg / e2e_max["pinned"], "unit": UNIT,
                    "h2d_bytes_per_step": bytes_in * world, "d2h_bytes_per_step": bytes_out * world,
                    "ms_per_step": 1e3 * e2e_max["pinned"],
                    "pinned": Kg / e2e_max["pinned"],
                    "pageable": (Kg / e2e_max["pageable"]) if e2e_max["pageable"] else None,
                    "note": "hqpcu_factor + 2 x hqpcu_step with host buffers; `value` = pinned (the "
                            "contract), `pageable` = malloc'ed buffers as Hqp_IpCuda::step passes them"},
            "gpu_launches": int(gpu_launches), "clocks": clocks}
    if update_ms is not None:
        line["details"]["update_ms_once_per_sqp_iteration"] = update_ms
        if update_sparse_ms is not None:
            line["details"]["update_ms_from_sparse_values"] = update_sparse_ms
    if wl.get("device_generated"):
        line["details"]["setup_ms_generate_and_upload"] = setup_ms
    if ip_solve is not None:
        line["ip_solve"] = ip_solve
    return line


def run_hl_bfgs(local):
    """Row f2 (next row of SURVEY section 8): the block-diagonal BFGS update of config 2's
    Hessian -- 10^4 blocks of nx+nu = 30 and one of nx = 20 -- with Powell's damping and
    eigenvalue control (Hqp_HL_BFGS::update, hqp/Hqp_HL_BFGS.C:149-243).  Device-resident
    time (CUDA events, L2 flushed), end to end with host buffers through hqphl_bfgs_update,
    and the reference's own update_b_Q on a sample of blocks on one host core."""
    import torch
    from hqp_b200 import hlcuda
    K, n, nl = 10000, 30, 20
    rng = np.random.default_rng(1234)
    M = rng.uniform(-1, 1, (K + 1, n, n))
    Qb = np.einsum("kij,kil->kjl", M, M) / n + 0.05 * np.eye(n)
    bs = np.array([n] * K + [nl], dtype=np.int32)
    Q = np.concatenate([Qb[:K].ravel(), Qb[K, :nl, :nl].ravel()])
    nv = K * n + nl
    s, u = rng.uniform(-1, 1, nv), rng.uniform(-1, 1, nv)
    dev = torch.device("cuda", local)
    qoff = np.concatenate([[0], np.cumsum(bs.astype(np.int64) ** 2)[:-1]])
    voff = np.concatenate([[0], np.cumsum(bs)[:-1]]).astype(np.int32)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    d_bs, d_qo, d_vo, d_Q0, d_s, d_u = t(bs), t(qoff), t(voff), t(Q), t(s), t(u)
    d_info = torch.zeros(3, dtype=torch.int32, device=dev)
    flush = torch.empty(64 << 20, dtype=torch.float32, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    ms = []
    for it in range(8):
        d_Q = d_Q0.clone()
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        hlcuda.bfgs_update_dev(d_bs, d_qo, d_vo, d_Q, d_s, d_u, d_info, n, 1.0, 0.1, 1e-8, True, stream)
        e1.record()
        torch.cuda.synchronize()
        if it >= 3:
            ms.append(e0.elapsed_time(e1))
    e2e_ms = float("inf")
    for _ in range(3):  # (best of three: the call allocates and frees its device buffers)
        t0 = time.perf_counter()
        got, info = hlcuda.bfgs_update(bs, Q, s, u, 1.0, 0.1, 1e-8, True, device=local)
        e2e_ms = min(e2e_ms, (time.perf_counter() - t0) * 1e3)
    out = {"workload": "Hessian of config 2: 10000 blocks of 30 + one of 20, eigenvalue control on",
           "blocks": int(bs.size), "ms_device": float(np.mean(ms)),
           "blocks_per_s_device": float(bs.size / (np.mean(ms) * 1e-3)),
           "ms_e2e_host_buffers": e2e_ms, "blocks_shifted": info[0], "sweep_limit_hits": info[2],
           "bytes_per_call": int(2 * Q.size * 8),
           "frac_of_hbm": float(2 * Q.size * 8 / (np.mean(ms) * 1e-3) / 1e9 / measured_peaks()[0])}
    # correctness of what was timed: sampled blocks against the restatement
    from oracle import hl_bfgs_oracle
    worst = 0.0
    for k in (0, 4321, K - 1):
        want, _, _ = hl_bfgs_oracle.update_block(Qb[k], s[k * n:(k + 1) * n], u[k * n:(k + 1) * n], 1.0, 0.1, 1e-8, True)
        g = got[k * n * n:(k + 1) * n * n].reshape(n, n)
        worst = max(worst, float(np.max(np.abs(np.triu(g - want))) / np.max(np.abs(want))))
    if not worst < 1e-10:
        raise RuntimeError(f"hl_bfgs: sampled blocks differ from the restatement ({worst})")
    out["max_rel_diff_vs_oracle_sampled"] = worst
    try:
        from oracle import refharness
        if refharness.available():
            nb = 300
            t0 = time.perf_counter()
            for k in range(nb):
                refharness.hl_bfgs_block(Qb[k], s[k * n:(k + 1) * n], u[k * n:(k + 1) * n], 1.0, 0.1, 1e-8, True)
            dt = time.perf_counter() - t0
            out["cpu_baseline"] = {"value": nb / dt, "unit": "blocks/s", "cores": 1, "kind": "reference",
                                   "sample": f"{nb} of {bs.size} blocks through Hqp_HL_BFGS::update_b_Q (ctypes call overhead included)"}
    except Exception as ex:
        out["cpu_baseline"] = {"error": str(ex)}
    return out


def run_docp_update(local):
    """Row f4: the stage loop of Hqp_Docp::update (values + derivatives of every stage, bounds)
    for the synthetic nonlinear model at config 2's shape (K = 10^4, nx 20, nu 10, one path
    constraint) and at a slice of config 5's (K = 10^5, nx 40, nu 10).  Device-resident time per
    call (CUDA events, L2 flushed), end to end with host buffers through hqpdocp_update, and the
    unmodified Hqp_Docp::update on one host core for a truncated horizon."""
    import torch
    from hqp_b200 import docpcuda as dc
    from oracle import docp_oracle
    dev = torch.device("cuda", local)
    flush = torch.empty(64 << 20, dtype=torch.float32, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    peak = measured_peaks()[0]
    out = {}
    for key, (K, nx, nu) in (("c2_shape", (10000, 20, 10)), ("c5_slice", (100000, 40, 10))):
        p = dc.synthnl_problem(K, nx, nu, 1, 1, seed=3)
        e = dc.DocpCuda(p, device=local)
        e.set_stream(stream)
        t = lambda n: torch.empty(max(1, n), dtype=torch.float64, device=dev)
        xd = torch.from_numpy(p.x_init).to(dev)
        fo, b, d, g = t(1), t(p.me), t(p.m), t(p.N)
        fx, fu, cx, cu = t(K * nx * nx), t(K * nx * nu), t(p.ncns * nx), t(K * p.nc * nu)
        res = {"workload": f"synthetic nonlinear DOCP, K={K} nx={nx} nu={nu} nc=1: N={p.N}, me={p.me}, m={p.m}"}
        # algorithmic bytes of one update: every output written once, x and the stage
        # parameters read once (the model matrices A, B are shared by all stages)
        nd = nx + nu
        bytes_upd = 8 * (K * (nx * nx + nx * nu + p.nc * nd) + 2 * p.N + p.me + p.m + (K + 1) * nx)
        bytes_fbd = 8 * (p.N + p.me + p.m + (K + 1) * nx)
        for name, call, nbytes in (
                ("update_ad", lambda: e.update_dev(xd, fo, b, d, g, fx, fu, cx, cu, dc.GRAD_AD), bytes_upd),
                ("update_fd", lambda: e.update_dev(xd, fo, b, d, g, fx, fu, cx, cu, dc.GRAD_FD), bytes_upd),
                ("update_fbd", lambda: e.update_fbd_dev(xd, fo, b, d), bytes_fbd)):
            ms = []
            for it in range(8):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                call()
                e1.record()
                torch.cuda.synchronize()
                if it >= 3:
                    ms.append(e0.elapsed_time(e1))
            m = float(np.mean(ms))
            res[name] = {"ms_device": m, "stages_per_s": (K + 1) / (m * 1e-3), "algorithmic_bytes": int(nbytes),
                         "frac_of_hbm": float(nbytes / (m * 1e-3) / 1e9 / peak)}
        if key == "c2_shape":
            e2e = float("inf")
            for _ in range(3):
                t0 = time.perf_counter()
                got = e.update(p.x_init, dc.GRAD_FD)
                e2e = min(e2e, (time.perf_counter() - t0) * 1e3)
            res["update_fd"]["ms_e2e_host_buffers"] = e2e
            # correctness of what was timed: sampled stages against the restatement
            worst = 0.0
            for k in (0, 4321, K - 1, K):
                x, u = docp_oracle._stage(p, p.x_init, k)
                jfx, jfu, f0x, f0u, jcx, jcu = docp_oracle.grds_fd(p, k, x, u)
                if k < K:
                    worst = max(worst, float(np.max(np.abs(got["fx"][k] - jfx))), float(np.max(np.abs(got["fu"][k] - jfu))))
                worst = max(worst, float(np.max(np.abs(got["g"][k * nd:k * nd + nx] - f0x))))
            if not worst < 1e-9:
                raise RuntimeError(f"docp_update: sampled stages differ from the restatement ({worst})")
            res["max_abs_diff_vs_oracle_sampled"] = worst
            try:
                from oracle import refharness
                if refharness.available():
                    Ks = 1000
                    ps = dc.synthnl_problem(Ks, nx, nu, 1, 1, seed=3)
                    r = refharness.RefDocp(ps)
                    r.update(ps.x_init, matrices=False)
                    t0 = time.perf_counter()
                    for _ in range(3):
                        r.update(ps.x_init, matrices=False)
                    dt = (time.perf_counter() - t0) / 3
                    t0 = time.perf_counter()
                    for _ in range(3):
                        r.update(ps.x_init, fbd_only=True)
                    dtv = (time.perf_counter() - t0) / 3
                    r.close()
                    res["cpu_baseline"] = {"value": (Ks + 1) / dt, "unit": "stages/s (update)", "cores": 1,
                                           "kind": "reference", "update_fbd_stages_per_s": (Ks + 1) / dtv,
                                           "sample": f"K={Ks} of {K}: unmodified Hqp_Docp::update (default update_grds) "
                                                     "with the model restated as an Hqp_Docp subclass"}
            except Exception as ex:
                res["cpu_baseline"] = {"error": str(ex)}
        res["bound"] = ("instruction issue, not HBM: nx+nu+1 model evaluations per stage (ncu of the FD kernel: "
                        "issue slots 75 % busy, FP64 pipe 25 %, DRAM < 1 %; profiles/r02_ncu_full_docp.md)")
        n0 = e.launches
        e.update_fbd_dev(xd, fo, b, d)
        res["gpu_launches_per_call"] = e.launches - n0
        torch.cuda.synchronize()
        e.close()
        del fx, fu, cx, cu, xd, b, d, g
        out[key] = res
    return out


def run_sqp_stack(local):
    """SQP-driven workload, end to end: hqp_solve of the synthetic nonlinear DOCP (K = 300, nx 20,
    nu 10) by the UNMODIFIED reference SQP solver (oracle/_ref is the host program here, as HQP
    would be), once with the reference's modules only (cpu_baseline) and once with every device
    module of this repository plugged in: program on Hqp_DocpCuda<> (row f4), sqp_hela CudaBFGS
    (row f2), sqp_qp_solver CudaMehrotra (device-resident IP solver on the KKT engine).  Wall clock
    of setup + solve; the host still assembles SPMATs between the device calls."""
    import subprocess
    K, nx, nu = 300, 20, 10
    runner = os.path.join(ROOT, "tests", "sqp_stack_runner.py")
    env = dict(os.environ, STACK_SIM="0", CUDA_VISIBLE_DEVICES=os.environ.get("CUDA_VISIBLE_DEVICES", str(local)))
    res = {}
    for which in ("ref", "cuda"):
        out = subprocess.run([sys.executable, runner, which, str(K), str(nx), str(nu), "0"], capture_output=True,
                             text=True, timeout=600, env=env)
        lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
        if not lines:
            raise RuntimeError(f"sqp_stack {which}: rc={out.returncode} {out.stderr[-400:]}")
        d = json.loads(lines[-1])
        d.pop("x", None)
        res[which] = d
    r, c = res["ref"], res["cuda"]
    return {"workload": f"hqp_solve, synthetic nonlinear DOCP K={K} nx={nx} nu={nu}, Powell SQP + BFGS + Mehrotra",
            "device_modules": {k: c[k] for k in ("result", "objective", "sqp_iters", "qp_iters", "seconds_setup_and_solve")},
            "cpu_baseline": {"value": r["seconds_setup_and_solve"], "unit": "s per solve", "cores": 1, "kind": "reference",
                             "sample": "the same solve with Prg_SynthNL / Hqp_HL_BFGS / Hqp_IpsMehrotra + Hqp_IpLQDOCP",
                             "result": r["result"], "objective": r["objective"], "sqp_iters": r["sqp_iters"],
                             "qp_iters": r["qp_iters"]},
            "same_iteration_counts": (r["sqp_iters"], r["qp_iters"]) == (c["sqp_iters"], c["qp_iters"]),
            "objective_rel_diff": abs(c["objective"] - r["objective"]) / abs(r["objective"])}


def run_sqp_ops(local):
    """Row f3: grd_L and the merit functions on config 2's QP (N = 300,020) through the
    host-pointer calls (H2D of the vectors inside the timed region), next to the reference's
    own Meschach code on one host core."""
    from hqp_b200.ipcuda import IpCuda
    from hqp_b200.problem import synth_lqdocp
    p = synth_lqdocp(20, 10, 10000)
    rng = np.random.default_rng(5)
    s, y, z = rng.uniform(-1, 1, p.N), rng.uniform(-1, 1, p.me), rng.uniform(0, 1, p.m)
    re, r = rng.uniform(0, 2, p.me), rng.uniform(0, 2, p.m)
    e = IpCuda(p, device=local)
    e.update()
    for _ in range(2):
        g = e.sqp_grd_L(p.c, y, z)
        m8 = e.sqp_merit(1.0, p.c, s, p.b, p.d, re, r)
    reps = 5
    t0 = time.perf_counter()
    for _ in range(reps):
        g = e.sqp_grd_L(p.c, y, z)
    t_g = (time.perf_counter() - t0) / reps * 1e3
    t0 = time.perf_counter()
    for _ in range(reps):
        m8 = e.sqp_merit(1.0, p.c, s, p.b, p.d, re, r)
    t_m = (time.perf_counter() - t0) / reps * 1e3
    e.close()
    out = {"workload": "config 2's QP: N=300020, me=200020, m=200000", "ms_grd_L_e2e": t_g,
           "ms_merit_e2e": t_m, "note": "host buffers in, results out; merit = phi, phi1, s'Qs, c's, norms in one call"}
    from oracle import sqp_oracle
    want = sqp_oracle.merit(p, 1.0, p.c, s, p.b, p.d, re, r)
    err = float(np.max(np.abs(m8[:4] - want[:4]) / np.maximum(1.0, np.abs(want[:4]))))
    errg = float(np.max(np.abs(g - sqp_oracle.grd_L(p, p.c, y, z))))
    if not (err < 1e-10 and errg < 1e-10):
        raise RuntimeError(f"sqp_ops: results differ from the restatement ({err}, {errg})")
    out["max_rel_diff_vs_oracle"] = max(err, errg)
    try:
        from oracle import refharness
        if refharness.available():
            qp = refharness.RefQP(p)
            refharness.sqp_eval(qp, 1.0, s, y, z, re, r)
            t0 = time.perf_counter()
            for _ in range(3):
                refharness.sqp_eval(qp, 1.0, s, y, z, re, r)
            out["cpu_baseline"] = {"value": (time.perf_counter() - t0) / 3 * 1e3, "unit": "ms per grd_L + phi + phi1 + s'Qs",
                                   "cores": 1, "kind": "reference",
                                   "sample": "Hqp_SqpSolver::grd_L, Hqp_SqpPowell::phi/phi1, sp_mv_symmlt on the same QP"}
            qp.close()
    except Exception as ex:
        out["cpu_baseline"] = {"error": str(ex)}
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist
    rank, world, local = dist_env()
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: there is no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    key = args.workload
    line = run_workload(args, key, args.steps, args.warmup, group)
    wl = WORKLOADS[key]
    if rank == 0:
        # reference CPU path on the host cores: rank 0, N = 1 only
        if world == 1:
            cap = (100 if wl["nx"] > 64 else 10000) if wl["batch"] == 1 else wl["K"]
            cb = time_reference(wl, 3, 1, max_stages=cap)
            line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
            if wl["batch"] > 1:
                try:
                    line["cpu_baseline_all_cores"] = time_reference_batched(wl)
                except Exception as ex:
                    line["cpu_baseline_all_cores"] = {"error": str(ex)}
        else:
            line["cpu_baseline"] = None
    # the other BASELINE configurations ride along as sub-objects of the default line
    if key == "c2" and not args.no_extra:
        extra = {}
        try:
            sub = run_workload(args, "c5", max(3, min(args.steps, 5)), 3, group)
            if rank == 0:
                if world == 1:
                    cb = time_reference(WORKLOADS["c5"], 2, 1, max_stages=5000)
                    sub["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
                extra["c5"] = sub
        except Exception as ex:
            if world > 1:
                raise
            extra["c5"] = {"error": str(ex)}
        if world == 1:
            try:
                sub = run_workload(args, "c4", 3, 3, group, extras=False)
                cb = time_reference(WORKLOADS["c4"], 2, 1, max_stages=100)
                sub["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
                extra["c4"] = sub
            except Exception as ex:
                extra["c4"] = {"error": str(ex)}
            try:
                sub = run_workload(args, "c3", max(3, min(args.steps, 10)), 3, group, extras=False)
                cb = time_reference(WORKLOADS["c3"], 3, 1, max_stages=WORKLOADS["c3"]["K"])
                sub["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
                try:
                    sub["cpu_baseline_all_cores"] = time_reference_batched(WORKLOADS["c3"])
                except Exception as ex:
                    sub["cpu_baseline_all_cores"] = {"error": str(ex)}
                extra["c3"] = sub
            except Exception as ex:
                extra["c3"] = {"error": str(ex)}
            try:
                extra["hl_bfgs"] = run_hl_bfgs(local)
            except Exception as ex:
                extra["hl_bfgs"] = {"error": str(ex)}
            try:
                extra["sqp_ops"] = run_sqp_ops(local)
            except Exception as ex:
                extra["sqp_ops"] = {"error": str(ex)}
            try:
                extra["docp_update"] = run_docp_update(local)
            except Exception as ex:
                extra["docp_update"] = {"error": str(ex)}
            try:
                extra["sqp_stack"] = run_sqp_stack(local)
            except Exception as ex:
                extra["sqp_stack"] = {"error": str(ex)}
        if rank == 0:
            line.update(extra)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--nseg", type=int, default=0)
    ap.add_argument("--no-extra", action="store_true",
                    help="default workload only: skip the c5 / c3 sub-objects")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference(args, WORKLOADS[args.workload])
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
