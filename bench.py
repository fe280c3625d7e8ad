#!/usr/bin/env python
"""KKT factor+solve throughput of the Hqp_IpCuda engine (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                  [--workload c2|c3|c5s]

One "step" = one unit of interior-point KKT work on one synthetic LQ-DOCP
horizon: 1 factor + 2 step (Mehrotra predictor + corrector).
`value` = stages/s = K_stages * batch * n_gpus / step time, device-timed with
inputs resident in HBM; `e2e` = the same through the host-pointer C-ABI calls
(pinned host buffers, H2D/D2H inside the timed region).

Workloads (BASELINE.json configs):
  c2   nx=20 nu=10 K=10,000, one instance  (the configuration the metric is quoted on)
  c3   4096 instances nx=12 nu=4 K=50, one CTA per instance
  c5s  nx=40 nu=10 K=100,000 (single-GPU slice of the long-horizon config)
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
    "c2": dict(nx=20, nu=10, K=10000, batch=1,
               name="synthetic LQ-DOCP nx=20 nu=10 K=10000 (BASELINE configs[1])"),
    "c3": dict(nx=12, nu=4, K=50, batch=4096,
               name="batched MPC QPs 4096 x (nx=12 nu=4 K=50) (BASELINE configs[2])"),
    "c5s": dict(nx=40, nu=10, K=100000, batch=1,
                name="long-horizon LQ-DOCP nx=40 nu=10 K=100000 (slice of BASELINE configs[4])"),
}
# dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the roofline kernel
# (seg_riccati_kernel) from the committed `ncu --set full` capture; null where no
# capture exists.  Compare with K * algorithmic bytes per stage (C2: 160.4 MB).
NCU_TRAFFIC = {"c2": (124.024832e6 + 50.681344e6, "profiles/r01_ncu_full_k1k3_v7.md (K3 section)")}
METRIC = "LQ-DOCP KKT factor+solve stages/s"
UNIT = "stages/s"


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
    sample = (f"{Ks} of {K * batch} stages (one instance), {steps} x (1 factor + 2 step), "
              f"single thread: the reference path is not thread-safe")
    return dict(value=Ks / t_step, unit=UNIT, cores=1, kind=kind, sample=sample,
                ms_per_unit=1e3 * t_step, stages=Ks)


def run_reference(args, wl):
    rank, world, _ = dist_env()
    if rank != 0:
        return
    cap = 10000 if wl["batch"] == 1 else wl["K"]
    cb = time_reference(wl, args.steps, min(args.warmup, 1), max_stages=cap)
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT,
            "n_gpus": args.gpus, "steps": args.steps, "warmup": min(args.warmup, 1),
            "ms_per_step": cb["ms_per_unit"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl["name"], "unit_of_work": "1 factor + 2 step",
                       "timed_stages": cb["stages"]},
            "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------- ours --
def run_ours(args, wl):
    import torch
    import torch.distributed as dist
    from hqp_b200.ipcuda import IpCuda
    from hqp_b200.problem import synth_lqdocp, synth_rhs

    rank, world, local = dist_env()
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: there is no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    nx, nu, K, batch = wl["nx"], wl["nu"], wl["K"], wl["batch"]

    N1 = 0  # bytes bookkeeping below
    update_sparse_ms = None
    sharded_instances = world > 1 and batch > 1
    if world == 1 or sharded_instances:
        # one horizon (or one batch of instances) on the GPU; with N > 1 GPUs a batch
        # workload shards its independent instances: `batch` per rank, no collective
        # in the KKT path
        p = synth_lqdocp(nx, nu, K, seed=1234)
        z, w, r1, r2, r3, r4 = synth_rhs(p, seed=4321)
        eng = IpCuda(p, batch=batch, device=local, nseg=args.nseg)
        stream = torch.cuda.current_stream()
        eng.set_stream(stream.cuda_stream)

        def rep(a):
            return np.ascontiguousarray(np.tile(a, batch)) if batch > 1 else a

        if batch > 1:
            eng.update(Q=np.broadcast_to(p.Q, (batch,) + p.Q.shape),
                       fx=np.broadcast_to(p.fx, (batch,) + p.fx.shape),
                       fu=np.broadcast_to(p.fu, (batch,) + p.fu.shape),
                       ineq_val=np.broadcast_to(p.ineq_val, (batch,) + p.ineq_val.shape))
            update_ms = None
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
        parallelism = (f"{world} x {batch} independent instances, sharded over the ranks, no collective"
                       if sharded_instances else "single")
    else:
        # horizon split: ONE horizon of world*K stages, contiguous
        # stage ranges per rank, boundary elements exchanged with all-gathers
        from hqp_b200.dist import CudaRangeEngine, RangeSolver, local_vectors, split_problem
        pg = synth_lqdocp(nx, nu, K * world, seed=1234)
        gvec = synth_rhs(pg, seed=4321)
        lp, rm = split_problem(pg, world)[rank]
        host = [np.ascontiguousarray(a) for a in local_vectors(pg, rm, *gvec)]
        del pg, gvec
        stream = torch.cuda.current_stream()
        t0 = time.perf_counter()
        reng = CudaRangeEngine(lp, rm, device=local, nseg=args.nseg)
        update_ms = 1e3 * (time.perf_counter() - t0)
        eng = reng.eng
        solver = RangeSolver(reng, rank, world)
        p = lp
        parallelism = f"horizon split, {world} contiguous stage ranges, 2 all-gathers per factor + 2 per step"
    pinned = [torch.from_numpy(np.array(a)).pin_memory() for a in host]
    dvec = [t.to(dev) for t in pinned]
    N, me, m = lp.N * batch, lp.me * batch, lp.m * batch
    outs = [torch.empty(n, dtype=torch.float64, device=dev) for n in (N, me, max(m, 1), max(m, 1))]
    hout = [torch.empty(n, dtype=torch.float64).pin_memory() for n in (N, me, max(m, 1), max(m, 1))]
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)  # > 126 MB L2

    hp = [t.numpy() for t in pinned]
    ho = [t.numpy() for t in hout]
    from hqp_b200 import ipcuda as _ic
    L = _ic.lib()

    if world == 1 or sharded_instances:
        def unit_dev():
            eng.factor_dev(dvec[0].data_ptr(), dvec[1].data_ptr())
            for _ in range(2):
                eng.step_dev(*[t.data_ptr() for t in dvec[2:]], *[t.data_ptr() for t in outs])

        def unit_host():
            _ic._check(L.hqpcu_factor(eng.h, _ic._hp(hp[0]), _ic._hp(hp[1])), "factor")
            for _ in range(2):
                _ic._check(L.hqpcu_step(eng.h, *[_ic._hp(a) for a in hp[2:]],
                                        *[_ic._hp(a) for a in ho]), "step")
    else:
        def unit_dev():
            solver.factor(dvec[0], dvec[1])
            for _ in range(2):
                solver.step(*dvec[2:])

        def unit_host():
            # host buffers in, host buffers out: pinned H2D of (z,w), then per step
            # H2D of r1..r4 and D2H of dx..dw around the same range calls
            dvec[0].copy_(pinned[0], non_blocking=True)
            dvec[1].copy_(pinned[1], non_blocking=True)
            solver.factor(dvec[0], dvec[1])
            for _ in range(2):
                for dt, ht in zip(dvec[2:], pinned[2:]):
                    dt.copy_(ht, non_blocking=True)
                res = solver.step(*dvec[2:])
                for ht, dt in zip(hout, res):
                    ht[:dt.numel()].copy_(dt, non_blocking=True)
            torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up ---------------------------------------------------------
    for _ in range(args.warmup):
        flush.zero_()
        unit_dev()
    st = eng.sync_status()
    if st != 0:
        raise RuntimeError(f"factor status {st}")

    # ---- timed region: device-resident inputs --------------------------------
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    launches0 = eng.launches
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
          for _ in range(args.steps)]
    for e0, e1 in ev:
        flush.zero_()            # L2 flush between timed iterations (not timed)
        e0.record(stream)
        unit_dev()
        e1.record(stream)
    barrier()
    gpu_launches = eng.launches - launches0
    ms = [e0.elapsed_time(e1) for e0, e1 in ev]
    ms_step = float(np.mean(ms))

    # ---- e2e: host buffers through the plugin-facing C ABI ----------------------
    unit_host()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        unit_host()
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / args.steps
    clocks = sampler.stop()
    barrier()

    # ---- per-kernel CUDA-event times of the same unit (roofline section) ------
    eng.profile(True)
    for _ in range(args.steps):
        flush.zero_()
        unit_dev()
    prof = eng.profile_read()
    eng.profile(False)

    # ---- whole QP solve on the device (Hqp_IpsMehrotra restated, SURVEY 8d:
    # "full-IP-iteration time including vector kernels") -- reported, not the metric
    ip_solve = None
    if world == 1 and batch == 1:
        try:
            eng.mehrotra_solve()                      # warm (graphs captured)
            s0 = eng.solve_stats()
            t0 = time.perf_counter()
            r = eng.mehrotra_solve()
            dt = time.perf_counter() - t0
            s1 = eng.solve_stats()
            ip_solve = {"result": r["result"], "iterations": r["iters"], "ms_total": 1e3 * dt,
                        "ms_per_iteration": 1e3 * dt / max(r["iters"], 1),
                        "stages_per_s": K * r["iters"] / dt, "gap": r["gap"],
                        "mean_kkt_steps_per_refined_solve": (s1[1] - s0[1]) / max(s1[0] - s0[0], 1),
                        "note": "hqpcu_mehrotra_solve: cold start + IP iterations, host c/b/d in, "
                                "x/y/z/w out; each iteration = 1 factor + 2 refined solves + vector kernels"}
            # the same QP through the unmodified reference on the host
            # (Hqp_IpsMehrotra + Hqp_IpLQDOCP, oracle/_ref), bounded to <= 10^4 stages
            if rank == 0 and K <= 10000:
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

    tmax = torch.tensor([ms_step, e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms_step, e2e_s = float(tmax[0]), float(tmax[1])
    stages = K * batch * world
    value = stages / (ms_step * 1e-3)
    bytes_in = 8 * sum(int(np.prod(a.shape)) for a in host[:2]) + \
        2 * 8 * sum(int(np.prod(a.shape)) for a in host[2:])
    bytes_out = 2 * 8 * (N + me + 2 * m)

    if rank == 0:
        mc = p.m / K
        bf, bs = algorithmic_bytes(nx, nu, mc)
        ff, fs = algorithmic_flops(nx, nu)
        peak, how = measured_peaks()
        per = {k.strip("()"): v["ms"] / args.steps for k, v in prof.items()}
        t_factor = sum(v for k, v in per.items() if not k.startswith("solve_"))
        t_solve = sum(v for k, v in per.items() if k.startswith("solve_")) / 2.0
        dom = max(per, key=per.get)
        # the kernel that performs the algorithmic factor work of every stage
        kname = [k for k in per if k.startswith("seg_riccati_kernel")][0]
        kms = per[kname]
        k1n = [k for k in per if k.strip("()").startswith("seg_element_kernel")]
        bf_read = 8 * ((nx + nu) * (nx + nu + 1) // 2 + nx * (nx + nu) + 2 * mc)
        ach = K * batch * bf / (kms * 1e-3) / 1e9
        roofline = {"bound": "hbm", "kernel": kname, "achieved": ach, "peak": peak,
                    "unit": "GB/s", "frac": ach / peak,
                    "traffic": NCU_TRAFFIC.get(args.workload, (None, None))[0],
                    "traffic_source": NCU_TRAFFIC.get(args.workload, (None, None))[1],
                    "algorithmic_bytes_per_launch": K * batch * bf, "peak_source": how,
                    "algorithmic_bytes_per_stage": {"factor": bf, "step": bs},
                    "algorithmic_flops_per_stage": {"factor": ff, "step": fs},
                    "kernel_ms_per_unit": per, "dominant_kernel": dom,
                    # K1 (seg_element) is the parallel-in-time condensation: it reads the
                    # same Q, fx, fu, z/w bytes as K3 and writes only P boundary elements;
                    # its time is overhead of the algorithm, reported next to K3
                    "k1_condensation": ({"kernel": k1n[0], "ms": per[k1n[0]],
                                         "read_bytes_per_launch": K * batch * bf_read,
                                         "frac_of_hbm": K * batch * bf_read / (per[k1n[0]] * 1e-3) / 1e9 / peak}
                                        if k1n else None),
                    "factor_ms": t_factor, "step_ms": t_solve,
                    "factor_frac_of_hbm": K * batch * bf / (t_factor * 1e-3) / 1e9 / peak,
                    "step_frac_of_hbm": K * batch * bs / (t_solve * 1e-3) / 1e9 / peak}
        cap = 10000 if batch == 1 else K
        # reference CPU path on the host cores: rank 0, N = 1 only
        cb = time_reference(wl, 3, 1, max_stages=cap) if world == 1 else None
        h2d, d2h = bytes_in * world, bytes_out * world  # whole job
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": wl["name"], "unit_of_work": "1 factor + 2 step",
                           "stages_per_gpu": K * batch, "segments_per_instance": eng.nseg,
                           "parallelism": parallelism,
                           "l2": "flushed (256 MiB write) between timed iterations"},
                "roofline": roofline,
                "cpu_baseline": ({k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
                                 if cb else None),
                "e2e": {"value": stages / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d,
                        "d2h_bytes_per_step": d2h, "ms_per_step": 1e3 * e2e_s},
                "gpu_launches": int(gpu_launches), "clocks": clocks}
        if update_ms is not None:
            line["config"]["update_ms_once_per_sqp_iteration"] = update_ms
            if update_sparse_ms is not None:
                line["config"]["update_ms_from_sparse_values"] = update_sparse_ms
        if ip_solve is not None:
            line["ip_solve"] = ip_solve
        print(json.dumps(line), flush=True)
    eng.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--nseg", type=int, default=0)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, wl)
    else:
        run_ours(args, wl)


if __name__ == "__main__":
    main()
