"""Scratch: where does a device-resident Mehrotra solve spend its time (C2)?"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from hqp_b200.ipcuda import IpCuda
from hqp_b200.problem import synth_lqdocp
p = synth_lqdocp(20, 10, 10000)
e = IpCuda(p); e.update()
e.mehrotra_solve()
t0 = time.perf_counter(); r = e.mehrotra_solve(); dt = time.perf_counter() - t0
print("plain:", r["iters"], "iterations", 1e3 * dt, "ms")
e.profile(True)
t0 = time.perf_counter(); r = e.mehrotra_solve(); dt2 = time.perf_counter() - t0
prof = e.profile_read(); e.profile(False)
tot = sum(v["ms"] for v in prof.values())
print("profiled wall", 1e3 * dt2, "ms; kernel sum", tot, "ms")
for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"]):
    print(f"  {k[:60]:<60} n={v['n']:>4} {v['ms']:8.3f} ms")
