"""Scratch: per-kernel CUDA-event times of one device-resident Mehrotra solve at C2."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hqp_b200.ipcuda import IpCuda
from hqp_b200.problem import synth_lqdocp
p = synth_lqdocp(20, 10, 10000)
e = IpCuda(p); e.update()
r = e.mehrotra_solve()
e.profile(True)
r = e.mehrotra_solve()
pr = e.profile_read()
it = r["iters"]
tot = 0
for k, v in sorted(pr.items(), key=lambda kv: -kv[1]["ms"]):
    print(f"  {k:58s} {1e3*v['ms']/it:8.1f} us/iter  x{v['n']/it:5.1f}  {1e3*v['ms']/v['n']:7.2f} us each")
    tot += v["ms"] / it
print("sum", tot * 1e3, "us per iteration;", it, "iterations")
