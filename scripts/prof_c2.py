"""Scratch: per-kernel CUDA-event times (plain launches) of one C2 unit."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from hqp_b200.ipcuda import IpCuda
from hqp_b200.problem import synth_lqdocp, synth_rhs
wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
nx, nu, K = {"c2": (20, 10, 10000), "c5s": (40, 10, 100000)}[wl]
p = synth_lqdocp(nx, nu, K); z, w, r1, r2, r3, r4 = synth_rhs(p)
e = IpCuda(p); e.update(); e.set_stream(torch.cuda.current_stream().cuda_stream)
dv = [torch.from_numpy(v).cuda() for v in (z, w, r1, r2, r3, r4)]
o = [torch.zeros_like(v) for v in dv[2:]]
def unit():
    e.factor_dev(dv[0].data_ptr(), dv[1].data_ptr())
    for _ in range(2): e.step_dev(*[t.data_ptr() for t in dv[2:]], *[t.data_ptr() for t in o])
for _ in range(3): unit()
e.profile(True)
n = 10
for _ in range(n): unit()
pr = e.profile_read()
tot = 0
for k, v in sorted(pr.items(), key=lambda kv: -kv[1]["ms"]):
    print(f"  {k:58s} {1e3*v['ms']/n:8.1f} us/unit  x{v['n']//n:3d}  {1e3*v['ms']/v['n']:7.2f} us each")
    tot += v["ms"] / n
print("sum", tot * 1e3, "us", "HS", os.environ.get("HQPCU_HS", "1"))
