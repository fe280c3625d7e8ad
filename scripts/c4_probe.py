"""Scratch: time factor / step of the large-block workload (nx=200 nu=50).
   python scripts/c4_probe.py [K]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from hqp_b200.ipcuda import IpCuda
from hqp_b200.problem import synth_lqdocp, synth_rhs
K = int(sys.argv[1]) if len(sys.argv) > 1 else 296
p = synth_lqdocp(200, 50, K)
z, w, r1, r2, r3, r4 = synth_rhs(p)
e = IpCuda(p)
e.update()
e.set_stream(torch.cuda.current_stream().cuda_stream)
dv = [torch.from_numpy(v).cuda() for v in (z, w, r1, r2, r3, r4)]
o = [torch.zeros_like(v) for v in dv[2:]]
def run(fn, n=3):
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))
f = run(lambda: e.factor_dev(dv[0].data_ptr(), dv[1].data_ptr()))
s = run(lambda: e.step_dev(*[t.data_ptr() for t in dv[2:]], *[t.data_ptr() for t in o]))
res, n = e.solve_dev(*[t.data_ptr() for t in dv[2:]], *[t.data_ptr() for t in o])
print(f"c4 K={K} nseg={e.nseg}: factor {f:.2f} ms step {s:.2f} ms unit {f+2*s:.2f} ms -> {K/((f+2*s)*1e-3):.0f} stages/s; "
      f"factor {K*45.54e6/(f*1e-3)/1e12:.3f} TFLOP/s; residual {res:.2e} ({n} steps)", flush=True)
e.profile(True)
e.factor_dev(dv[0].data_ptr(), dv[1].data_ptr())
e.step_dev(*[t.data_ptr() for t in dv[2:]], *[t.data_ptr() for t in o])
for k, v in sorted(e.profile_read().items(), key=lambda kv: -kv[1]["ms"]):
    print(f"   {k:60s} {v['ms']:9.3f} ms  x{v['n']}")
e.close()
