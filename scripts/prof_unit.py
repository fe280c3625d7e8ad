"""Scratch: run a few units of a workload with plain launches (for ncu).
   python scripts/prof_unit.py <c2|c3|c4|c5s|ips|franke> [units]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from hqp_b200.ipcuda import IpCuda
from hqp_b200.problem import synth_lqdocp, synth_rhs
wl = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 3
nx, nu, K, B = {"c2": (20, 10, 10000, 1), "c3": (12, 4, 50, 4096), "c4": (200, 50, 296, 1), "c5s": (40, 10, 100000, 1),
                "ips": (20, 10, 10000, 1), "franke": (20, 10, 10000, 1)}[wl]
p = synth_lqdocp(nx, nu, K); z, w, r1, r2, r3, r4 = synth_rhs(p)
e = IpCuda(p, batch=B)
if B > 1:
    bc = lambda a: np.broadcast_to(a, (B,) + a.shape)
    e.update(Q=bc(p.Q), fx=bc(p.fx), fu=bc(p.fu), ineq_val=bc(p.ineq_val))
else:
    e.update()
if wl == "ips":
    print(e.mehrotra_solve()["iters"]); sys.exit(0)
if wl == "franke":
    print(e.franke_solve()["iters"]); sys.exit(0)
e.set_stream(torch.cuda.current_stream().cuda_stream)
dv = [torch.from_numpy(np.tile(v, B)).cuda() for v in (z, w, r1, r2, r3, r4)]
o = [torch.zeros_like(v) for v in dv[2:]]
for _ in range(n):
    e.factor_dev(dv[0].data_ptr(), dv[1].data_ptr())
    for _ in range(2): e.step_dev(*[t.data_ptr() for t in dv[2:]], *[t.data_ptr() for t in o])
torch.cuda.synchronize()
print("done", e.nseg)
