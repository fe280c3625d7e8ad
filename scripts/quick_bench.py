"""Scratch: time factor / step of a workload with a given library build and nseg.
   python scripts/quick_bench.py <lib.so|-> <workload> <nseg,nseg,...>"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from hqp_b200 import ipcuda
if sys.argv[1] != "-":
    ipcuda.LIB_PATH = os.path.abspath(sys.argv[1])
import torch
from hqp_b200.problem import synth_lqdocp, synth_rhs
wl = sys.argv[2]
cfg = {"c2": (20, 10, 10000, 1), "c3": (12, 4, 50, 4096), "c5s": (40, 10, 100000, 1)}[wl]
nx, nu, K, batch = cfg
p = synth_lqdocp(nx, nu, K)
z, w, r1, r2, r3, r4 = synth_rhs(p)
for nseg in [int(s) for s in sys.argv[3].split(",")]:
    e = ipcuda.IpCuda(p, nseg=nseg, batch=batch) if batch > 1 else ipcuda.IpCuda(p, nseg=nseg)
    if batch > 1:
        e.update(Q=np.broadcast_to(p.Q, (batch,) + p.Q.shape), fx=np.broadcast_to(p.fx, (batch,) + p.fx.shape),
                 fu=np.broadcast_to(p.fu, (batch,) + p.fu.shape),
                 ineq_val=np.broadcast_to(p.ineq_val, (batch,) + p.ineq_val.shape))
    else:
        e.update()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    def run(fn, n=20):
        ts = []
        for _ in range(n):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        return float(np.median(ts))
    e.set_stream(torch.cuda.current_stream().cuda_stream)
    rep = lambda v: torch.from_numpy(np.tile(v, batch)).cuda()
    dz, dw = rep(z), rep(w)
    d = [rep(v) for v in (r1, r2, r3, r4)]
    o = [torch.zeros_like(v) for v in d]
    f = run(lambda: e.factor_dev(dz.data_ptr(), dw.data_ptr()))
    s = run(lambda: e.step_dev(*[t.data_ptr() for t in d], *[t.data_ptr() for t in o]))
    print(f"{wl} lib={os.path.basename(ipcuda.LIB_PATH)} nseg={e.nseg}: factor {f*1e3:.1f} us step {s*1e3:.1f} us unit {(f+2*s)*1e3:.1f} us -> {K*batch/((f+2*s)*1e-3)/1e6:.2f} M stages/s", flush=True)
    e.close()
