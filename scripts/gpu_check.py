"""Scratch GPU check: parity of the CUDA path vs the oracles + rough timings."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from hqp_b200.problem import synth_lqdocp, synth_rhs
from hqp_b200.ipcuda import IpCuda
from oracle.portoracle import PortOracle

def relerr(a, b):
    if a.size == 0:
        return 0.0
    return float(np.max(np.abs(a - b)) / max(1e-300, np.max(np.abs(b))))

def check(nx, nu, K, nseg, bounds=True, fixed=True):
    p = synth_lqdocp(nx, nu, K, bounds=bounds)
    if not fixed:
        p.fixed_x0 = False
        p.b = p.b[:K * nx].copy()
    z, w, r1, r2, r3, r4 = synth_rhs(p)
    o = PortOracle(p); o.factor(z, w)
    ref = o.step(r1, r2, r3, r4)
    g = IpCuda(p, nseg=nseg); g.update(); g.factor(z, w)
    mine = g.step(r1, r2, r3, r4)
    V, R = g.get_factor()
    errs = [relerr(a, b) for a, b in zip(mine, ref)]
    res = g.residuum(r1, r2, r3, r4, *mine)
    res_o = o.residuum(r1, r2, r3, r4, *mine)
    s = g.solve(r1, r2, r3, r4)
    print(f"nx={nx} nu={nu} K={K} nseg={g.nseg} fixed={fixed}: step relerr {['%.1e'%e for e in errs]} "
          f"V err {relerr(V[0], o.Vxx()):.1e} Rux err {relerr(R[0], o.Rux()):.1e} "
          f"res gpu {res:.2e} (oracle-eval {res_o:.2e}) solve res {s[4]:.2e} steps {s[5]}", flush=True)
    g.close(); o.close()

if __name__ == "__main__":
    check(2, 1, 8, 1)
    check(2, 1, 8, 2)
    check(5, 3, 40, 1)
    check(5, 3, 40, 4)
    check(5, 3, 41, 5, fixed=False)
    check(12, 4, 50, 1)
    check(12, 4, 50, 5)
    check(20, 10, 200, 1)
    check(20, 10, 200, 0)
    check(20, 10, 2000, 0)
    check(7, 9, 64, 6, bounds=False)
    check(40, 10, 300, 0)
    # timing at C2
    import torch
    p = synth_lqdocp(20, 10, 10000)
    z, w, r1, r2, r3, r4 = synth_rhs(p)
    for nseg in (0, 60, 240, 1):
        g = IpCuda(p, nseg=nseg); g.update()
        g.factor(z, w)
        t0 = time.time(); g.factor(z, w); t1 = time.time()
        out = g.step(r1, r2, r3, r4)
        t2 = time.time(); out = g.step(r1, r2, r3, r4); t3 = time.time()
        res = g.residuum(r1, r2, r3, r4, *out)
        print(f"C2 nseg={g.nseg}: factor {1e3*(t1-t0):.2f} ms step {1e3*(t3-t2):.2f} ms (host-pointer API, wall) res {res:.2e}", flush=True)
        g.close()
