"""Small factor + step cases for compute-sanitizer (memcheck / racecheck)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from hqp_b200.ipcuda import IpCuda
from hqp_b200.problem import rhs_for, synth_lqdocp
from oracle.portoracle import PortOracle
for (nx, nu, K, nseg) in ((20, 10, 64, 4), (12, 4, 40, 1), (40, 10, 32, 2), (5, 3, 40, 3)):
    p = synth_lqdocp(nx, nu, K)
    z, w, r1, r2, r3, r4 = rhs_for(p, seed=1)
    o = PortOracle(p); o.factor(z, w); ref = o.step(r1, r2, r3, r4)
    e = IpCuda(p, nseg=nseg); e.update()
    for _ in range(2):
        e.factor(z, w)
        out = e.step(r1, r2, r3, r4)
    err = max(float(np.max(np.abs(a - b)) / np.max(np.abs(b))) for a, b in zip(out, ref))
    print(nx, nu, K, nseg, "relerr", err, flush=True)
    e.close(); o.close()
