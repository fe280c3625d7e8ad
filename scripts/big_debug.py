import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from common import make_problem, relerr
from hqp_b200.ipcuda import IpCuda
from hqp_b200.problem import rhs_for
from oracle.portoracle import PortOracle
nx, nu, K, nseg = [int(a) for a in sys.argv[1:5]]
p = make_problem(nx, nu, K, 1, 0, 1)
z, w, r1, r2, r3, r4 = rhs_for(p, seed=41)
o = PortOracle(p); o.factor(z, w); ref = o.step(r1, r2, r3, r4)
e = IpCuda(p, nseg=nseg); e.update(); e.factor(z, w)
print("status", e.sync_status(), "nseg", e.nseg)
V, R = e.get_factor()
print("V err", relerr(V[0], o.Vxx()), "R err", relerr(R[0], o.Rux()))
for k in range(K + 1):
    print("  V[%d] err %.2e" % (k, relerr(V[0][k], o.Vxx()[k])), end="")
print()
mine = e.step(r1, r2, r3, r4)
for a, b, n in zip(mine, ref, "dx dy dz dw".split()):
    print(n, relerr(a, b), np.abs(a).max(), np.abs(b).max())
