import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from hqp_b200 import ipcuda
lib = sys.argv[1]
ipcuda.LIB_PATH = os.path.join(os.path.dirname(ipcuda.LIB_PATH), lib)
from hqp_b200.problem import synth_lqdocp, synth_rhs
p = synth_lqdocp(20, 10, 10000)
z, w, r1, r2, r3, r4 = synth_rhs(p)
e = ipcuda.IpCuda(p); e.update()
L = ipcuda.lib()
for _ in range(3): L.hqpcu_factor(e.h, ipcuda._hp(z), ipcuda._hp(w))
e.profile(True)
for _ in range(10): L.hqpcu_factor(e.h, ipcuda._hp(z), ipcuda._hp(w))
pr = e.profile_read()
print(lib, {k: round(v["ms"] / 10, 4) for k, v in pr.items()})
