import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from hqp_b200 import ipcuda
ipcuda.LIB_PATH = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scratch_bin", "libhqpcuda_timing.so")
from hqp_b200.problem import synth_lqdocp, synth_rhs
K = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
p = synth_lqdocp(200, 50, K)
z, w, r1, r2, r3, r4 = synth_rhs(p)
e = ipcuda.IpCuda(p); e.update()
for _ in range(2): e.factor(z, w)
out = (ctypes.c_longlong * 32)()
ipcuda.lib().hqpcu_debug_stamps(e.h, out)
st = list(out)[16:]
names = ["acquire", "T=V F (+W)", "G+=F'T", "LDL(+solve K3)", "Rux solve", "V,Phi", "sym/store"]
print("nseg", e.nseg, "K3 last stage of CTA 7 (cycles):", {n: st[i+1]-st[i] for i, n in enumerate(names)}, "stage total", st[7]-st[0])
e.close()
