import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from hqp_b200 import ipcuda
ipcuda.LIB_PATH = os.path.join(os.path.dirname(ipcuda.LIB_PATH), "libhqpcuda_timing.so")
from hqp_b200.problem import synth_lqdocp, synth_rhs
p = synth_lqdocp(20, 10, 10000)
z, w, r1, r2, r3, r4 = synth_rhs(p)
for nseg in (0, 148):
    e = ipcuda.IpCuda(p, nseg=nseg); e.update()
    for _ in range(3): e.factor(z, w)
    out = (ctypes.c_longlong * 32)()
    ipcuda.lib().hqpcu_debug_stamps(e.h, out)
    full = list(out)
    cst = full[:16]
    print("compose (top group):", [cst[i+1]-cst[i] for i in range(6)], "total", cst[6]-cst[0], "[load acc, load child, pre-mm, GJ, post-mm, tail]")
    st = full[16:]
    names = ["acquire", "T=V F (+W)", "G+=F'T", "LDL", "Rux solve", "V,Phi", "sym/store/Psi"]
    print("nseg", e.nseg, "K3 last stage of CTA 7 (cycles):", {n: st[i+1]-st[i] for i, n in enumerate(names)}, "stage total", st[7]-st[0])
    e.close()
