"""Cycle stamps of one K3 stage (LQ_TIMING build in scratch_bin/libhqpcuda_timing.so):
   python scripts/stamp_any.py nx nu K"""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from hqp_b200 import ipcuda
ipcuda.LIB_PATH = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scratch_bin", "libhqpcuda_timing.so")
from hqp_b200.problem import synth_lqdocp, synth_rhs
nx, nu, K = [int(a) for a in sys.argv[1:4]]
p = synth_lqdocp(nx, nu, K)
z, w, r1, r2, r3, r4 = synth_rhs(p)
e = ipcuda.IpCuda(p); e.update()
for _ in range(2): e.factor(z, w)
out = (ctypes.c_longlong * 32)()
ipcuda.lib().hqpcu_debug_stamps(e.h, out)
st = list(out)[16:]
print("nseg", e.nseg, "stamps 0..7 diffs:", [st[i+1]-st[i] for i in range(7)], "total", st[7]-st[0])
e.close()
