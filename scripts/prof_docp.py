"""Scratch: row f4's kernels for ncu / for a quick timing.
  python scripts/prof_docp.py bench      the docp_update sub-object of bench.py, as JSON
  python scripts/prof_docp.py ncu        two updates (AD, FD) + one update_fbd at config 2's shape"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
if sys.argv[1:] == ["bench"]:
    import bench
    print(json.dumps(bench.run_docp_update(0)))
else:
    import torch
    from hqp_b200 import docpcuda as dc
    K, nx, nu = (100000, 40, 10) if "c5" in sys.argv else (10000, 20, 10)
    p = dc.synthnl_problem(K, nx, nu, 1, 1, seed=3)
    e = dc.DocpCuda(p)
    dev = torch.device("cuda:0")
    t = lambda n: torch.empty(max(1, n), dtype=torch.float64, device=dev)
    xd = torch.from_numpy(p.x_init).to(dev)
    fo, b, d, g = t(1), t(p.me), t(p.m), t(p.N)
    fx, fu, cx, cu = t(K * nx * nx), t(K * nx * nu), t(p.ncns * nx), t(K * p.nc * nu)
    e.set_stream(torch.cuda.current_stream().cuda_stream)
    for _ in range(2):
        e.update_dev(xd, fo, b, d, g, fx, fu, cx, cu, dc.GRAD_AD)
        e.update_dev(xd, fo, b, d, g, fx, fu, cx, cu, dc.GRAD_FD)
        e.update_fbd_dev(xd, fo, b, d)
    torch.cuda.synchronize()
    print("done", fo.item())
