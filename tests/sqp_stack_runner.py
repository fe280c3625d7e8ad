"""Helper of tests/test_docp_update.py (run as a subprocess: the SQP solver keeps global state):
one hqp_solve of the synthetic nonlinear DOCP, either with the reference's own modules
(program Prg_SynthNL, Hqp_HL_BFGS, Hqp_IpsMehrotra + Hqp_IpLQDOCP) or with every device module of
this repository plugged into the unmodified SQP solver (program on Hqp_DocpCuda<>: row f4;
sqp_hela CudaBFGS: row f2; sqp_qp_solver CudaMehrotra: the device-resident IP solver on the KKT
engine).  Prints one JSON line.
    python tests/sqp_stack_runner.py ref|cuda K nx nu [grad]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

if __name__ == "__main__":
    from hqp_b200 import build, docpcuda as dc
    from oracle import refharness as rh
    which, K, nx, nu = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
    grad = int(sys.argv[5]) if len(sys.argv) > 5 else 1
    p = dc.synthnl_problem(K, nx, nu, 1, 0)
    SIM = os.environ.get("STACK_SIM", "1") != "0"  # prg_simulate before sqp_init (Docp_Main.C:69)
    import time
    t0 = time.perf_counter()
    L = rh.lib()
    if which == "cuda":
        rh.load_plugin(os.path.join(build.LIB, "libhqp_ipcuda_plugin.so"))
        prg = rh.RefDocp(p, cuda=True)
        assert L.ref_set_int(b"prg_cuda_grad", grad) == 0
        hela = os.environ.get("STACK_HELA", "CudaBFGS")
        assert L.ref_set_string(b"sqp_hela", hela.encode()) == 0
        out = prg.solve(simulate=SIM, qp_solver=os.environ.get("STACK_QPS", "CudaMehrotra"), mat_solver=os.environ.get("STACK_MAT", ""))
    else:
        prg = rh.RefDocp(p)
        out = prg.solve(simulate=SIM, qp_solver=os.environ.get("STACK_QPS", "Mehrotra"), mat_solver=os.environ.get("STACK_MAT", "LQDOCP"))
    out["seconds_setup_and_solve"] = time.perf_counter() - t0
    out["x"] = [float(v) for v in out["x"]]
    print(json.dumps(out))
