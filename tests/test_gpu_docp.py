"""Full-stack drop-in test: the reference's own host code (Hqp_Docp,
Hqp_SqpPowell, Hqp_IpsMehrotra/Franke, iftcl; all unmodified, oracle/_ref) drives
the Hqp_IpCuda module selected with  qp_mat_solver Cuda  on the hqp_docp example
(hqp_docp/Docp_Main.C), and must reproduce the LQDOCP runs recorded in
tests/golden/docp.json: identical SQP / IP iteration counts, objective within
1e-8 relative (BASELINE.json north star)."""
import json
import os

import pytest

from oracle import refharness

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PLUGIN = os.path.join(ROOT, "hqp_b200", "lib", "libhqp_ipcuda_plugin.so")

needs_ref = pytest.mark.skipif(
    not (refharness.available() and os.path.exists(PLUGIN)),
    reason="oracle/_ref or the Hqp_IpCuda plugin were not built (needs /root/reference at build time)")


def gold():
    with open(os.path.join(ROOT, "tests", "golden", "docp.json")) as f:
        return json.load(f)


@needs_ref
@pytest.mark.parametrize("key,kmax,qps", [
    ("K60_Mehrotra_LQDOCP", 60, "Mehrotra"),
    ("K200_Mehrotra_LQDOCP", 200, "Mehrotra"),
    ("K1000_Mehrotra_LQDOCP", 1000, "Mehrotra"),
])
def test_docp_mehrotra_cuda_matches_lqdocp(key, kmax, qps):
    g = gold()[key]
    r = refharness.docp_did(kmax, qps, "Cuda", plugin=PLUGIN)
    assert r["result"] == "optimal"
    assert r["sqp_iters"] == g["sqp_iters"]
    assert r["qp_iters"] == g["qp_iters"]
    assert abs(r["objective"] - g["objective"]) <= 1e-8 * abs(g["objective"])


@needs_ref
def test_docp_as_shipped_franke_cuda():
    """as shipped: Powell + Franke (hqp_docp/Docp_Main.C:42); Franke's stop test
    consumes the refinement residual, so only the objective and optimality are
    required to match exactly; the IP iteration count may differ by rounding."""
    g = gold()["K60_asShipped_Franke_LQDOCP"]
    r = refharness.docp_did(60, "", "Cuda", plugin=PLUGIN)
    assert r["result"] == "optimal"
    assert r["sqp_iters"] == g["sqp_iters"]
    assert abs(r["objective"] - g["objective"]) <= 1e-8 * abs(g["objective"])
    assert abs(r["qp_iters"] - g["qp_iters"]) <= 3


@needs_ref
@pytest.mark.parametrize("key,kmax", [("K60_Mehrotra_LQDOCP", 60), ("K200_Mehrotra_LQDOCP", 200),
                                      ("K1000_Mehrotra_LQDOCP", 1000)])
def test_docp_device_resident_solver_matches_mehrotra(key, kmax):
    """sqp_qp_solver CudaMehrotra (Hqp_IpsCuda): the whole IP iteration on the
    device -- cold start, hot starts between SQP iterations, cold restarts -- under
    the unmodified SQP driver must reproduce Hqp_IpsMehrotra + Hqp_IpLQDOCP: same
    SQP and (total) IP iteration counts, same objective."""
    g = gold()[key]
    r = refharness.docp_did(kmax, "CudaMehrotra", "", plugin=PLUGIN)
    assert r["result"] == "optimal"
    assert r["sqp_iters"] == g["sqp_iters"]
    assert r["qp_iters"] == g["qp_iters"]
    assert abs(r["objective"] - g["objective"]) <= 1e-8 * abs(g["objective"])


@needs_ref
def test_docp_device_resident_franke_reproduces_as_shipped_run():
    """sqp_qp_solver CudaFranke (hqpcu_franke_solve): Hqp_IpsFranke's iteration on the
    device.  With the as-shipped qp_eps (1e-9, hqp/Hqp_SqpSolver.C:80) the shipped
    docp run is reproduced: same SQP iterations and objective, 56 IP iterations
    against the reference's 57 -- Franke's stop test consumes the residual of the
    refined KKT solve (:372), which depends on the linear solver's rounding: the
    reference itself needs 57 iterations with LQDOCP and 59 with RedSpBKP / SpBKP
    (tests/golden/docp.json).  On the synthetic QPs (no equality rows) the counts
    are identical (tests/test_gpu_ips.py::test_franke_matches_live_reference)."""
    import os
    g = gold()["K60_asShipped_Franke_LQDOCP"]
    env = dict(os.environ, HQP_QP_EPS="1e-9")
    r = refharness.docp_did(60, "CudaFranke", "", plugin=PLUGIN, env=env)
    assert r["result"] == "optimal"
    assert r["sqp_iters"] == g["sqp_iters"]
    assert g["qp_iters"] == 57 and abs(r["qp_iters"] - 57) <= 1
    assert abs(r["objective"] - g["objective"]) <= 1e-8 * abs(g["objective"])


def _vardim_worker(solver, mat, q):
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__))))
    from common import VarDimQP
    from oracle import refharness as R
    if "Cuda" in solver + mat:
        R.load_plugin(PLUGIN)
    p = VarDimQP([3, 3, 5, 2, 4, 4, 6, 3, 5, 4, 4, 2, 3], [2, 1, 3, 2, 1, 4, 2, 3, 1, 2, 2, 3], seed=4)
    r = R.ips_solve(R.RefQP(p), solver, mat, 1e-9)
    q.put((r["iters"], r["result"], r["x"], r["y"], r["z"]))


@needs_ref
def test_non_uniform_stage_dimensions_through_the_plugin():
    """stage-dependent nx_k / nu_k (Hqp_IpLQDOCP::Get_Dim's _nk[k], _mk[k],
    hqp/Hqp_IpLQDOCP.C:201-287): Hqp_IpCuda pads every stage to the largest
    dimensions with decoupled unit blocks.  Under the unmodified Hqp_IpsMehrotra and
    as the device-resident solver modules: the reference's iteration count and
    solution with Hqp_IpLQDOCP."""
    import multiprocessing as mp
    from common import relerr
    ctx = mp.get_context("spawn")
    out = {}
    for solver, mat in (("Mehrotra", "LQDOCP"), ("Mehrotra", "Cuda"), ("CudaMehrotra", ""),
                        ("Franke", "LQDOCP"), ("CudaFranke", "")):
        q = ctx.Queue()
        pr = ctx.Process(target=_vardim_worker, args=(solver, mat, q))
        pr.start()
        out[(solver, mat)] = q.get(timeout=300)
        pr.join()
    ref = out[("Mehrotra", "LQDOCP")]
    for key in (("Mehrotra", "Cuda"), ("CudaMehrotra", "")):
        it, res, x, y, z = out[key]
        assert res == ref[1] == "optimal" and it == ref[0], key
        assert relerr(x, ref[2]) < 1e-8 and relerr(y, ref[3]) < 1e-7 and relerr(z, ref[4]) < 1e-7, key
    reff, fr = out[("Franke", "LQDOCP")], out[("CudaFranke", "")]
    assert fr[1] == reff[1] == "optimal" and abs(fr[0] - reff[0]) <= 1
    assert relerr(fr[2], reff[2]) < 1e-7
