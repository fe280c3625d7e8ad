"""GPU tests of the device-resident Mehrotra solver (hqpcu_mehrotra_solve) against
full cold-started Hqp_IpsMehrotra solves of the unmodified reference: identical
IP iteration counts, solutions within 1e-8 (tests/golden/ips_*.npz; also live
against oracle/_ref when it travelled to the box)."""
import os

import numpy as np
import pytest

from common import make_eq_problem, relerr
from hqp_b200.ipcuda import IpCuda
from hqp_b200.problem import synth_lqdocp
from oracle import refharness

pytestmark = pytest.mark.gpu


def solve_gpu(p, nseg=0):
    e = IpCuda(p, nseg=nseg)
    e.update()
    r = e.mehrotra_solve(eps=1e-9)
    e.close()
    return r


@pytest.mark.parametrize("name,dims", [("ips_n5m3K40", (5, 3, 40)), ("ips_n20m10K200", (20, 10, 200))])
@pytest.mark.parametrize("nseg", [1, 0])
def test_mehrotra_matches_reference_golden(name, dims, nseg, golden_dir):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    p = synth_lqdocp(*dims)
    r = solve_gpu(p, nseg)
    assert r["result"] == "optimal" == str(g["result"])
    assert r["iters"] == int(g["iters"])  # IP iteration counts identical
    assert relerr(r["x"], g["x"]) < 1e-8
    assert relerr(r["y"], g["y"]) < 1e-7
    assert relerr(r["z"], g["z"]) < 1e-7
    # final objective within 1e-8 relative
    def obj(x):
        Q, _, _ = p.dense_kkt_blocks() if p.N < 3000 else (None, None, None)
        return None if Q is None else 0.5 * x @ Q @ x + p.c @ x
    if p.N < 3000:
        assert abs(obj(r["x"]) - obj(g["x"])) <= 1e-8 * abs(obj(g["x"]))


def test_mehrotra_with_terminal_equality_and_general_rows(golden_dir):
    g = np.load(os.path.join(golden_dir, "ips_eq_n6m2K30.npz"))
    p = make_eq_problem(6, 2, 30, True, 1, True, [30], 3)
    r = solve_gpu(p)
    assert r["result"] == "optimal" and r["iters"] == int(g["iters"])
    assert relerr(r["x"], g["x"]) < 1e-8


def test_mehrotra_without_inequalities_is_one_newton_step():
    p = synth_lqdocp(6, 3, 25, bounds=False)
    r = solve_gpu(p)
    assert r["result"] == "optimal" and r["iters"] == 1
    Q, A, _ = p.dense_kkt_blocks()
    kkt = np.block([[Q, -A.T], [A, np.zeros((p.me, p.me))]])
    ref = np.linalg.solve(kkt, np.concatenate([-p.c, -p.b]))
    assert relerr(r["x"], ref[:p.N]) < 1e-9


@pytest.mark.skipif(not refharness.available(), reason="compiled reference not present")
@pytest.mark.parametrize("dims", [(12, 4, 50), (20, 10, 1000)])
def test_mehrotra_matches_live_reference(dims):
    p = synth_lqdocp(*dims)
    qp = refharness.RefQP(p)
    ref = refharness.ips_solve(qp, "Mehrotra", "LQDOCP", 1e-9)
    qp.close()
    r = solve_gpu(p)
    assert r["result"] == ref["result"] and r["iters"] == ref["iters"]
    assert relerr(r["x"], ref["x"]) < 1e-8


@pytest.mark.parametrize("name", ["ipshot_n5m3K40", "ipshot_n20m10K100"])
@pytest.mark.parametrize("nseg", [1, 0])
def test_hot_started_sequence_matches_reference_golden(name, nseg, golden_dir):
    """Hqp_IpsMehrotra::hot_start + solve on the device (hqpcu_mehrotra_hot_solve):
    the SQP pattern -- one cold-started QP, then hot-started ones with changed
    linear terms, one of them changed so much that the reference restarts cold
    inside solve() -- must take the reference's iteration counts (its fail_iters
    included) and reach its solutions."""
    import sys
    sys.path.insert(0, golden_dir)
    from make_golden_hot import sequence
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    p = synth_lqdocp(*[int(v) for v in g["cfg"]])
    cs, bs, ds = sequence(p, [float(s) for s in g["scales"]], int(g["seed"]))
    e = IpCuda(p, nseg=nseg)
    e.update()
    prev = None
    for k in range(len(cs)):
        r = e.mehrotra_solve(c=cs[k], b=bs[k], d=ds[k], eps=1e-9, hot=prev)
        assert r["result"] == str(g["result"][k]), k
        assert r["iters"] == int(g["iters"][k]), (k, r["iters"], int(g["iters"][k]))
        assert relerr(r["x"], g["x"][k]) < 1e-7
        assert relerr(r["y"], g["y"][k]) < 1e-6
        prev = (r["x"], r["y"])
    e.close()


@pytest.mark.skipif(not refharness.available(), reason="compiled reference not present")
@pytest.mark.parametrize("dims", [(5, 3, 40), (20, 10, 200), (12, 4, 50)])
@pytest.mark.parametrize("nseg", [1, 0])
def test_franke_matches_live_reference(dims, nseg):
    """hqpcu_franke_solve against a cold-started Hqp_IpsFranke + Hqp_IpLQDOCP solve of
    the unmodified reference: identical iteration count, same solution."""
    p = synth_lqdocp(*dims)
    qp = refharness.RefQP(p)
    ref = refharness.ips_solve(qp, "Franke", "LQDOCP", 1e-9)
    e = IpCuda(p, nseg=nseg)
    e.update()
    r = e.franke_solve(eps=1e-9)
    assert r["result"] == ref["result"] == "optimal"
    assert r["iters"] == ref["iters"]
    assert relerr(r["x"], ref["x"]) < 1e-7
    assert relerr(r["z"], ref["z"]) < 1e-6
    # hot start from the solution: converges again (cold restarts allowed), same x
    r2 = e.franke_solve(eps=1e-9, hot=(r["x"], r["y"], r["z"], r["w"]))
    assert r2["result"] == "optimal" and relerr(r2["x"], ref["x"]) < 1e-6
    e.close(); qp.close()
