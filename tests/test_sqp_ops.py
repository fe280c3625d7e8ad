"""Row f3: SQP-level vector operations on the device-resident QP
(grd_L, x'Qx / s'Qs, the merit functions phi / phi1, the infeasibility norm).

CPU part: the numpy restatement (oracle/sqp_oracle.py) against the UNMODIFIED
Hqp_SqpSolver::grd_L / ::norm_inf and Hqp_SqpPowell::phi / ::phi1 (oracle/_ref).
GPU part (-m gpu): hqpcu_sqp_* through the C ABI against the restatement, the live
reference, and at config 2's full size."""
import numpy as np
import pytest

from hqp_b200.problem import synth_lqdocp
from oracle import sqp_oracle as O

TOL = 1e-10


def vectors(p, seed):
    rng = np.random.default_rng(seed)
    return dict(s=rng.uniform(-1, 1, p.N), y=rng.uniform(-1, 1, p.me), z=rng.uniform(0, 1, p.m),
                re=rng.uniform(0, 2, p.me), r=rng.uniform(0, 2, p.m), c=rng.uniform(-1, 1, p.N),
                b=rng.uniform(-1, 1, p.me), d=rng.uniform(-0.5, 1, p.m))


def rel(a, b):
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b))) / max(1.0, float(np.max(np.abs(b)))))


def test_oracle_matches_the_live_reference():
    from oracle import refharness as R
    if not R.available():
        pytest.skip("compiled reference (oracle/_ref) not present")
    for nx, nu, K in ((5, 3, 6), (12, 4, 30), (20, 10, 25)):
        p = synth_lqdocp(nx, nu, K)
        v = vectors(p, 3)
        qp = R.RefQP(p)  # (the reference reads c, b, d from the QP itself)
        g, o = R.sqp_eval(qp, 1.25, v["s"], v["y"], v["z"], v["re"], v["r"])
        qp.close()
        assert rel(O.grd_L(p, p.c, v["y"], v["z"]), g) < 1e-13
        assert rel(O.merit(p, 1.25, p.c, v["s"], p.b, p.d, v["re"], v["r"]), o[:5]) < 1e-13


@pytest.mark.gpu
@pytest.mark.parametrize("cfg", [(5, 3, 6, 1), (12, 4, 64, 4), (20, 10, 300, 0), (40, 10, 64, 0)])
def test_gpu_matches_the_oracle(cfg):
    from hqp_b200.ipcuda import IpCuda
    nx, nu, K, nseg = cfg
    p = synth_lqdocp(nx, nu, K)
    v = vectors(p, 5)
    e = IpCuda(p, nseg=nseg)
    e.update()
    assert rel(e.sqp_grd_L(v["c"], v["y"], v["z"]), O.grd_L(p, v["c"], v["y"], v["z"])) < TOL
    got = e.sqp_merit(0.75, v["c"], v["s"], v["b"], v["d"], v["re"], v["r"])
    want = O.merit(p, 0.75, v["c"], v["s"], v["b"], v["d"], v["re"], v["r"])
    assert rel(got[:4], want[:4]) < TOL
    assert rel(max(got[4], got[5]), want[4]) < TOL
    assert abs(got[1] - (0.75 + got[3] + got[6] + got[7])) <= 1e-12 * max(1.0, abs(got[1]))
    assert rel(e.sqp_quad(v["s"]), O.quad(p, v["s"])) < TOL
    e.close()


@pytest.mark.gpu
def test_gpu_matches_the_live_reference():
    from oracle import refharness as R
    if not R.available():
        pytest.skip("compiled reference (oracle/_ref) not present")
    from hqp_b200.ipcuda import IpCuda
    p = synth_lqdocp(20, 10, 120)
    v = vectors(p, 8)
    qp = R.RefQP(p)
    g, o = R.sqp_eval(qp, -2.0, v["s"], v["y"], v["z"], v["re"], v["r"])
    qp.close()
    e = IpCuda(p)
    e.update()
    assert rel(e.sqp_grd_L(p.c, v["y"], v["z"]), g) < TOL
    got = e.sqp_merit(-2.0, p.c, v["s"], p.b, p.d, v["re"], v["r"])
    assert rel(got[:4], o[:4]) < TOL and rel(max(got[4], got[5]), o[4]) < TOL
    e.close()


@pytest.mark.gpu
def test_gpu_full_size_config2():
    """K = 10^4: against the restatement (numpy finishes in a second)"""
    from hqp_b200.ipcuda import IpCuda
    p = synth_lqdocp(20, 10, 10000)
    v = vectors(p, 1)
    e = IpCuda(p)
    e.update()
    assert rel(e.sqp_grd_L(p.c, v["y"], v["z"]), O.grd_L(p, p.c, v["y"], v["z"])) < TOL
    got = e.sqp_merit(0.0, p.c, v["s"], p.b, p.d, v["re"], v["r"])
    want = O.merit(p, 0.0, p.c, v["s"], p.b, p.d, v["re"], v["r"])
    assert rel(got[:4] / np.maximum(1.0, np.abs(want[:4])), want[:4] / np.maximum(1.0, np.abs(want[:4]))) < TOL
    e.close()
