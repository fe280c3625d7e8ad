"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU
oracle on identical seeded inputs, against the golden fixtures produced by the
unmodified reference, and -- at full BASELINE sizes -- through size-independent
properties (KKT residual, linearity, refinement fixed point)."""
import numpy as np
import pytest

from common import golden_step_cases, load_golden, make_problem, relerr
from hqp_b200.ipcuda import IpCuda, SingularError
from hqp_b200.problem import add_random_stage_ineq, rhs_for, synth_lqdocp, synth_rhs
from oracle.portoracle import PortOracle

pytestmark = pytest.mark.gpu

TOL = 1e-10  # north star: KKT solutions within relative 1e-10 of the reference


@pytest.mark.parametrize("name", golden_step_cases())
@pytest.mark.parametrize("nseg", [1, 0, 3])
def test_step_and_solve_match_reference_golden(name, nseg):
    g, cfg = load_golden(name)
    p = make_problem(*cfg)
    e = IpCuda(p, nseg=nseg)
    e.update()
    e.factor(g["z"], g["w"])
    dx, dy, dz, dw = e.step(g["r1"], g["r2"], g["r3"], g["r4"])
    for mine, key in ((dx, "dx"), (dy, "dy"), (dz, "dz"), (dw, "dw")):
        assert relerr(mine, g[key]) < TOL, (key, nseg)
    sx, sy, sz, sw, res, nsteps = e.solve(g["r1"], g["r2"], g["r3"], g["r4"])
    for mine, key in ((sx, "sx"), (sy, "sy"), (sz, "sz"), (sw, "sw")):
        assert relerr(mine, g[key]) < TOL, (key, nseg)
    assert res <= 1e-10 and nsteps == 1  # the reference needs no refinement either
    e.close()


SHAPES = [  # nx nu K bounds gen fixed nseg
    (2, 1, 1, 1, 0, 1, 1), (2, 1, 2, 1, 0, 1, 0), (1, 1, 9, 1, 0, 1, 3), (3, 1, 33, 1, 1, 1, 4),
    (12, 4, 50, 1, 0, 1, 1), (12, 4, 50, 1, 2, 0, 7), (20, 10, 128, 1, 0, 1, 0),
    (20, 10, 131, 0, 0, 0, 9), (33, 7, 40, 1, 0, 1, 5), (40, 10, 64, 1, 0, 1, 0),
    (8, 24, 40, 1, 0, 1, 4),
]


@pytest.mark.parametrize("cfg", SHAPES)
def test_matches_cpu_oracle(cfg):
    *pc, nseg = cfg
    p = make_problem(*pc)
    z, w, r1, r2, r3, r4 = rhs_for(p, seed=17)
    o = PortOracle(p)
    o.factor(z, w)
    ref = o.step(r1, r2, r3, r4)
    e = IpCuda(p, nseg=nseg)
    e.update()
    e.factor(z, w)
    mine = e.step(r1, r2, r3, r4)
    for a, b in zip(mine, ref):
        assert relerr(a, b) < TOL
    V, R = e.get_factor()
    assert relerr(V[0], o.Vxx()) < TOL and relerr(R[0], o.Rux()) < TOL
    # both residuum implementations agree on the same vectors
    rg = e.residuum(r1, r2, r3, r4, *mine)
    ro = o.residuum(r1, r2, r3, r4, *mine)
    assert abs(rg - ro) <= 1e-12 * max(1.0, ro)
    e.close(); o.close()


def test_batched_instances_match_single():
    """config 3 shape: independent MPC QPs, one CTA per instance."""
    nx, nu, K, B = 12, 4, 50, 5
    probs = [synth_lqdocp(nx, nu, K, seed=100 + i) for i in range(B)]
    rhs = [rhs_for(p, seed=200 + i) for i, p in enumerate(probs)]
    e = IpCuda(probs[0], batch=B)
    e.update(Q=np.stack([p.Q for p in probs]), fx=np.stack([p.fx for p in probs]),
             fu=np.stack([p.fu for p in probs]), ineq_val=np.stack([p.ineq_val for p in probs]))
    cat = [np.concatenate([r[i] for r in rhs]) for i in range(6)]
    e.factor(cat[0], cat[1])
    out = e.step(*cat[2:])
    sizes = (probs[0].N, probs[0].me, probs[0].m, probs[0].m)
    for i, p in enumerate(probs):
        o = PortOracle(p)
        o.factor(rhs[i][0], rhs[i][1])
        ref = o.step(*rhs[i][2:])
        for a, b, n in zip(out, ref, sizes):
            assert relerr(a[i * n:(i + 1) * n], b) < TOL
        o.close()
    e.close()


def test_ill_conditioned_ip_iterate_with_refinement():
    """late-IP regime: z/w over 20 decades; solve() must converge like the oracle's."""
    p = synth_lqdocp(10, 4, 96)
    _, _, r1, r2, r3, r4 = rhs_for(p, seed=8)
    rng = np.random.default_rng(1)
    z = 10.0 ** rng.uniform(-10, 0, p.m)
    w = 10.0 ** rng.uniform(-10, 0, p.m)
    o = PortOracle(p)
    o.factor(z, w)
    ox, oy, oz, ow, ores, _ = o.solve(r1, r2, r3, r4)
    for nseg in (1, 6):
        e = IpCuda(p, nseg=nseg)
        e.update()
        e.factor(z, w)
        dx, dy, dz, dw, res, nsteps = e.solve(r1, r2, r3, r4)
        assert np.isfinite(res) and res <= max(10 * ores, 1e-10)
        assert relerr(dx, ox) < 1e-8 and relerr(dy, oy) < 1e-8
        e.close()
    o.close()


def test_singular_stage_reports_e_sing():
    """Guu == 0 (no cost on u, fu = 0): the reference's BKPsolve raises E_SING
    (meschach/bkpfacto.c:273-286); the ABI must return the same code."""
    p = synth_lqdocp(3, 2, 6, bounds=False)
    p.Q[:] = 0.0
    p.fu[:] = 0.0
    e = IpCuda(p, nseg=1)
    e.update()
    with pytest.raises(SingularError):
        e.factor(np.zeros(0), np.zeros(0))
    e.close()


def test_update_changes_values_factor_reuses_structure():
    p = synth_lqdocp(6, 3, 40, seed=1)
    p2 = synth_lqdocp(6, 3, 40, seed=2)
    z, w, r1, r2, r3, r4 = rhs_for(p, seed=3)
    e = IpCuda(p)
    for prob in (p, p2, p):
        e.update(Q=prob.Q, fx=prob.fx, fu=prob.fu, ineq_val=prob.ineq_val)
        e.factor(z, w)
        mine = e.step(r1, r2, r3, r4)
        o = PortOracle(prob)
        o.factor(z, w)
        ref = o.step(r1, r2, r3, r4)
        for a, b in zip(mine, ref):
            assert relerr(a, b) < TOL
        o.close()
    e.close()


def test_no_inequalities_m_zero():
    """factor/solve with m = 0 (z,w,r3,r4 of dim 0) must work."""
    p = synth_lqdocp(5, 2, 20, bounds=False)
    _, _, r1, r2, r3, r4 = rhs_for(p, seed=4)
    e = IpCuda(p)
    e.update()
    e.factor(np.zeros(0), np.zeros(0))
    dx, dy, dz, dw, res, _ = e.solve(r1, r2, r3, r4)
    assert dz.size == 0 and res <= 1e-10
    e.close()


# ---- full BASELINE sizes: properties that need no CPU oracle ------------------
@pytest.fixture(scope="module")
def c2():
    p = synth_lqdocp(20, 10, 10000)
    e = IpCuda(p)
    e.update()
    rhs = synth_rhs(p)
    e.factor(rhs[0], rhs[1])
    yield p, e, rhs
    e.close()


def test_c2_residual_and_refinement(c2):
    p, e, (z, w, r1, r2, r3, r4) = c2
    assert e.nseg > 1  # the parallel-in-time path is the one under test
    dx, dy, dz, dw, res, nsteps = e.solve(r1, r2, r3, r4)
    assert res <= 1e-10 and nsteps == 1
    assert e.residuum(r1, r2, r3, r4, dx, dy, dz, dw) == res


def test_c2_linearity_and_segment_independence(c2):
    p, e, (z, w, r1, r2, r3, r4) = c2
    a = e.step(r1, r2, r3, r4)
    b = e.step(2 * r1, 2 * r2, 2 * r3, 2 * r4)
    for u, v in zip(a, b):
        assert relerr(v, 2 * u) < 1e-12
    e2 = IpCuda(p, nseg=37)
    e2.update()
    e2.factor(z, w)
    c = e2.step(r1, r2, r3, r4)
    for u, v in zip(a, c):
        assert relerr(v, u) < 1e-10
    e2.close()


def test_c2_prefix_matches_oracle_on_truncated_horizon():
    """the first 300 stages of the C2 workload, checked against the CPU oracle"""
    p = synth_lqdocp(20, 10, 300)
    z, w, r1, r2, r3, r4 = synth_rhs(p)
    o = PortOracle(p); o.factor(z, w)
    ref = o.step(r1, r2, r3, r4)
    e = IpCuda(p); e.update(); e.factor(z, w)
    mine = e.step(r1, r2, r3, r4)
    for a, b in zip(mine, ref):
        assert relerr(a, b) < TOL
    e.close(); o.close()


# ---- full BASELINE sizes against the live reference / the port oracle -----------
def _ref_available():
    from oracle import refharness
    return refharness.available()


@pytest.mark.skipif(not _ref_available(), reason="compiled reference (oracle/_ref) not present")
@pytest.mark.parametrize("shape", [(20, 10, 10000), (40, 10, 2500)])
def test_full_size_step_and_solve_match_live_reference(shape):
    """C2 at its full K = 10^4 and the C5 stage shape at K = 2500: factor, step and
    the refined solve against the UNMODIFIED Hqp_IpLQDOCP (oracle/_ref) run here on
    the same seeded inputs."""
    from oracle import refharness
    nx, nu, K = shape
    p = synth_lqdocp(nx, nu, K)
    z, w, r1, r2, r3, r4 = synth_rhs(p)
    qp = refharness.RefQP(p)
    M = refharness.RefMatrix("LQDOCP", qp)
    M.factor(z, w)
    ref = M.step(z, w, r1, r2, r3, r4)
    rs = M.solve(z, w, r1, r2, r3, r4)
    e = IpCuda(p)
    assert e.nseg > 1
    e.update()
    e.factor(z, w)
    mine = e.step(r1, r2, r3, r4)
    for a, b, key in zip(mine, ref, ("dx", "dy", "dz", "dw")):
        assert relerr(a, b) < TOL, key
    sx, sy, sz, sw, res, nsteps = e.solve(r1, r2, r3, r4)
    for a, b, key in zip((sx, sy, sz, sw), rs[:4], ("sx", "sy", "sz", "sw")):
        assert relerr(a, b) < TOL, key
    assert res <= 1e-10 and rs[4] <= 1e-10
    e.close(); M.close(); qp.close()


# ---- large stage blocks (BASELINE config 4 shape): global-workspace kernels ------
@pytest.mark.parametrize("cfg", [(70, 20, 24, 1, 0, 1, 0), (96, 32, 20, 1, 1, 0, 5),
                                 (200, 50, 16, 1, 0, 1, 0), (130, 7, 12, 0, 0, 1, 1)])
def test_large_blocks_match_cpu_oracle(cfg):
    """nx > 64: the stage blocks do not fit shared memory (hqpcu_create used to
    refuse them); FormGxx / BKP path of hqp/Hqp_IpLQDOCP.C:1077-1111, 1854-1882."""
    *pc, nseg = cfg
    p = make_problem(*pc)
    z, w, r1, r2, r3, r4 = rhs_for(p, seed=41)
    o = PortOracle(p)
    o.factor(z, w)
    ref = o.step(r1, r2, r3, r4)
    e = IpCuda(p, nseg=nseg)
    e.update()
    e.factor(z, w)
    mine = e.step(r1, r2, r3, r4)
    for a, b in zip(mine, ref):
        assert relerr(a, b) < TOL
    V, R = e.get_factor()
    assert relerr(V[0], o.Vxx()) < TOL and relerr(R[0], o.Rux()) < TOL
    sx, sy, sz, sw, res, nsteps = e.solve(r1, r2, r3, r4)
    assert res <= 1e-10
    e.close(); o.close()


def test_large_blocks_two_sweep_tree(monkeypatch):
    """the up / down sweep of the binary tree (HQPCU_HS=0) on the global-workspace path:
    elem_compose_kernel<0> / elem_scan_kernel<0> with the blocked elimination"""
    monkeypatch.setenv("HQPCU_HS", "0")
    p = make_problem(70, 20, 48, 1, 0, 1)
    z, w, r1, r2, r3, r4 = rhs_for(p, seed=43)
    o = PortOracle(p)
    o.factor(z, w)
    ref = o.step(r1, r2, r3, r4)
    e = IpCuda(p)
    assert e.nseg > 2
    e.update()
    e.factor(z, w)
    mine = e.step(r1, r2, r3, r4)
    for a, b in zip(mine, ref):
        assert relerr(a, b) < TOL
    e.close(); o.close()


@pytest.mark.skipif(not _ref_available(), reason="compiled reference (oracle/_ref) not present")
def test_c4_truncated_horizon_matches_live_reference():
    """BASELINE config 4 (nx=200 nu=50) on a K=100 truncation against the
    UNMODIFIED Hqp_IpLQDOCP, parallel in time (nseg chosen automatically)."""
    from oracle import refharness
    p = synth_lqdocp(200, 50, 100)
    z, w, r1, r2, r3, r4 = synth_rhs(p)
    qp = refharness.RefQP(p)
    M = refharness.RefMatrix("LQDOCP", qp)
    M.factor(z, w)
    ref = M.step(z, w, r1, r2, r3, r4)
    e = IpCuda(p)
    assert e.nseg > 1
    e.update()
    e.factor(z, w)
    mine = e.step(r1, r2, r3, r4)
    for a, b, key in zip(mine, ref, ("dx", "dy", "dz", "dw")):
        assert relerr(a, b) < TOL, key
    e.close(); M.close(); qp.close()


def test_c3_full_batch_sampled_instances_match_oracle():
    """config 3 at its full size: 4096 DIFFERENT instances nx=12 nu=4 K=50 in one
    batch; sampled instances against the port oracle."""
    nx, nu, K, B = 12, 4, 50, 4096
    probs = [synth_lqdocp(nx, nu, K, seed=5000 + i) for i in range(B)]
    p0 = probs[0]
    rng = np.random.default_rng(3)
    z = 0.5 + rng.uniform(0, 1, (B, p0.m)); w = 0.5 + rng.uniform(0, 1, (B, p0.m))
    r1 = rng.uniform(-1, 1, (B, p0.N)); r2 = rng.uniform(-1, 1, (B, p0.me))
    r3 = rng.uniform(-1, 1, (B, p0.m)); r4 = rng.uniform(-1, 1, (B, p0.m))
    e = IpCuda(p0, batch=B)
    e.update(Q=np.stack([p.Q for p in probs]), fx=np.stack([p.fx for p in probs]),
             fu=np.stack([p.fu for p in probs]), ineq_val=np.stack([p.ineq_val for p in probs]))
    e.factor(z.ravel(), w.ravel())
    out = e.step(r1.ravel(), r2.ravel(), r3.ravel(), r4.ravel())
    sizes = (p0.N, p0.me, p0.m, p0.m)
    for i in (0, 1, 777, 2048, 4095):
        o = PortOracle(probs[i])
        o.factor(z[i], w[i])
        ref = o.step(r1[i], r2[i], r3[i], r4[i])
        for a, b, n in zip(out, ref, sizes):
            assert relerr(a[i * n:(i + 1) * n], b) < TOL
        o.close()
    e.close()


# ---- general stage equality rows (terminal constraints) ------------------------
from common import dense_kkt_solve, eq_cases, make_eq_problem  # noqa: E402
import os  # noqa: E402


@pytest.mark.parametrize("name", sorted(eq_cases()))
@pytest.mark.parametrize("nseg", [1, 0])
def test_equality_rows_match_reference_golden(name, nseg, golden_dir):
    """The reference eliminates stage equalities with its GE_QP null-space step
    (hqp/Hqp_IpLQDOCP.C:1883-1938); the block-elimination path must give the
    same KKT solution."""
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    p = make_eq_problem(*eq_cases()[name])
    e = IpCuda(p, nseg=nseg)
    e.update()
    e.factor(g["z"], g["w"])
    dx, dy, dz, dw, res, nsteps = e.solve(g["r1"], g["r2"], g["r3"], g["r4"])
    for mine, key in ((dx, "sx"), (dy, "sy"), (dz, "sz"), (dw, "sw")):
        assert relerr(mine, g[key]) < 1e-9, (key, relerr(mine, g[key]))
    assert res <= 1e-10
    e.close()


def test_equality_rows_match_dense_kkt():
    p = make_eq_problem(4, 2, 7, True, 1, True, [7, 3], 2)
    z, w, r1, r2, r3, r4 = rhs_for(p, seed=12)
    ref = dense_kkt_solve(p, z, w, r1, r2, r3, r4)
    e = IpCuda(p)
    e.update()
    e.factor(z, w)
    mine = e.step(r1, r2, r3, r4)
    for a, b in zip(mine, ref):
        assert relerr(a, b) < 1e-9
    assert e.residuum(r1, r2, r3, r4, *mine) < 1e-10
    e.close()


@pytest.mark.parametrize("env", [{"HQPCU_SEG_WARPS": "1"}, {"HQPCU_GRAPHS": "0"},
                                 {"HQPCU_PDL": "1"}, {"HQPCU_CHAIN_CHUNK": "7"},
                                 {"HQPCU_LANE_GROUP": "16"}, {"HQPCU_SPW": "2"}, {"HQPCU_HS": "0"}])
@pytest.mark.parametrize("cfg", [(20, 10, 131, 1, 0, 1, 9), (12, 4, 50, 1, 2, 0, 1),
                                 (40, 10, 64, 1, 0, 1, 4)])
def test_launch_variants_match_cpu_oracle(cfg, env, monkeypatch):
    """The tuning switches read at hqpcu_create select other kernels / launch
    paths of the same algorithm (one warp per segment, plain launches instead
    of CUDA graphs, programmatic dependent launch, another TMA chunk size): all
    must give the oracle's answer, twice in a row (graph replay)."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    *pc, nseg = cfg
    p = make_problem(*pc)
    z, w, r1, r2, r3, r4 = rhs_for(p, seed=23)
    o = PortOracle(p)
    o.factor(z, w)
    ref = o.step(r1, r2, r3, r4)
    e = IpCuda(p, nseg=nseg)
    e.update()
    for _ in range(2):
        e.factor(z, w)
        mine = e.step(r1, r2, r3, r4)
        for a, b in zip(mine, ref):
            assert relerr(a, b) < TOL
    e.close(); o.close()


def test_many_small_instances_one_warp_each():
    """enough instances for the automatic one-warp-per-instance mode (>= 8 per SM)"""
    nx, nu, K, B = 12, 4, 20, 1500
    p = synth_lqdocp(nx, nu, K, seed=7)
    z, w, r1, r2, r3, r4 = rhs_for(p, seed=8)
    o = PortOracle(p)
    o.factor(z, w)
    ref = o.step(r1, r2, r3, r4)
    e = IpCuda(p, batch=B)
    bc = lambda a: np.broadcast_to(a, (B,) + a.shape)
    e.update(Q=bc(p.Q), fx=bc(p.fx), fu=bc(p.fu), ineq_val=bc(p.ineq_val))
    rep = lambda a: np.tile(a, B)
    e.factor(rep(z), rep(w))
    out = e.step(rep(r1), rep(r2), rep(r3), rep(r4))
    for got, want in zip(out, ref):
        got = got.reshape(B, -1)
        assert relerr(got[0], want) < TOL and relerr(got[B - 1], want) < TOL
        assert np.array_equal(got[0], got[B // 2])  # identical inputs, identical bits
    e.close(); o.close()


def test_update_from_sparse_values_equals_dense_update():
    """row f1: hqpcu_set_value_map + hqpcu_update_values (device-side scatter of
    the CSR values) must leave exactly the factor that the dense update leaves."""
    p = make_problem(20, 10, 64, 1, 2, 1)
    z, w, r1, r2, r3, r4 = rhs_for(p, seed=31)
    a = IpCuda(p, nseg=4)
    a.update()
    a.factor(z, w)
    ref = a.step(r1, r2, r3, r4)
    Va, Ra = a.get_factor()
    b = IpCuda(p, nseg=4)
    vals, dst, dst2 = p.value_map()
    b.set_value_map(dst, dst2)
    for _ in range(2):                   # the map is registered once, values per update
        b.update_values(vals)
        b.factor(z, w)
        out = b.step(r1, r2, r3, r4)
        Vb, Rb = b.get_factor()
        assert np.array_equal(Va, Vb) and np.array_equal(Ra, Rb)
        for x, y in zip(out, ref):
            assert np.array_equal(x, y)
    with pytest.raises(Exception):
        b.set_value_map(np.array([10 ** 12]), np.array([-1]))
    a.close(); b.close()


# ---- indefinite stage blocks: scaled, pivoted factorisation of Guu (a16) ----------
@pytest.mark.parametrize("dims", [(12, 4, 50), (20, 10, 96), (40, 10, 64)])
@pytest.mark.parametrize("nseg", [1, 0])
def test_indefinite_guu_matches_bunch_kaufman_oracle(dims, nseg):
    """Huu made indefinite at a few stages (a non-convex Hessian block as BFGS can
    deliver it before eigenvalue control): the reference factors Guu with a scaled
    Bunch-Kaufman (hqp/Hqp_IpLQDOCP.C:1860-1879, meschach/bkpfacto.c:102-226) and
    solves the KKT system all the same; the oracle restates that factorisation.
    The device path must switch from LDL^T without interchanges to its scaled,
    pivoted inverse for those blocks -- and, parallel in time, fall back to the
    sequential sweep for THIS factor only (the next update restores the segments)."""
    nx, nu, K = dims
    p = synth_lqdocp(nx, nu, K, seed=21)
    bad = [3, K // 2, K - 2]
    for k in bad:
        p.Q[k, nx:, nx:] -= 4.0 * np.eye(nu)       # Huu indefinite (eigenvalues ~ -3.9 .. -2)
        p.Q[k, nx, nx + 1] += 0.7                   # and not diagonal
        p.Q[k, nx + 1, nx] += 0.7
    z, w, r1, r2, r3, r4 = rhs_for(p, seed=5)
    z *= 0.05                                       # weak barrier terms: Guu stays indefinite
    o = PortOracle(p)
    o.factor(z, w)
    ref = o.step(r1, r2, r3, r4)
    e = IpCuda(p, nseg=nseg)
    nseg0 = e.nseg
    e.update()
    e.factor(z, w)
    mine = e.step(r1, r2, r3, r4)
    for a, b, key in zip(mine, ref, ("dx", "dy", "dz", "dw")):
        assert relerr(a, b) < 1e-9, (key, relerr(a, b))
    sx, sy, sz, sw, res, nsteps = e.solve(r1, r2, r3, r4)
    assert res <= 1e-10
    if nseg == 0:
        assert e.nseg == 1 and nseg0 > 1            # demoted for these matrix values ...
        p2 = synth_lqdocp(nx, nu, K, seed=22)       # ... and restored by the next update
        e.update(Q=p2.Q, fx=p2.fx, fu=p2.fu, ineq_val=p2.ineq_val)
        assert e.nseg == nseg0
        o2 = PortOracle(p2)
        o2.factor(z, w)
        ref2 = o2.step(r1, r2, r3, r4)
        e.factor(z, w)
        for a, b in zip(e.step(r1, r2, r3, r4), ref2):
            assert relerr(a, b) < TOL
        o2.close()
    e.close(); o.close()
