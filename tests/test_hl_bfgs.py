"""Row f2: block-diagonal BFGS update of the Lagrangian Hessian.

CPU part: the numpy restatement (oracle/hl_bfgs_oracle.py) against the golden vectors of
the UNMODIFIED Hqp_HL_BFGS::update_b_Q (hqp/Hqp_HL_BFGS.C:149-213) and, where oracle/_ref
is present, against the live reference; the C-ABI library exports what its header declares.
GPU part (-m gpu): hqphl_bfgs_update through the C ABI against the golden vectors, the
oracle on mixed block structures, and properties at full size."""
import ctypes
import os
import re

import numpy as np
import pytest

from oracle import hl_bfgs_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "hl_bfgs_blocks.npz")
TOL = 1e-10


def golden_cases():
    d = np.load(GOLD)
    for i in range(int(d["n"])):
        alpha, gamma, eps, ec = d[f"par{i}"]
        yield d[f"Q{i}"], d[f"s{i}"], d[f"u{i}"], float(alpha), float(gamma), float(eps), bool(ec), d[f"out{i}"]


def upper_err(a, b):
    n = a.shape[0]
    iu = np.triu_indices(n)
    return float(np.max(np.abs(a[iu] - b[iu])) / max(1.0, np.max(np.abs(b))))


def test_oracle_matches_the_reference_goldens():
    n = 0
    for Q, s, u, alpha, gamma, eps, ec, want in golden_cases():
        got, _, _ = O.update_block(Q, s, u, alpha, gamma, eps, ec)
        assert upper_err(got, want) < 1e-12
        n += 1
    assert n == 52


def test_oracle_matches_the_live_reference():
    from oracle import refharness as R
    if not R.available():
        pytest.skip("compiled reference (oracle/_ref) not present")
    rng = np.random.default_rng(11)
    for t in range(12):
        n = int(rng.integers(2, 45))
        M = rng.uniform(-1, 1, (n, n))
        Q = M.T @ M / n + 0.05 * np.eye(n)
        s, u = rng.uniform(-1, 1, n), rng.uniform(-1, 1, n) * (1 if t % 2 else -1)
        for ec in (True, False):
            want = R.hl_bfgs_block(Q, s, u, 0.6, -0.3, 1e-8, ec)
            got, _, _ = O.update_block(Q, s, u, 0.6, -0.3, 1e-8, ec)
            assert upper_err(got, want) < 1e-12


def test_hl_library_exports_every_declared_symbol():
    from hqp_b200 import hlcuda
    text = open(os.path.join(ROOT, "include", "hqp_hlcuda.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = sorted(set(re.findall(r"\b(hqphl_\w+)\s*\(", text)))
    assert "hqphl_bfgs_update" in names and "hqphl_bfgs_update_dev" in names
    lib = ctypes.CDLL(hlcuda.LIB_PATH)
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/hqp_hlcuda.h but not exported"


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from hqp_b200 import hlcuda
    with pytest.raises(RuntimeError):
        hlcuda.bfgs_update([2], np.eye(2).ravel(), np.ones(2), np.ones(2), 1.0)


# ------------------------------------------------------------------ GPU
@pytest.mark.gpu
def test_gpu_matches_the_reference_goldens():
    from hqp_b200 import hlcuda
    for Q, s, u, alpha, gamma, eps, ec, want in golden_cases():
        n = Q.shape[0]
        got, info = hlcuda.bfgs_update([n], Q.ravel(), s, u, alpha, gamma, eps, ec)
        assert upper_err(got.reshape(n, n), want) < TOL
        assert info[2] == 0


@pytest.mark.gpu
def test_gpu_mixed_blocks_match_the_oracle():
    """a DOCP's Hessian: K blocks of nx+nu and a last one of nx (Hqp_HL_BFGS::update walks
    them with next_block, hqp/Hqp_HL_BFGS.C:216-243), plus ragged sizes"""
    from hqp_b200 import hlcuda
    rng = np.random.default_rng(7)
    for bs in ([30] * 40 + [20], [5, 1, 17, 64, 2, 33, 8], [50] * 9 + [40]):
        Q = np.concatenate([(lambda m, n: (m.T @ m / n + 0.05 * np.eye(n)).ravel())(rng.uniform(-1, 1, (n, n)), n)
                            for n in bs])
        nv = sum(bs)
        s, u = rng.uniform(-1, 1, nv), rng.uniform(-1, 1, nv)
        u[: nv // 2] *= -1.0  # half of the blocks see negative curvature
        for ec in (True, False):
            got, info = hlcuda.bfgs_update(bs, Q, s, u, 0.5, -0.1, 1e-8, ec)
            want, nshift, nskip = O.update(bs, Q, s, u, 0.5, -0.1, 1e-8, ec)
            qo = 0
            for n in bs:
                assert upper_err(got[qo:qo + n * n].reshape(n, n), want[qo:qo + n * n].reshape(n, n)) < TOL
                qo += n * n
            assert info[0] == nshift and info[1] == nskip and info[2] == 0


@pytest.mark.gpu
def test_gpu_full_size_properties():
    """config 2's Hessian (10^4 blocks of 30 and one of 20): with eigenvalue control every
    updated block is symmetric and its smallest eigenvalue is >= eps^2 (the property the
    reference enforces, :204-212); sampled blocks agree with the oracle."""
    from hqp_b200 import hlcuda
    rng = np.random.default_rng(3)
    K, n = 10000, 30
    bs = [n] * K + [20]
    M = rng.uniform(-1, 1, (K + 1, n, n))
    Qb = np.einsum("kij,kil->kjl", M, M) / n + 0.05 * np.eye(n)
    Q = np.concatenate([Qb[:K].ravel(), Qb[K, :20, :20].ravel()])
    nv = K * n + 20
    s, u = rng.uniform(-1, 1, nv), rng.uniform(-1, 1, nv)
    got, info = hlcuda.bfgs_update(bs, Q, s, u, 1.0, 0.1, 1e-8, True)
    assert info[2] == 0
    G = got[: K * n * n].reshape(K, n, n)
    assert np.max(np.abs(G - G.transpose(0, 2, 1))) == 0.0
    lam = np.linalg.eigvalsh(G[::97])
    assert lam.min() >= 1e-16 - 1e-12
    for k in (0, 1234, K - 1):
        want, _, _ = O.update_block(Qb[k], s[k * n:(k + 1) * n], u[k * n:(k + 1) * n], 1.0, 0.1, 1e-8, True)
        assert upper_err(G[k], want) < TOL


def _module_worker(hela, q):
    """one process per module: the Hessian of a small DOCP after Hqp_HL::setup + ::update"""
    import os as _os
    import sys as _sys
    _sys.path.insert(0, ROOT)
    from hqp_b200.problem import synth_lqdocp
    from oracle import refharness as R
    if hela != "BFGS":
        R.load_plugin(_os.path.join(ROOT, "hqp_b200", "lib", "libhqp_ipcuda_plugin.so"))
    nx, nu, K = 7, 4, 40
    p = synth_lqdocp(nx, nu, K)
    rng = np.random.default_rng(9)
    s, u = rng.uniform(-1, 1, p.N), rng.uniform(-1, 1, p.N)
    u[::3] *= -1.0
    out = []
    for ec in (True, False):
        qp = R.RefQP(p)
        R.hl_update(qp, hela, s, u, 0.6, -0.2, 1e-8, ec)
        out.append([R.qp_get_Q_block(qp, k * (nx + nu), nx + nu) for k in range(K)] +
                   [R.qp_get_Q_block(qp, K * (nx + nu), nx)])
        qp.close()
    q.put(out)


@pytest.mark.gpu
def test_module_cudabfgs_matches_the_reference_module():
    """drop-in at the module boundary: sqp_hela CudaBFGS (hqp_b200/host/Hqp_HL_CudaBFGS.C, in
    the plugin library) against the reference's sqp_hela BFGS, both driven through
    Hqp_HL::setup / ::update on the same Hqp_Program by the unmodified host code"""
    from oracle import refharness as R
    plugin = os.path.join(ROOT, "hqp_b200", "lib", "libhqp_ipcuda_plugin.so")
    if not (R.available() and os.path.exists(plugin)):
        pytest.skip("oracle/_ref or the plugin library were not built (needs /root/reference at build time)")
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    res = {}
    for hela in ("BFGS", "CudaBFGS"):
        q = ctx.Queue()
        pr = ctx.Process(target=_module_worker, args=(hela, q))
        pr.start()
        res[hela] = q.get(timeout=300)
        pr.join()
    for a_blocks, b_blocks in zip(res["BFGS"], res["CudaBFGS"]):
        for a, b in zip(a_blocks, b_blocks):
            assert upper_err(b, a) < TOL
