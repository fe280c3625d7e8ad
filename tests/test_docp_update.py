"""Row f4: the stage loop of Hqp_Docp::update / ::update_fbd on the GPU
(libhqpdocp.so, include/hqp_docpcuda.h).

CPU (-m "not gpu"): the plain-Python restatement oracle/docp_oracle.py against the golden
vectors made by the unmodified reference (tests/golden/docp_update_*.npz) and, where
oracle/_ref exists, against the live reference on more shapes; the host-side bound parsing;
the C-ABI library's exports.
GPU (-m gpu): the CUDA path through the C ABI against the golden vectors and the restatement.

Tolerances.  Values (f_k, c_k, b, d) and forward differences are the same IEEE operations in
the same order on both sides (the library is built with -fmad=false): bit-exact is expected
and asserted for b, d; the objective f is a re-associated sum over stages (1e-13 relative);
difference quotients are asserted to 1e-9 absolute (a last-bit difference in a model value
would show up as 1e-10 after the division by dv >= 1e-6).  Dual-number derivatives against
the closed form: 1e-12."""
import ctypes
import glob
import os
import re

import numpy as np
import pytest

from hqp_b200 import docpcuda as dc
from oracle import docp_oracle as do
from oracle import refharness as rh

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CASES = {
    "did_k12": lambda: dc.did_problem(12, True),
    "did_k60_nocns": lambda: dc.did_problem(60, False),
    "synthnl_n4m2c1K6": lambda: dc.synthnl_problem(6, 4, 2, 1, 0),
    "synthnl_n6m3c3K5": lambda: dc.synthnl_problem(5, 6, 3, 3, 2),
    "synthnl_n20m10c1K8": lambda: dc.synthnl_problem(8, 20, 10, 1, 1),
}


def golden(golden_dir, name):
    return np.load(os.path.join(golden_dir, f"docp_update_{name}.npz"))


def grads_kind(p):
    return "did" if p.model == dc.MODEL_DID else "fd"


# ------------------------------------------------------------------ CPU: oracle pinned

def test_golden_files_present(golden_dir):
    assert len(glob.glob(os.path.join(golden_dir, "docp_update_*.npz"))) == len(CASES)


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_matches_reference_golden(golden_dir, name):
    p, g = CASES[name](), golden(golden_dir, name)
    assert (p.N, p.me, p.m) == (g["x"].size, g["b"].size, g["d"].size)  # parse_constr restated
    o = do.update(p, g["x"], grads_kind(p))
    A, C = do.assemble_AC(p, o["fx"], o["fu"], o["cx"], o["cu"])
    assert o["f"] == float(g["f"])
    for key, want in (("b", g["b"]), ("d", g["d"]), ("g", g["c"])):
        assert np.array_equal(o[key], want), key
    assert np.array_equal(A, g["A"]) and np.array_equal(C, g["C"])
    f2, b2, d2, _ = do.update_fbd(p, g["x2"])
    assert f2 == float(g["f2"]) and np.array_equal(b2, g["b2"]) and np.array_equal(d2, g["d2"])


@pytest.mark.skipif(not rh.available(), reason="oracle/_ref not built")
@pytest.mark.parametrize("dims", [(7, 3, 1, 0, 0), (3, 1, 1, 1, 1), (9, 5, 5, 4, 3), (4, 12, 4, 2, 0)])
def test_oracle_matches_live_reference(dims):
    p = dc.synthnl_problem(*dims, seed=77)
    r = rh.RefDocp(p)
    try:
        assert (r.N, r.me, r.m) == (p.N, p.me, p.m)
        assert np.array_equal(r.x(), p.x_init)
        x = p.x_init + 0.2 * np.random.default_rng(1).uniform(-1, 1, p.N)
        ref, o = r.update(x), do.update(p, x, "fd")
        A, C = do.assemble_AC(p, o["fx"], o["fu"], o["cx"], o["cu"])
        assert o["f"] == ref["f"] and np.array_equal(o["b"], ref["b"]) and np.array_equal(o["d"], ref["d"])
        assert np.array_equal(o["g"], ref["c"]) and np.array_equal(A, ref["A"]) and np.array_equal(C, ref["C"])
    finally:
        r.close()


def test_exact_derivatives_agree_with_differences():
    p = dc.synthnl_problem(5, 6, 3, 3, 2)
    x = p.x_init
    fd, ex = do.update(p, x, "fd"), do.update(p, x, "exact")
    for key in ("fx", "fu", "g", "cx", "cu"):
        assert np.max(np.abs(fd[key] - ex[key])) < 5e-4, key


def test_parse_constr_tables():
    p = dc.did_problem(4, True)
    # x0 fixed (2), final state fixed (2); x[1] <= 0.01 at k = 1..3; c <= 0.01 at k = 0..3
    assert p.xu_eq.idxs == [0, 1, 12, 13] and p.xu_eq.vals == [1.0, 0.0, -1.0, 0.0]
    assert p.xu_lb.idxs == [] and p.xu_ub.idxs == [4, 7, 10] and p.cns_ub.idxs == [0, 1, 2, 3]
    assert (p.N, p.me, p.m) == (14, 12, 7)
    with pytest.raises(ValueError):
        dc.parse_constr([dc.INF], [dc.INF], 0, dc.Assoc(), dc.Assoc(), dc.Assoc())


# ------------------------------------------------------------------ CPU: boundary

def declared():
    text = open(os.path.join(ROOT, "include", "hqp_docpcuda.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(hqpdocp_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(dc.LIB_PATH) if os.path.exists(dc.LIB_PATH) else dc.lib()
    names = declared()
    assert {"hqpdocp_create", "hqpdocp_update", "hqpdocp_update_fbd", "hqpdocp_update_dev"} <= set(names)
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/hqp_docpcuda.h but not exported"


def test_library_is_sm100a_sass_without_local_memory():
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", dc.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    assert "sm_100a" in out.stdout
    # the views keep every model evaluation in registers / shared memory
    res = subprocess.run(["cuobjdump", "-res-usage", dc.LIB_PATH], capture_output=True, text=True).stdout
    kernels = re.findall(r"Function (\S*docp_(?:stage|vals)_kernel\S*):\s*\n\s*(.*)", res)
    assert kernels
    for name, usage in kernels:
        assert "LOCAL:0" in usage, (name, usage)


def test_bad_arguments_and_no_cpu_fallback():
    lib = dc.lib()
    h = ctypes.c_void_p()
    assert lib.hqpdocp_create(None, ctypes.byref(h)) == 1
    p = dc.did_problem(4)
    p.model = 99
    with pytest.raises(RuntimeError, match="status 2"):
        dc.DocpCuda(p)
    p = dc.did_problem(4)
    p.k_first, p.K_total = 3, 5  # a stage range that does not fit its horizon
    with pytest.raises(RuntimeError, match="status 1"):
        dc.DocpCuda(p)
    p = dc.did_problem(4)
    p.nx = 3  # does not fit the model
    with pytest.raises(RuntimeError, match="status 2"):
        dc.DocpCuda(p)
    import torch
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="status 100"):
            dc.DocpCuda(dc.did_problem(4))


# ------------------------------------------------------------------ GPU: parity

def _check_against(p, out, want_f, want_b, want_d, want_c, want_A, want_C, jac_tol):
    assert abs(out["f"] - want_f) <= 1e-13 * max(1.0, abs(want_f))
    assert np.array_equal(out["b"], want_b), float(np.max(np.abs(out["b"] - want_b)))
    assert np.array_equal(out["d"], want_d)
    assert np.max(np.abs(out["g"] - want_c)) <= jac_tol
    A, C = do.assemble_AC(p, out["fx"], out["fu"], out["cx"], out["cu"])
    assert np.max(np.abs(A - want_A)) <= jac_tol
    if p.m:
        assert np.max(np.abs(C - want_C)) <= jac_tol


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_gpu_update_matches_reference_golden(golden_dir, name):
    p, g = CASES[name](), golden(golden_dir, name)
    e = dc.DocpCuda(p)
    try:
        if p.model == dc.MODEL_DID:
            # Prg_DID supplies its own derivatives (hqp_docp/Prg_DID.C:101-165): dual numbers
            # reproduce them exactly
            out = e.update(g["x"], dc.GRAD_AD)
            _check_against(p, out, float(g["f"]), g["b"], g["d"], g["c"], g["A"], g["C"], 0.0)
        else:
            out = e.update(g["x"], dc.GRAD_FD)  # the reference's default update_grds
            _check_against(p, out, float(g["f"]), g["b"], g["d"], g["c"], g["A"], g["C"], 1e-9)
        f2, b2, d2 = e.update_fbd(g["x2"])
        assert abs(f2 - float(g["f2"])) <= 1e-13 * max(1.0, abs(float(g["f2"])))
        assert np.array_equal(b2, g["b2"]) and np.array_equal(d2, g["d2"])
        assert e.launches > 0
    finally:
        e.close()


@pytest.mark.gpu
@pytest.mark.parametrize("dims", [(7, 3, 1, 0, 0), (3, 1, 1, 1, 1), (40, 12, 4, 3, 2), (33, 40, 10, 1, 0),
                                  (5, 64, 32, 8, 8)])
def test_gpu_update_matches_restatement(dims):
    p = dc.synthnl_problem(*dims, seed=11)
    x = p.x_init + 0.2 * np.random.default_rng(2).uniform(-1, 1, p.N)
    e = dc.DocpCuda(p)
    try:
        fd, ad = e.update(x, dc.GRAD_FD), e.update(x, dc.GRAD_AD)
    finally:
        e.close()
    o, ex = do.update(p, x, "fd"), do.update(p, x, "exact")
    A, C = do.assemble_AC(p, o["fx"], o["fu"], o["cx"], o["cu"])
    _check_against(p, fd, o["f"], o["b"], o["d"], o["g"], A, C, 1e-9)
    for key in ("fx", "fu", "g", "cx", "cu"):
        if ex[key].size:
            assert np.max(np.abs(ad[key] - ex[key])) <= 1e-12 * max(1.0, np.max(np.abs(ex[key]))), key
    assert np.array_equal(ad["b"], o["b"]) and np.array_equal(ad["d"], o["d"])


def _numpy_values(p, x):
    """Vectorised values of the synthetic model (rounding differs from the scalar order)."""
    nx, nu, K, nd = p.nx, p.nu, p.K, p.nd
    par = p.par
    eps = par[0]
    A = par[1:1 + nx * nx].reshape(nx, nx)
    B = par[1 + nx * nx:1 + nx * nx + nx * nu].reshape(nx, nu)
    qw = par[1 + nx * nx + nx * nu:1 + nx * nx + nx * nu + nx]
    rw = par[1 + nx * nx + nx * nu + nx:]
    xs = np.concatenate([x, np.zeros(nu)]).reshape(K + 1, nd)
    X, U = xs[:, :nx], xs[:K, nx:]
    f = X[:K] @ A.T + U @ B.T + eps * X[:K] / (1 + X[:K] ** 2)
    e = X - p.spar
    f0 = 0.5 * (e * e * qw).sum(1)
    nm = min(nx, nu)
    f0[:K] += 0.5 * (U * U * rw).sum(1) + eps * (X[:K, :nm] * U[:, :nm]).sum(1)
    return f, f0.sum(), X, U, A, B, eps


@pytest.mark.gpu
def test_gpu_update_full_size_properties():
    """Config 2's horizon (K = 10^4, nx 20, nu 10): values against a vectorised numpy
    evaluation, dual-number Jacobians against the closed form, differences against them, and the
    device-pointer entry point against the host one."""
    import torch
    p = dc.synthnl_problem(10_000, 20, 10, 1, 1, seed=3)
    x = p.x_init
    e = dc.DocpCuda(p)
    try:
        ad, fd = e.update(x, dc.GRAD_AD), e.update(x, dc.GRAD_FD)
        f, fsum, X, U, A, B, eps = _numpy_values(p, x)
        nd = p.nd
        xn = np.concatenate([x, np.zeros(p.nu)]).reshape(p.K + 1, nd)[1:, :p.nx]
        assert np.max(np.abs(ad["b"][:p.K * p.nx].reshape(p.K, p.nx) - (f - xn))) < 1e-13
        assert abs(ad["f"] - fsum) < 1e-12 * abs(fsum)
        Xk = X[:p.K]
        want_fx = A[None] + eps * ((1 - Xk ** 2) / (1 + Xk ** 2) ** 2)[:, :, None] * np.eye(p.nx)[None]
        assert np.max(np.abs(ad["fx"] - want_fx)) < 1e-13
        assert np.max(np.abs(ad["fu"] - B[None])) < 1e-15
        for key in ("fx", "fu", "g", "cx", "cu"):
            assert np.max(np.abs(fd[key] - ad[key])) < 1e-3, key
        # fixed x0 rows and the box bounds on u through the association tables
        o = p.K * p.nx
        assert np.array_equal(ad["b"][o:o + p.nx], np.zeros(p.nx))
        # device-pointer entry point
        dev = torch.device("cuda:0")
        t = lambda n: torch.empty(max(1, n), dtype=torch.float64, device=dev)
        xd = torch.from_numpy(x).to(dev)
        fo, b, d, g = t(1), t(p.me), t(p.m), t(p.N)
        fx, fu, cx, cu = t(p.K * p.nx * p.nx), t(p.K * p.nx * p.nu), t(p.ncns * p.nx), t(p.K * p.nc * p.nu)
        e.set_stream(torch.cuda.current_stream().cuda_stream)
        e.update_dev(xd, fo, b, d, g, fx, fu, cx, cu, dc.GRAD_AD)
        torch.cuda.synchronize()
        assert fo.item() == ad["f"]
        assert np.array_equal(b.cpu().numpy(), ad["b"]) and np.array_equal(d.cpu().numpy()[:p.m], ad["d"])
        assert np.array_equal(fx.cpu().numpy().reshape(ad["fx"].shape), ad["fx"])
        assert np.array_equal(g.cpu().numpy(), ad["g"])
        b.zero_()
        e.update_fbd_dev(xd, fo, b, d)
        torch.cuda.synchronize()
        assert np.array_equal(b.cpu().numpy(), ad["b"])
    finally:
        e.close()


# ------------------------------------------------------------------ GPU: the host mix-in
# hqp_b200/host/Hqp_DocpCuda.h under the UNMODIFIED reference: the same program once as the
# reference runs it (Hqp_Docp::update on the host) and once with the stage loop on the GPU.

def _ref_and_cuda(p, x, grad=None):
    r = rh.RefDocp(p)
    try:
        want = r.update(x)
        want_fbd = r.update(x * 0.9, fbd_only=True)
    finally:
        r.close()
    c = rh.RefDocp(p, cuda=True)
    try:
        if grad is not None:
            assert rh.lib().ref_set_int(b"prg_cuda_grad", grad) == 0
        got = c.update(x)
        got_fbd = c.update(x * 0.9, fbd_only=True)
    finally:
        c.close()
    return want, got, want_fbd, got_fbd


@pytest.mark.gpu
@pytest.mark.skipif(not rh.available(), reason="oracle/_ref not built")
def test_gpu_host_module_did_update_is_the_reference_update():
    p = dc.did_problem(60, True)
    x = p.x_init + 0.05 * np.random.default_rng(4).uniform(-1, 1, p.N)
    want, got, wf, gf = _ref_and_cuda(p, x)
    assert abs(got["f"] - want["f"]) <= 1e-13 * max(1.0, abs(want["f"]))
    for key in ("b", "d", "c", "A", "C"):
        assert np.array_equal(got[key], want[key]), key
    assert np.array_equal(gf["b"], wf["b"]) and np.array_equal(gf["d"], wf["d"])


@pytest.mark.gpu
@pytest.mark.skipif(not rh.available(), reason="oracle/_ref not built")
@pytest.mark.parametrize("grad,tol", [(0, 1e-9), (1, 5e-4)])
def test_gpu_host_module_synthnl_update(grad, tol):
    p = dc.synthnl_problem(9, 6, 3, 3, 2, seed=21)
    x = p.x_init + 0.1 * np.random.default_rng(6).uniform(-1, 1, p.N)
    want, got, wf, gf = _ref_and_cuda(p, x, grad)
    assert abs(got["f"] - want["f"]) <= 1e-13 * max(1.0, abs(want["f"]))
    assert np.array_equal(got["b"], want["b"]) and np.array_equal(got["d"], want["d"])
    for key in ("c", "A", "C"):
        assert np.max(np.abs(got[key] - want[key])) <= tol, key
    assert np.array_equal(gf["b"], wf["b"]) and np.array_equal(gf["d"], wf["d"])


@pytest.mark.gpu
@pytest.mark.skipif(not rh.available(), reason="oracle/_ref not built")
def test_gpu_host_module_sqp_solves_match_the_reference():
    """hqp_solve (the unmodified Hqp_SqpPowell + Hqp_IpsFranke + Hqp_IpLQDOCP) on the program with
    its stage loop on the GPU: the shipped example reproduces the reference run (SURVEY.md
    section 6: objective 98.3999988628, 1 SQP iteration, 57 QP iterations), the nonlinear
    program takes the reference's iterations to the reference's optimum."""
    for make, grad in ((lambda: dc.did_problem(60, True), 1), (lambda: dc.synthnl_problem(20, 4, 2, 1, 0), 0)):
        p = make()
        r = rh.RefDocp(p)
        try:
            want = r.solve()
        finally:
            r.close()
        c = rh.RefDocp(p, cuda=True)
        try:
            assert rh.lib().ref_set_int(b"prg_cuda_grad", grad) == 0
            got = c.solve()
        finally:
            c.close()
        assert want["result"] == "optimal" and got["result"] == "optimal"
        assert got["sqp_iters"] == want["sqp_iters"] and got["qp_iters"] == want["qp_iters"]
        assert abs(got["objective"] - want["objective"]) <= 1e-9 * abs(want["objective"])
        assert np.max(np.abs(got["x"] - want["x"])) <= 1e-7
    assert abs(want["objective"] - 14.814107952743184) < 1e-9  # (the CPU run of this container)


def _run_stack(which, K, nx, nu, grad=1):
    import json
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "sqp_stack_runner.py"), which, str(K), str(nx),
                          str(nu), str(grad)], capture_output=True, text=True, timeout=900)
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert lines, f"sqp_stack_runner {which} failed: rc={out.returncode}\n{out.stdout[-2000:]}\n{out.stderr[-2000:]}"
    return json.loads(lines[-1])


@pytest.mark.gpu
@pytest.mark.skipif(not rh.available(), reason="oracle/_ref not built")
@pytest.mark.parametrize("K,nx,nu,grad", [(20, 4, 2, 0), (40, 6, 3, 0), (100, 8, 4, 0), (40, 6, 3, 1)])
def test_gpu_sqp_with_every_device_module(K, nx, nu, grad):
    """The SQP-driven workload (config 5's kind, small): the unmodified Hqp_SqpPowell runs the
    nonlinear DOCP once with the reference's modules only (Prg_SynthNL, Hqp_HL_BFGS,
    Hqp_IpsMehrotra + Hqp_IpLQDOCP) and once with every device module of this repository
    plugged in -- program on Hqp_DocpCuda<> (row f4), sqp_hela CudaBFGS (row f2), sqp_qp_solver
    CudaMehrotra (device-resident IP solver + KKT engine).  The programs are ones on which the
    reference's own solver variants agree (LQDOCP / RedSpBKP / Franke: same SQP iterations,
    objective to 1e-13); on less well-posed instances (K = 60, nx = 4: 114 SQP iterations) the
    reference's variants already end 0.4 % apart and no parity can be asked for.
    grad 0 (the reference's forward differences): identical SQP and IP iteration counts;
    grad 1 (dual numbers): derivatives differ by the truncation error of the differences."""
    want, got = _run_stack("ref", K, nx, nu), _run_stack("cuda", K, nx, nu, grad)
    print({k: (want[k], got[k]) for k in ("result", "objective", "sqp_iters", "qp_iters")})
    assert want["result"] == "optimal" and got["result"] == "optimal"
    if grad == 0:
        assert (got["sqp_iters"], got["qp_iters"]) == (want["sqp_iters"], want["qp_iters"])
        assert abs(got["objective"] - want["objective"]) <= 1e-8 * abs(want["objective"])
        assert np.max(np.abs(np.array(got["x"]) - np.array(want["x"]))) <= 1e-6
    else:
        assert abs(got["objective"] - want["objective"]) <= 1e-5 * abs(want["objective"])
        assert abs(got["sqp_iters"] - want["sqp_iters"]) <= 3


@pytest.mark.gpu
@pytest.mark.parametrize("make,world", [(lambda: dc.synthnl_problem(11, 5, 2, 3, 2, seed=9), 3),
                                        (lambda: dc.synthnl_problem(300, 20, 10, 1, 1, seed=2), 4),
                                        (lambda: dc.did_problem(60, True), 2)])
def test_gpu_update_over_stage_ranges_equals_full_horizon(make, world):
    """One handle per stage range (hqpdocp_dims.k_first / K_total: one per GPU on a split horizon;
    here all on device 0): the ranges' results, concatenated group by group, are bit-identical
    to the full-horizon handle's; the objective is the sum of the ranges'."""
    p = make()
    x = p.x_init + 0.1 * np.random.default_rng(0).uniform(-1, 1, p.N)
    e = dc.DocpCuda(p)
    try:
        full = {m: e.update(x, m) for m in (dc.GRAD_FD, dc.GRAD_AD)}
        full_fbd = e.update_fbd(x)
    finally:
        e.close()
    shards = [p.shard(r, world) for r in range(world)]
    outs, fbd = {dc.GRAD_FD: [], dc.GRAD_AD: []}, []
    for q in shards:
        eq = dc.DocpCuda(q)
        try:
            xs = p.x_slice(x, q.k_first, q.k_first + q.K)
            for m in outs:
                outs[m].append(eq.update(xs, m))
            fbd.append(eq.update_fbd(xs))
        finally:
            eq.close()
    for m in outs:
        asm = dc.assemble_shards(p, shards, outs[m])
        for key in ("b", "d", "g", "fx", "fu", "cx", "cu"):
            assert np.array_equal(asm[key], full[m][key]), (m, key)
        assert abs(asm["f"] - full[m]["f"]) <= 1e-13 * max(1.0, abs(full[m]["f"]))
    a2 = dc.assemble_shards(p, shards, fbd)
    assert np.array_equal(a2["b"], full_fbd[1]) and np.array_equal(a2["d"], full_fbd[2])
