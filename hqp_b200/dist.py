"""Horizon split across ranks: one process per GPU, contiguous
stage ranges, `torch.distributed` for the plumbing.

Per factor every rank condenses its range to one boundary element (A, C, J) --
its Schur complement onto (state at the range start, costate at its end) --
and ONE all-gather of 4*nx*nx doubles makes every rank able to compute the value
Hessian at its own range end (replicated chain over the ranks behind it, at most
world-1 steps); a second all-gather (nx*nx) publishes the closed-loop transition
of each range for the solves.  Per solve two all-gathers of 2*nx doubles (the
backward and the forward boundary vectors).  No other data crosses NVLink.

The protocol is written once (`RangeSolver`) over an engine interface with two
implementations: `CudaRangeEngine` (the CUDA C ABI, device tensors, NCCL) and,
in tests/, a numpy engine used to exercise the protocol on CPU with gloo.
"""
from __future__ import annotations

import ctypes
import dataclasses

import numpy as np

from .problem import LQProblem


# --------------------------------------------------------------------------
# partition of a global problem / its vectors into stage ranges
# --------------------------------------------------------------------------
@dataclasses.dataclass
class RangeMap:
    rank: int
    world: int
    k0: int
    k1: int
    x_sl: slice            # global x entries owned by the range (K_loc*nm [+nx on the last])
    dyn_sl: slice          # global dynamics rows of the range
    x0_rows: slice | None  # global rows fixing x0 (rank 0 only)
    ineq_rows: np.ndarray  # global inequality rows whose stage lies in the range

    @property
    def last(self):
        return self.rank == self.world - 1


def stage_ranges(K, world):
    base, rem = divmod(K, world)
    out, k = [], 0
    for r in range(world):
        n = base + (1 if r < rem else 0)
        out.append((k, k + n))
        k += n
    return out


def split_problem(prob: LQProblem, world: int):
    """-> [(local LQProblem, RangeMap)] ; general equality rows are not supported."""
    if prob.n_eq:
        raise NotImplementedError("horizon split with general equality rows")
    nx, nu, nm, K = prob.nx, prob.nu, prob.nm, prob.K
    stage, lcol = prob.ineq_stage_local()
    out = []
    for r, (k0, k1) in enumerate(stage_ranges(K, world)):
        last = r == world - 1
        Kl = k1 - k0
        Q = np.zeros((Kl + 1, nm, nm))
        Q[:Kl] = prob.Q[k0:k1]
        if last:
            Q[Kl] = prob.Q[K]
        fixed = prob.fixed_x0 and r == 0
        b = prob.b[k0 * nx:k1 * nx]
        if fixed:
            b = np.concatenate([b, prob.b[K * nx:K * nx + nx]])
        c = np.zeros(Kl * nm + nx)
        c[:Kl * nm] = prob.c[k0 * nm:k1 * nm]
        if last:
            c[Kl * nm:] = prob.c[K * nm:K * nm + nx]
        hi = k1 + (1 if last else 0)
        rows = np.nonzero((stage >= k0) & (stage < hi))[0]
        ptr = [0]
        col, val = [], []
        for i in rows:
            e0, e1 = prob.ineq_ptr[i], prob.ineq_ptr[i + 1]
            col.extend((prob.ineq_col[e0:e1] - k0 * nm).tolist())
            val.extend(prob.ineq_val[e0:e1].tolist())
            ptr.append(len(col))
        lp = LQProblem(nx, nu, Kl, Q, c, np.ascontiguousarray(prob.fx[k0:k1]),
                       np.ascontiguousarray(prob.fu[k0:k1]), np.ascontiguousarray(b),
                       fixed_x0=fixed, ineq_ptr=np.asarray(ptr, np.int32),
                       ineq_col=np.asarray(col, np.int32), ineq_val=np.asarray(val, np.float64),
                       d=prob.d[rows] if prob.m else np.zeros(0))
        rm = RangeMap(r, world, k0, k1, slice(k0 * nm, k1 * nm + (nx if last else 0)),
                      slice(k0 * nx, k1 * nx),
                      slice(K * nx, K * nx + nx) if fixed else None, rows)
        out.append((lp, rm))
    return out


def local_vectors(prob: LQProblem, rm: RangeMap, z, w, r1, r2, r3, r4):
    """global IP vectors -> the range's local ones (local layout of LQProblem)."""
    nx = prob.nx
    n_loc = (rm.k1 - rm.k0) * prob.nm + nx
    l1 = np.zeros(n_loc)
    seg = r1[rm.x_sl]
    l1[:len(seg)] = seg
    l2 = r2[rm.dyn_sl]
    if rm.x0_rows is not None:
        l2 = np.concatenate([l2, r2[rm.x0_rows]])
    rows = rm.ineq_rows
    return (np.ascontiguousarray(z[rows]), np.ascontiguousarray(w[rows]), l1,
            np.ascontiguousarray(l2), np.ascontiguousarray(r3[rows]),
            np.ascontiguousarray(r4[rows]))


def scatter_solution(prob: LQProblem, rm: RangeMap, local, dx, dy, dz, dw):
    """write the range's local solution into the global vectors"""
    ldx, ldy, ldz, ldw = local
    n = rm.x_sl.stop - rm.x_sl.start
    dx[rm.x_sl] = ldx[:n]
    nd = rm.dyn_sl.stop - rm.dyn_sl.start
    dy[rm.dyn_sl] = ldy[:nd]
    if rm.x0_rows is not None:
        dy[rm.x0_rows] = ldy[nd:nd + prob.nx]
    dz[rm.ineq_rows] = ldz
    dw[rm.ineq_rows] = ldw


# --------------------------------------------------------------------------
# the exchange protocol
# --------------------------------------------------------------------------
class RangeSolver:
    """factor / step of one rank's stage range inside a torch.distributed group.

    `engine` implements factor_begin / factor_finish / step_begin / step_mid /
    step_finish on torch tensors that live where the process group needs them
    (CUDA for nccl, CPU for gloo)."""

    def __init__(self, engine, rank, world, group=None):
        self.e, self.rank, self.world, self.group = engine, rank, world, group
        self.gpsi = None

    def _gather(self, t):
        import torch
        import torch.distributed as dist
        flat = t.contiguous().view(-1)
        out = torch.empty(self.world * flat.numel(), dtype=t.dtype, device=t.device)
        if self.world == 1:
            out.copy_(flat)
        else:
            dist.all_gather_into_tensor(out, flat, group=self.group)
        return out.view((self.world,) + tuple(t.shape))

    def _gather_async(self, t):
        """all-gather that is only waited for at its first use (overlaps with the
        kernels enqueued meanwhile); -> (output, work or None)"""
        import torch
        import torch.distributed as dist
        flat = t.contiguous().view(-1)
        out = torch.empty(self.world * flat.numel(), dtype=t.dtype, device=t.device)
        if self.world == 1:
            out.copy_(flat)
            return out.view((self.world,) + tuple(t.shape)), None
        work = dist.all_gather_into_tensor(out, flat, group=self.group, async_op=True)
        return out.view((self.world,) + tuple(t.shape)), work

    def factor(self, z, w):
        xf = self.e.factor_begin(z, w)
        gathered = self._gather(xf)
        xpsi = self.e.factor_finish(gathered, self.rank, self.world)
        # the range transitions are first needed in the middle of the next step
        self.gpsi, self._gpsi_work = self._gather_async(xpsi)

    def step(self, r1, r2, r3, r4):
        xv = self.e.step_begin(r1, r2, r3, r4)
        gv = self._gather(xv)
        if getattr(self, "_gpsi_work", None) is not None:
            self._gpsi_work.wait()
            self._gpsi_work = None
        xx = self.e.step_mid(gv, self.gpsi, self.rank, self.world)
        gx = self._gather(xx)
        return self.e.step_finish(gx, self.gpsi, self.rank, self.world)


class CudaRangeEngine:
    """the CUDA C ABI (hqpcu_range_*) on device tensors of the current stream"""

    def __init__(self, local_prob: LQProblem, rm: RangeMap, device=0, nseg=0):
        import torch
        from . import ipcuda
        self.torch, self.ic = torch, ipcuda
        self.dev = torch.device("cuda", device)
        self.p, self.rm = local_prob, rm
        self.eng = ipcuda.IpCuda(local_prob, device=device, nseg=nseg)
        self.eng.set_stream(torch.cuda.current_stream(self.dev).cuda_stream)
        ipcuda._check(ipcuda.lib().hqpcu_range_config(self.eng.h, int(rm.rank > 0),
                                                      int(not rm.last)), "hqpcu_range_config")
        self.eng.update()
        nx = local_prob.nx
        f64 = dict(dtype=torch.float64, device=self.dev)
        self.xf = torch.zeros(4 * nx * nx, **f64)
        self.xpsi = torch.zeros(nx * nx, **f64)
        self.xv = torch.zeros(2 * nx, **f64)
        self.xx = torch.zeros(2 * nx, **f64)
        m = max(local_prob.m, 1)
        self.out = [torch.zeros(n, **f64) for n in (local_prob.N, local_prob.me, m, m)]

    def _p(self, t):
        return ctypes.c_void_p(t.data_ptr())

    def _ck(self, rc, what):
        self.ic._check(rc, what)

    def factor_begin(self, z, w):
        self._ck(self.ic.lib().hqpcu_range_factor_begin(self.eng.h, self._p(z), self._p(w),
                                                        self._p(self.xf)), "range_factor_begin")
        return self.xf

    def factor_finish(self, gathered, rank, world):
        self._ck(self.ic.lib().hqpcu_range_factor_finish(self.eng.h, self._p(gathered), rank, world,
                                                         self._p(self.xpsi)), "range_factor_finish")
        return self.xpsi

    def step_begin(self, r1, r2, r3, r4):
        self._keep = (r1, r2, r3, r4)
        self._ck(self.ic.lib().hqpcu_range_step_begin(self.eng.h, self._p(r1), self._p(r2),
                                                      self._p(r3), self._p(r4), self._p(self.xv)),
                 "range_step_begin")
        return self.xv

    def step_mid(self, gv, gpsi, rank, world):
        self._ck(self.ic.lib().hqpcu_range_step_mid(self.eng.h, self._p(gv), self._p(gpsi), rank,
                                                    world, self._p(self.xx)), "range_step_mid")
        return self.xx

    def step_finish(self, gx, gpsi, rank, world):
        o = self.out
        self._ck(self.ic.lib().hqpcu_range_step_finish(self.eng.h, self._p(gx), self._p(gpsi), rank,
                                                       world, *[self._p(t) for t in o]),
                 "range_step_finish")
        return o

    def status(self):
        return self.eng.sync_status()

    def close(self):
        self.eng.close()


# --------------------------------------------------------------------------
# the exchanges inside the library (hqpcu_comm_init): what bench.py and the
# multi-process GPU tests use
# --------------------------------------------------------------------------
def exchange_unique_id(rank, group=None):
    """NCCL unique id of a new library communicator: made on rank 0, broadcast over
    the existing torch.distributed group (nccl or gloo)."""
    import torch
    import torch.distributed as dist
    from . import ipcuda
    if dist.get_world_size(group) == 1:
        return None
    on_gpu = dist.get_backend(group) == "nccl"
    dev = torch.device("cuda", torch.cuda.current_device()) if on_gpu else torch.device("cpu")
    t = torch.zeros(128, dtype=torch.uint8, device=dev)
    if rank == 0:
        t.copy_(torch.frombuffer(bytearray(ipcuda.IpCuda.comm_unique_id()), dtype=torch.uint8))
    dist.broadcast(t, src=0, group=group)
    return bytes(t.cpu().numpy().tobytes())


class DistIpCuda:
    """One rank's stage range of a horizon split over the GPUs of the process group.

    Wraps an `IpCuda` handle created for the LOCAL problem and turned into rank
    `rank` of `world` by `hqpcu_comm_init`: afterwards the handle's own methods
    (update / factor / step / solve / residuum / mehrotra_solve and their `_dev`
    flavours) run the split algorithm on the local slices, NCCL calls included, and
    must be called by all ranks together."""

    def __init__(self, local_prob: LQProblem, rank, world, device=0, nseg=0, group=None):
        from . import ipcuda
        self.rank, self.world = rank, world
        self.eng = ipcuda.IpCuda(local_prob, device=device, nseg=nseg)
        uid = exchange_unique_id(rank, group) if world > 1 else None
        self.eng.comm_init(uid, rank, world)

    def __getattr__(self, name):
        return getattr(self.eng, name)

    def close(self):
        self.eng.close()
