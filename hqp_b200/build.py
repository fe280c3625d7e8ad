"""In-tree build of the native pieces (no JIT cache: the .so files travel to the
GPU box with the repository snapshot).

  hqp_b200/lib/libhqpcuda.so   CUDA kernels + C ABI  (nvcc, sm_100a)
  hqp_b200/lib/libhqphl.so     block-diagonal BFGS update, CUDA + C ABI (nvcc, sm_100a)
  hqp_b200/lib/libhqpdocp.so   stage loop of Hqp_Docp::update for device models (nvcc, sm_100a)
  hqp_b200/lib/libhqpsynth.so  seeded synthetic-workload generator (g++)
  hqp_b200/lib/libhqp_ipcuda_plugin.so   Hqp_IpCuda host module; only where the
                                reference headers exist (/root/reference)
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIB = os.path.join(HERE, "lib")
CSRC = os.path.join(HERE, "csrc")
REF = os.environ.get("HQP_REFERENCE", "/root/reference")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3",
              "-std=c++17", "--use_fast_math=false", "-Xcompiler", "-fPIC", "-shared"]


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _run(cmd, **kw):
    r = subprocess.run(cmd, capture_output=True, text=True, **kw)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        raise RuntimeError(f"build step failed: {cmd[0]}")
    return r


def build_cuda(force=False, verbose=False):
    os.makedirs(LIB, exist_ok=True)
    out = os.path.join(LIB, "libhqpcuda.so")
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC)
            if f.endswith((".cu", ".cuh", ".inc")) and f not in ("hl_bfgs.cu", "docp_update.cu", "docp_models.cuh")]  # (own libraries)
    srcs.append(os.path.join(ROOT, "include", "hqp_ipcuda.h"))
    if force or _newer(out, srcs):
        flags = [f for f in NVCC_FLAGS if f != "--use_fast_math=false"]
        cmd = ["nvcc", *flags, "-I" + os.path.join(ROOT, "include")]
        if verbose:
            cmd += ["-Xptxas", "-v"]
        cmd += [os.path.join(CSRC, "hqp_ipcuda.cu"), "-o", out, "-lcudart"]
        r = _run(cmd)
        if verbose:
            print(r.stderr)
    return out


def build_hl(force=False):
    """libhqphl.so: block-diagonal BFGS update on the GPU (include/hqp_hlcuda.h)."""
    os.makedirs(LIB, exist_ok=True)
    out = os.path.join(LIB, "libhqphl.so")
    src = os.path.join(CSRC, "hl_bfgs.cu")
    if force or _newer(out, [src, os.path.join(ROOT, "include", "hqp_hlcuda.h")]):
        flags = [f for f in NVCC_FLAGS if f != "--use_fast_math=false"]
        _run(["nvcc", *flags, "-I" + os.path.join(ROOT, "include"), src, "-o", out, "-lcudart"])
    return out


def build_docp(force=False):
    """libhqpdocp.so: the stage loop of Hqp_Docp::update on the GPU (include/hqp_docpcuda.h).
    -fmad=false: the model code must round like its CPU restatements (docp_models.cuh)."""
    os.makedirs(LIB, exist_ok=True)
    out = os.path.join(LIB, "libhqpdocp.so")
    src = os.path.join(CSRC, "docp_update.cu")
    if force or _newer(out, [src, os.path.join(CSRC, "docp_models.cuh"),
                             os.path.join(ROOT, "include", "hqp_docpcuda.h")]):
        flags = [f for f in NVCC_FLAGS if f != "--use_fast_math=false"]
        _run(["nvcc", *flags, "-fmad=false", "-I" + os.path.join(ROOT, "include"), src, "-o", out, "-lcudart"])
    return out


def build_synth(force=False):
    os.makedirs(LIB, exist_ok=True)
    out = os.path.join(LIB, "libhqpsynth.so")
    src = os.path.join(CSRC, "synth.cpp")
    if force or _newer(out, [src]):
        _run(["g++", "-O2", "-fPIC", "-shared", "-std=c++11", src, "-o", out])
    return out


def build_plugin(force=False):
    """Hqp_IpCuda : Hqp_IpMatrix.  Needs the reference headers; the resulting
    .so has no link-time dependency on the reference (symbols resolve in the
    host process that loads it, like any HQP plugin)."""
    src = os.path.join(HERE, "host", "Hqp_IpCuda.C")
    hdr = os.path.join(HERE, "host", "Hqp_IpCuda.h")
    out = os.path.join(LIB, "libhqp_ipcuda_plugin.so")
    if not os.path.isdir(os.path.join(REF, "hqp")) or not os.path.exists(src):
        return out if os.path.exists(out) else None
    hl_src = os.path.join(HERE, "host", "Hqp_HL_CudaBFGS.C")
    did_src = os.path.join(HERE, "host", "Prg_DIDCuda.C")
    if force or _newer(out, [src, hdr, os.path.join(HERE, "host", "Hqp_IpsCuda.C"),
                             os.path.join(HERE, "host", "Hqp_IpsCuda.h"), hl_src,
                             os.path.join(HERE, "host", "Hqp_HL_CudaBFGS.h"), did_src,
                             os.path.join(HERE, "host", "Hqp_DocpCuda.h"),
                             os.path.join(ROOT, "include", "hqp_docpcuda.h"),
                             os.path.join(ROOT, "include", "hqp_hlcuda.h"),
                             os.path.join(ROOT, "include", "hqp_ipcuda.h")]):
        build_hl()
        build_docp()
        shim = os.path.join(ROOT, "oracle", "tclshim")
        _run(["g++", "-std=c++11", "-O2", "-fPIC", "-shared", "-w", "-fpermissive",
              f"-I{REF}", f"-I{shim}", f"-I{REF}/iftcl", f"-I{REF}/hqp", f"-I{REF}/hqp_docp",
              "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(HERE, "host"), src,
              os.path.join(HERE, "host", "Hqp_IpsCuda.C"), hl_src, did_src, "-o", out,
              "-L" + LIB, "-lhqpcuda", "-lhqphl", "-lhqpdocp", "-Wl,-rpath,$ORIGIN"])
    return out


def build_all(force=False):
    build_synth(force)
    build_cuda(force)
    build_hl(force)
    build_docp(force)
    build_plugin(force)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv)
    print("built:", sorted(os.listdir(LIB)))
