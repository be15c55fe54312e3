"""In-tree build of the native libraries (nvcc -> sm_100a, g++ for the host side).

`python -m dsopp_b200.build` or `__graft_entry__.build()`.  Outputs go to dsopp_b200/lib/ (git-ignored,
shipped to the GPU box by gpurun).
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "lib")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ARCH + ["-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden"]

CUDA_LIB = os.path.join(LIB, "libdsopp_pba_cuda.so")
HOST_LIB = os.path.join(LIB, "libdsopp_pba_host.so")


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _run(cmd):
    print("+", " ".join(cmd), flush=True)
    subprocess.check_call(cmd)


def build_cuda(force=False, verbose_ptxas=False):
    os.makedirs(LIB, exist_ok=True)
    srcs = [os.path.join(CSRC, "pba_kernels.cu"), os.path.join(CSRC, "pba_capi.cu"), os.path.join(CSRC, "pose_alignment.cu"),
            os.path.join(CSRC, "peer_exchange.cu"), os.path.join(CSRC, "energy_quantile.cu"),
            os.path.join(CSRC, "depth_maps.cu"), os.path.join(CSRC, "image_tma.cu")]
    deps = srcs + [os.path.join(CSRC, "pba_internal.h"), os.path.join(CSRC, "depth_maps_body.h"),
                   os.path.join(CSRC, "energy_quantile_body.h"), os.path.join(CSRC, "peer_exchange_body.h"), os.path.join(ROOT, "include", "dsopp_cuda_pba.h"),
                   os.path.join(ROOT, "include", "dsopp_cuda_pose_alignment.h")]
    if not force and not _newer(CUDA_LIB, deps):
        return CUDA_LIB
    objs = []
    for s in srcs:
        o = os.path.join(LIB, os.path.basename(s) + ".o")
        flags = list(NVCC_FLAGS)
        if verbose_ptxas:
            flags += ["-Xptxas", "-v"]
        _run([NVCC] + flags + ["-I", os.path.join(ROOT, "include"), "-c", s, "-o", o])
        objs.append(o)
    _run([NVCC] + ARCH + ["-shared", "-o", CUDA_LIB] + objs + ["-ldl"])
    return CUDA_LIB


def build_host(force=False):
    """C++ host side above the C ABI (LM driver, NormalLinearSystem, solver class)."""
    src = os.path.join(CSRC, "host", "host_capi.cpp")
    if not os.path.exists(src):
        return None
    hdrs = [os.path.join(CSRC, "host", f) for f in os.listdir(os.path.join(CSRC, "host"))]
    if not force and not _newer(HOST_LIB, hdrs + [CUDA_LIB]):
        return HOST_LIB
    _run(["g++", "-O2", "-std=c++20", "-fPIC", "-shared", "-Wall", "-Wextra", "-I", os.path.join(ROOT, "include"),
          "-o", HOST_LIB, src, "-L", LIB, "-ldsopp_pba_cuda", "-Wl,-rpath,$ORIGIN"])
    return HOST_LIB


def build_all(force=False):
    build_cuda(force)
    build_host(force)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv)
