"""Builds scalce_b200/libscalce_b200.so (hand-written sm_100a CUDA + the C ABI) in-tree with nvcc."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libscalce_b200.so")
SOURCES = ["api.cu", "core_table.cpp"]
# every header under csrc/ is a dependency of the one translation unit (api.cu includes them all)
HEADERS = sorted(f for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))) + [os.path.join("..", "..", "include", "scalce_b200.h")]
NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC,-O2,-Wall,-Wno-unused-function", "-shared", "--expt-relaxed-constexpr",
]


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [__file__]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build_lib(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("nvcc failed")
    if verbose:
        print(r.stdout)
    return LIB


NCCL_LIB = os.path.join(HERE, "libscalce_b200_nccl.so")
NCCL_SRC = os.path.join(CSRC, "comm_nccl.cpp")


def build_nccl_comm(force: bool = False) -> str:
    """nvcc -> scalce_b200/libscalce_b200_nccl.so: an ncclComm_t wrapped into the scb_comm of scb_shard_flush (csrc/comm_nccl.cpp).
    Separate from the main library so that the transform itself carries no NCCL dependency."""
    deps = [NCCL_SRC, os.path.join(HERE, "..", "include", "scalce_b200.h")]
    if not force and os.path.exists(NCCL_LIB) and all(os.path.getmtime(d) <= os.path.getmtime(NCCL_LIB) for d in deps):
        return NCCL_LIB
    nvcc = os.environ.get("NVCC", "nvcc")
    inc, libdirs = [], []
    try:   # prefer the NCCL that ships with torch (the one the GPU box loads); fall back to the system one
        import nvidia.nccl as _n
        base = os.path.dirname(_n.__file__) if getattr(_n, "__file__", None) else list(_n.__path__)[0]
        inc, libdirs = ["-I" + os.path.join(base, "include")], [os.path.join(base, "lib")]
    except Exception:
        pass
    cmd = [nvcc, "-O2", "-std=c++17", "-Xcompiler", "-fPIC,-Wall", "-shared", "-o", NCCL_LIB, NCCL_SRC] + inc
    for d in libdirs:
        cmd += ["-L" + d, "-Xlinker", "-rpath," + d]
    cmd += ["-l:libnccl.so.2"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("libscalce_b200_nccl.so build failed")
    return NCCL_LIB


HOST_SRC = os.path.join(HERE, "host", "scb_boost.cpp")
HOST_BIN = os.path.join(HERE, "host", "scb_boost")


def build_host_tool(force: bool = False) -> str:
    """g++ -> scalce_b200/host/scb_boost: the C++ host side (FASTQ -> C ABI -> the reference's temp files), linked
    against libscalce_b200.so through an rpath relative to the binary."""
    build_lib()
    deps = [HOST_SRC, LIB, os.path.join(HERE, "..", "include", "scalce_b200.h")]
    if not force and os.path.exists(HOST_BIN) and all(os.path.getmtime(d) <= os.path.getmtime(HOST_BIN) for d in deps):
        return HOST_BIN
    cxx = os.environ.get("CXX", "g++")
    cmd = [cxx, "-O2", "-std=c++17", "-Wall", "-o", HOST_BIN, HOST_SRC, "-L" + HERE, "-lscalce_b200", "-Wl,-rpath,$ORIGIN/.."]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("host tool build failed")
    return HOST_BIN


if __name__ == "__main__":
    print(build_lib(force="--force" in sys.argv, verbose="-v" in sys.argv))
    print(build_host_tool(force="--force" in sys.argv))
    print(build_nccl_comm(force="--force" in sys.argv))
