"""In-tree build of the CUDA library (sm_100a).

``python -m acvd_b200.build`` compiles ``acvd_b200/libacvd_b200.so`` with nvcc; nvcc cross-compiles
without a GPU, and the built ``.so`` travels to the GPU box with the repository snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libacvd_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
]
SOURCES = ["acvd_capi.cu"]


def _newest_source_mtime() -> float:
    m = 0.0
    for d in (CSRC, os.path.join(ROOT, "include")):
        for f in os.listdir(d):
            m = max(m, os.path.getmtime(os.path.join(d, f)))
    return m


def find_nvcc() -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: the acvd_b200 CUDA library cannot be built")
    return nvcc


def build_library(force: bool = False, verbose: bool = False) -> str:
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= _newest_source_mtime():
        return LIB
    cmd = [find_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB]
    cmd += [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    cmd += ["-ldl"]   # NCCL is resolved with dlopen at run time (csrc/nccl_dyn.hpp)
    subprocess.check_call(cmd, cwd=CSRC)
    return LIB


HOST = os.path.join(CSRC, "host")
BIN = os.path.join(HERE, "bin")
CLIS = ["ACVD", "ACVDQ", "AnisotropicRemeshingQ"]


def build_host(force: bool = False) -> list:
    """Host C++ front-ends (VTK-free vtkSurface + remeshing classes + the three CLIs), linked against
    libacvd_b200.so through its C ABI only."""
    os.makedirs(BIN, exist_ok=True)
    build_library()
    srcs = [os.path.join(HOST, "vtkSurface.cpp"), os.path.join(HOST, "vtkDiscreteRemeshing.cpp")]
    newest = max(os.path.getmtime(os.path.join(dp, f)) for dp, _, fs in os.walk(HOST) for f in fs)
    newest = max(newest, os.path.getmtime(os.path.join(ROOT, "include", "acvd_b200.h")))
    out = []
    for cli in CLIS:
        exe = os.path.join(BIN, cli)
        out.append(exe)
        if not force and os.path.exists(exe) and os.path.getmtime(exe) >= newest:
            continue
        cmd = ["g++", "-std=c++17", "-O2", "-o", exe, os.path.join(HOST, "Examples", cli + ".cxx")] + srcs
        cmd += ["-L" + HERE, "-lacvd_b200", "-Wl,-rpath,$ORIGIN/.."]
        subprocess.check_call(cmd)
    return out


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
    print(build_host(force="--force" in sys.argv))
