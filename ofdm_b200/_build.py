"""Build the engine's shared library in-tree with nvcc (sm_100a only)."""
from __future__ import annotations

import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(_HERE)
CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.path.join(_HERE, "libofdm_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared", "-diag-suppress", "177",
]


def _sources():
    out = [os.path.join(ROOT, "include", "ofdm_engine.h")]
    for f in sorted(os.listdir(CSRC)):
        if f.endswith((".cu", ".cuh", ".h")):
            out.append(os.path.join(CSRC, f))
    return out


def engine_is_stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(s) > t for s in _sources())


def build_engine(force: bool = False, verbose: bool = False) -> str:
    """nvcc -gencode arch=compute_100a,code=sm_100a ... -> ofdm_b200/libofdm_b200.so (cross-compiles without a GPU)."""
    if not force and not engine_is_stale():
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc, *NVCC_FLAGS, "-ccbin", "/usr/bin/g++", "-o", LIB_PATH, os.path.join(CSRC, "ofdm_engine.cu")]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stderr)
    return LIB_PATH


HOST_DIR = os.path.join(_HERE, "host")
LAB3C = os.path.join(HOST_DIR, "lab3c")


def build_host_example(force: bool = False) -> str:
    """g++ the C++ host mirror's lab3c example against libofdm_b200.so (rpath $ORIGIN/..)."""
    srcs = [os.path.join(HOST_DIR, "lab3c.cpp"), os.path.join(HOST_DIR, "ofdm.hpp"), LIB_PATH]
    if not force and os.path.exists(LAB3C) and all(os.path.getmtime(LAB3C) >= os.path.getmtime(x) for x in srcs):
        return LAB3C
    cmd = ["/usr/bin/g++", "-O2", "-std=c++17", "-Wall", "-o", LAB3C, srcs[0], "-L" + _HERE, "-lofdm_b200", "-Wl,-rpath,$ORIGIN/.."]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("g++ failed:\n" + r.stdout + r.stderr)
    return LAB3C
