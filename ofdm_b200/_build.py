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
    "-Xcompiler", "-fPIC", "-diag-suppress", "177",
]
# one translation unit per kernel family (csrc/kernels.h) + the C ABI; compiled in parallel, linked into one library
UNITS = ["rx64_m2", "wide_rx_m2", "rx64_m0", "rx64_m1", "wide_rx_m0", "wide_rx_m1", "wide_tx", "wide_txr", "tx64", "tx64r", "tx64w", "rx64", "wide_rx",
         "rs", "sync", "ofdm_engine"]
OBJ_DIR = os.path.join(_HERE, "build")


def _sources():
    out = [os.path.join(ROOT, "include", "ofdm_engine.h")]
    for f in sorted(os.listdir(CSRC)):
        if f.endswith((".cu", ".cuh", ".h", ".inc")):
            out.append(os.path.join(CSRC, f))
    return out


def engine_is_stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(s) > t for s in _sources())


def _unit_deps(unit: str):
    """Files a unit really includes: its .cu and, recursively, every quoted #include that lives in csrc/ or include/."""
    import re
    seen, todo = set(), [os.path.join(CSRC, unit + ".cu")]
    while todo:
        f = os.path.normpath(todo.pop())
        if f in seen or not os.path.exists(f):
            continue
        seen.add(f)
        with open(f, "r", encoding="utf-8", errors="replace") as fh:
            for inc in re.findall(r'^\s*#\s*include\s+"([^"]+)"', fh.read(), flags=re.M):
                todo.append(os.path.join(os.path.dirname(f), inc))
    return sorted(seen)


def build_engine(force: bool = False, verbose: bool = False, units=None) -> str:
    """nvcc -gencode arch=compute_100a,code=sm_100a ... -> ofdm_b200/libofdm_b200.so (cross-compiles without a GPU).

    Every unit of UNITS is compiled to ofdm_b200/build/<unit>.o by its own nvcc process (in parallel), then linked."""
    if not force and not engine_is_stale():
        return LIB_PATH
    from concurrent.futures import ThreadPoolExecutor
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    os.makedirs(OBJ_DIR, exist_ok=True)

    def compile_unit(unit):
        obj = os.path.join(OBJ_DIR, unit + ".o")
        if not force and os.path.exists(obj) and all(os.path.getmtime(obj) >= os.path.getmtime(d) for d in _unit_deps(unit)):
            return unit, 0, ""
        cmd = [nvcc, *NVCC_FLAGS, "-ccbin", "/usr/bin/g++", "-c", "-o", obj, os.path.join(CSRC, unit + ".cu")]
        if os.environ.get("OFDM_NVCC_EXTRA"):
            cmd += os.environ["OFDM_NVCC_EXTRA"].split()
        if verbose:
            cmd += ["-Xptxas", "-v"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        return unit, r.returncode, r.stdout + r.stderr

    with ThreadPoolExecutor(max_workers=max(1, min(len(UNITS), os.cpu_count() or 1))) as ex:
        results = list(ex.map(compile_unit, units or UNITS))
    for unit, rc, log in results:
        if rc != 0:
            raise RuntimeError(f"nvcc failed on {unit}.cu:\n{log}")
        if verbose and log:
            print(f"==== {unit}.cu\n{log}")
    objs = [os.path.join(OBJ_DIR, u + ".o") for u in UNITS]
    r = subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-ccbin", "/usr/bin/g++", "-o", LIB_PATH, *objs, "-ldl", "-lpthread"],
                       capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
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
