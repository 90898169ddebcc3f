"""Build recipe of the kernel library: nvcc for sm_100a only, in-tree output (eqsolver_b200/lib/).

    python -m eqsolver_b200.build            # incremental
    python -m eqsolver_b200.build --force

`EQS_USER_OBJECTIVE_HEADER=/path/mine.cuh` swaps in a user objective (csrc/user_objective.cuh documents the
contract).  There is exactly one target architecture and no fallback path.
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(ROOT, "csrc")
OBJ_DIR = os.path.join(os.path.dirname(ROOT), "build")
LIB_DIR = os.path.join(ROOT, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libeqsolver_b200.so")

SOURCES = ["api.cu", "pso.cu", "ce.cu", "mc.cu", "comm.cu"]
# -fmad=false: rustc never contracts a*b+c, and replay parity with the CPU restatement is bit-level.
NVCC_FLAGS = ["-std=c++20", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-fmad=false",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden"]


def _nvcc() -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: the kernel library can only be built with the CUDA toolkit")
    return nvcc


def _deps_mtime() -> float:
    newest = 0.0
    for d in (CSRC, os.path.join(os.path.dirname(ROOT), "include")):
        for f in os.listdir(d):
            newest = max(newest, os.path.getmtime(os.path.join(d, f)))
    user = os.environ.get("EQS_USER_OBJECTIVE_HEADER")
    if user:
        newest = max(newest, os.path.getmtime(user))
    return newest


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and os.path.exists(LIB_PATH) and os.path.getmtime(LIB_PATH) >= _deps_mtime():
        return LIB_PATH
    nvcc = _nvcc()
    os.makedirs(OBJ_DIR, exist_ok=True)
    os.makedirs(LIB_DIR, exist_ok=True)
    flags = list(NVCC_FLAGS)
    user = os.environ.get("EQS_USER_OBJECTIVE_HEADER")
    if user:
        # objectives.cuh includes this file instead of csrc/user_objective.cuh
        flags += [f'-DEQS_USER_OBJECTIVE_HEADER="{os.path.abspath(user)}"']
    flags += os.environ.get("EQS_NVCC_DEFINES", "").split()  # e.g. "-DEQS_PSO_UQ=3" for tuning sweeps
    if verbose:
        flags += ["-Xptxas", "-v"]

    def compile_one(src: str) -> str:
        obj = os.path.join(OBJ_DIR, src.replace(".cu", ".o"))
        cmd = [nvcc, *flags, "-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with cf.ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    cmd = [nvcc, "-shared", "-o", LIB_PATH, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-ldl"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
