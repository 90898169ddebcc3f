#!/usr/bin/env python3
"""Builds a differently configured copy of the kernel library for a tuning sweep:

    python benchmarks/build_variant.py NAME -DEQS_CE_DIMS=8 -DEQS_PSO_UQ=3 ...
    EQS_B200_LIBRARY=eqsolver_b200/lib/variants/libeqsolver_b200_NAME.so python benchmarks/ce_bench.py

Same sources, same flags as eqsolver_b200/build.py plus the given defines; objects under build/variants/NAME/.  The
variants are measurement artefacts (profiles/r02_sweep.md), never what the package loads by default."""
import concurrent.futures as cf
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from eqsolver_b200 import build as B  # noqa: E402


def main():
    name, defines = sys.argv[1], sys.argv[2:]
    obj_dir = os.path.join(B.OBJ_DIR, "variants", name)
    lib_dir = os.path.join(B.LIB_DIR, "variants")
    os.makedirs(obj_dir, exist_ok=True)
    os.makedirs(lib_dir, exist_ok=True)
    nvcc = B._nvcc()

    def one(src):
        obj = os.path.join(obj_dir, src.replace(".cu", ".o"))
        subprocess.run([nvcc, *B.NVCC_FLAGS, *defines, "-c", os.path.join(B.CSRC, src), "-o", obj], check=True)
        return obj

    with cf.ThreadPoolExecutor(max_workers=len(B.SOURCES)) as ex:
        objs = list(ex.map(one, B.SOURCES))
    out = os.path.join(lib_dir, f"libeqsolver_b200_{name}.so")
    subprocess.run([nvcc, "-shared", "-o", out, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-ldl"], check=True)
    print(out)


if __name__ == "__main__":
    main()
