"""The C++ header-only builder (include/eqsolver_b200.hpp) and the Rust FFI declarations (rust/, source only)."""
import os
import re
import subprocess

import pytest

from conftest import ROOT

BIN = os.path.join(ROOT, "build", "test_builder")


def build_cpp():
    from eqsolver_b200 import _ffi
    os.makedirs(os.path.dirname(BIN), exist_ok=True)
    libdir = os.path.dirname(_ffi.LIB_PATH)
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "test_builder.cpp"), "-o", BIN, "-L", libdir,
                           "-leqsolver_b200", f"-Wl,-rpath,{libdir}"])


def test_cpp_builder_host_side():
    build_cpp()
    r = subprocess.run([BIN, "host"], capture_output=True, text=True)
    assert r.returncode == 0 and "host tests ok" in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
def test_cpp_builder_on_gpu():
    build_cpp()
    r = subprocess.run([BIN, "gpu"], capture_output=True, text=True)
    assert r.returncode == 0 and "gpu tests ok" in r.stdout, r.stdout + r.stderr


def test_tile_cyclic_distribution_helpers_on_the_host():
    """csrc/pso.cuh: pso_gidx / pso_lidx / pso_lcount / pso_cyclic_shard, the index arithmetic of the sequential-exact mode
    over several ranks, as a host program (the functions are __host__ __device__): the ranks' index sets partition 0..N-1
    for N around the tile size and up to 2^18, worlds 1..9; local order is global order; window cuts count right."""
    from eqsolver_b200 import build as kbuild
    exe = os.path.join(ROOT, "build", "test_cyclic")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    subprocess.check_call([kbuild._nvcc(), "-std=c++20", "-O1", "-gencode", "arch=compute_100a,code=sm_100a", "-x", "cu",
                           os.path.join(ROOT, "tests", "cpp", "test_cyclic.cu"), "-o", exe])
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and "cyclic distribution ok" in r.stdout, r.stdout + r.stderr


def test_rust_extern_block_matches_header():
    """rust/eqsolver-cuda-sys cannot be compiled here (no cargo); at least its symbol set must equal the header's."""
    from test_abi import header_symbols
    src = open(os.path.join(ROOT, "rust", "eqsolver-cuda-sys", "src", "lib.rs")).read()
    rust = sorted(set(re.findall(r"pub fn (eqs_[a-z0-9_]+)\s*\(", src)))
    assert rust == header_symbols()
    # struct field order of the #[repr(C)] mirrors == the C structs
    hdr = open(os.path.join(ROOT, "include", "eqsolver_b200.h")).read()
    for name in ("eqs_pso_params", "eqs_ce_params", "eqs_mc_params", "eqs_miser_params", "eqs_report"):
        c_body = re.search(r"typedef struct \{([^{}]*)\} %s;" % name, hdr).group(1)
        c_body = re.sub(r"/\*.*?\*/", "", c_body, flags=re.S)
        c_fields = []
        for decl in c_body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            names = decl.split(None, 1)[1] if not decl.startswith("const") else decl.split(None, 2)[2]
            c_fields += [f.strip().lstrip("*") for f in names.split(",")]
        r_body = re.search(r"pub struct %s \{(.*?)\n\}" % name, src, re.S).group(1)
        r_fields = re.findall(r"pub ([a-z0-9_]+):", r_body)
        assert r_fields == c_fields, (name, r_fields, c_fields)
