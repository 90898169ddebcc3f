"""CPU: the arithmetic of csrc/devmath.cuh / philox.cuh::normal_pair, restated in C (tests/devmath_model.c), against glibc.

The model is test infrastructure: it cannot see the device code, it restates it (exact fma, a 20-bit stand-in for the MUFU
seeds), so what this pins is the FORMULAS' accuracy -- the figures DESIGN.md section 4 quotes -- while the device code is
pinned on the GPU (test_pso_gpu.py::test_device_cos_accuracy, ::test_box_muller_normals_stay_within_ulps_of_the_oracle)."""
import json
import os
import shutil
import subprocess

import pytest

from conftest import ROOT


def test_formulas_against_glibc(tmp_path):
    gcc = shutil.which("gcc")
    if not gcc:
        pytest.skip("no gcc")
    exe = tmp_path / "devmath_model"
    subprocess.run([gcc, "-O2", "-ffp-contract=off", "-o", str(exe), os.path.join(ROOT, "tests", "devmath_model.c"), "-lm"],
                   check=True, capture_output=True, timeout=120)
    r = subprocess.run([str(exe), "1500000"], check=True, capture_output=True, text=True, timeout=120)
    d = json.loads(r.stdout)
    assert d["log_max_ulp"] <= 1.0 + 1e-9                 # fdlibm's log with fused tail (log_unit): within 1 ulp of glibc
    assert d["log_tab_max_ulp"] <= 2.0 + 1e-9             # the table-driven log of Box-Muller (log_unit_tab): within 2 ulp
    # one even polynomial on [-pi/2, pi/2]: absolute bound 1.25 ulp of 1.0 (round 1's two-kernel cosine: 0.5 ulp of 1.0)
    assert d["cos_max_abs"] <= 2.8e-16 and d["cos_max_ulp_beyond_2.8e-16"] == 0.0   # the bound of the GPU test
    assert d["cos_round1_max_abs"] <= 1.2e-16
    assert d["sqrt_mismatches"] == 0                      # the branch-free root is the correctly rounded one
    assert d["sqrt_zero"] == 0 and d["sqrt_minus_zero_is_negative"] == 1 and d["log_one"] == 0
    assert d["normals_equal_share"] > 0.88 and d["normals_max_units"] <= 4.0   # DESIGN: 91 % identical, within 3
