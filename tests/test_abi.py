"""CPU: the C-ABI library loads and exports every symbol include/eqsolver_b200.h declares; host-only entry points
behave; compute entry points fail loudly without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT


def header_symbols():
    text = open(os.path.join(ROOT, "include", "eqsolver_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(eqs_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from eqsolver_b200 import _ffi
    lib = C.CDLL(_ffi.LIB_PATH)
    names = header_symbols()
    assert len(names) >= 40
    for name in names:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert set(names) == set(_ffi.SYMBOLS), "ctypes table and header disagree"
    assert _ffi.lib().eqs_abi_version() == 1


def test_host_only_entry_points():
    import eqsolver_b200 as E
    assert E.shard_range(10, 3, 0) == (0, 4) and E.shard_range(10, 3, 1) == (4, 3) and E.shard_range(10, 3, 2) == (7, 3)
    assert E.shard_range(2, 4, 3) == (2, 0)
    covered = 0
    for r in range(8):
        f, l = E.shard_range(1000003, 8, r)
        assert f == covered
        covered += l
    assert covered == 1000003
    from eqsolver_b200 import _ffi
    assert _ffi.lib().eqs_record_doubles(32) == 4 + 50 + 32


def test_new_rejects_bad_bounds_like_the_reference():
    import eqsolver_b200 as E
    # ParticleSwarm::new -> Err(SolverError::ExternalError(..)) when lower > upper (particle_swarm.rs:124-139)
    with pytest.raises(E.ExternalError):
        E.ParticleSwarm.new(E.Objective.RASTRIGIN, [1.0, 0.0], [0.0, 1.0])
    with pytest.raises(E.ExternalError):
        E.ParticleSwarm.new(E.Objective.RASTRIGIN, [0.0], [float("nan")])
    with pytest.raises(E.IncorrectInput):
        E.ParticleSwarm.new(E.Objective.THREE_CIRCLES, [0.0] * 3, [1.0] * 3)
    ps = E.ParticleSwarm.new(E.Objective.RASTRIGIN, [-1.0] * 4, [1.0] * 4)
    # reference defaults: particle_swarm.rs:8-12, lib.rs:13,16
    assert (ps.inertia_weight, ps.cognitive_coefficient, ps.social_coefficient) == (0.5, 1.0, 1.0)
    assert (ps.particle_count, ps.iter_max, ps.tolerance) == (100, 50, 1e-6)
    assert ps.with_particle_count(7).with_iter_max(3).with_tol(0.5).with_inertia_weight(0.1) is ps
    ce = E.CrossEntropy.new(E.Objective.RASTRIGIN)
    assert (ce.sample_size, ce.importance_selection_size, ce.iter_max, ce.tolerance) == (100, 10, 50, 1e-6)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: the loud-failure path is for GPU-less hosts")
    import eqsolver_b200 as E
    with pytest.raises(E.ExternalError, match="no CUDA device"):
        E.ParticleSwarm.new(E.Objective.SPHERE, [-1.0], [1.0]).solve([0.0])
