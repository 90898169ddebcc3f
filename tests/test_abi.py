"""CPU: the C-ABI library loads and exports every symbol include/eqsolver_b200.h declares; host-only entry points
behave; compute entry points fail loudly without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
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
    assert _ffi.lib().eqs_record_doubles(32) == 4 + 50 + 50 + 32


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


def test_validate_entry_points_survive_garbage():
    """The host-only *_validate entry points (what `new` / the builders call before any GPU work) on a seeded stream of
    hostile parameter blocks: null pointers, NaN / infinite / inverted bounds, zero and negative sizes, huge counts.
    They must answer with a status (never crash), and agree with the reference's rules where those are decidable."""
    import ctypes as C
    from eqsolver_b200 import _ffi
    L = _ffi.lib()
    rng = np.random.default_rng(99)
    specials = [0.0, -0.0, 1.0, -1.0, np.inf, -np.inf, np.nan, 1e308, -1e308, 5e-324]
    ints = [0, 1, 2, -1, 3, 7, 100, 12800, 12801, 2 ** 31 - 1, -2 ** 31]
    bigs = [0, 1, 2, -1, 10, 2 ** 40, 2 ** 40 + 1, 2 ** 62, -2 ** 62]
    statuses = set(range(7))
    for case in range(600):
        n = int(rng.choice(ints))
        m = max(1, min(abs(n), 64))
        lo = np.array([rng.choice(specials) if rng.random() < 0.3 else rng.uniform(-5, 5) for _ in range(m)])
        hi = np.array([rng.choice(specials) if rng.random() < 0.3 else rng.uniform(-5, 5) for _ in range(m)])
        plo = None if rng.random() < 0.1 else lo.ctypes.data_as(_ffi.dp)
        phi = None if rng.random() < 0.1 else hi.ctypes.data_as(_ffi.dp)
        obj = int(rng.integers(-2, 13))
        p = _ffi.PsoParams()
        p.n, p.objective_id, p.mode, p.exchange = n, obj, int(rng.integers(-1, 4)), int(rng.integers(-1, 4))
        p.particle_count, p.iter_max = int(rng.choice(bigs)), int(rng.choice(bigs))
        p.tol, p.w, p.c1, p.c2 = (float(rng.choice(specials)) for _ in range(4))
        p.lower, p.upper = plo, phi
        p.virtual_rank, p.virtual_world = int(rng.integers(-2, 5)), int(rng.integers(-2, 5))
        st = L.eqs_pso_validate(C.byref(p)) if 0 < n <= 64 or plo is None or phi is None or n <= 0 or n > 12800 else None
        if st is not None:
            assert st in statuses
            if st == _ffi.OK:   # what the reference's `new` requires (particle_swarm.rs:124-139)
                assert np.all(lo[:n] <= hi[:n]) and np.all(np.isfinite(lo[:n])) and np.all(np.isfinite(hi[:n]))
        c = _ffi.CeParams()
        sd = np.array([rng.choice(specials) if rng.random() < 0.3 else rng.uniform(0.1, 3) for _ in range(m)])
        c.n, c.objective_id, c.mean_divisor, c.exchange = n, obj, int(rng.integers(-1, 3)), int(rng.integers(-1, 4))
        c.sample_size, c.elite_count, c.iter_max = int(rng.choice(bigs)), int(rng.choice(bigs)), int(rng.choice(bigs))
        c.tol = float(rng.choice(specials))
        c.std_dev = None if rng.random() < 0.3 or not (0 < n <= 64) else sd.ctypes.data_as(_ffi.dp)
        c.virtual_rank, c.virtual_world = int(rng.integers(-2, 5)), int(rng.integers(-2, 5))
        st = L.eqs_ce_validate(C.byref(c))
        assert st in statuses
        if st == _ffi.OK:
            assert 2 <= c.elite_count <= c.sample_size and 0 < n <= 12800
        if 0 < n <= 64 or plo is None or phi is None or n <= 0 or n > 12800:
            q = _ffi.McParams()
            q.n, q.integrand_id, q.sample_count = n, obj, int(rng.choice(bigs))
            q.from_, q.to = plo, phi
            st = L.eqs_mc_validate(C.byref(q))
            assert st in statuses
            if st == _ffi.OK:
                assert q.sample_count >= 2 and np.all(lo[:n] <= hi[:n])
            r = _ffi.MiserParams()
            r.n, r.integrand_id, r.sample_count = n, obj, int(rng.choice(bigs))
            r.minimum_dimensional_sample_count, r.maximum_depth = int(rng.choice(bigs)), int(rng.choice(ints))
            r.alpha, r.dither = float(rng.choice(specials)), float(rng.choice(specials))
            r.from_, r.to = plo, phi
            st = L.eqs_miser_validate(C.byref(r))
            assert st in statuses
            if st == _ffi.OK:
                assert 0 <= r.maximum_depth <= 30 and 0 <= r.sample_count <= 2 ** 40
    assert L.eqs_pso_validate(None) != _ffi.OK and L.eqs_ce_validate(None) != _ffi.OK
    assert L.eqs_mc_validate(None) != _ffi.OK and L.eqs_miser_validate(None) != _ffi.OK


def test_null_handles_are_refused_not_dereferenced():
    """Every entry point taking a handle answers EQS_ERR_INCORRECT_INPUT for NULL (destroy(NULL) is a no-op)."""
    from eqsolver_b200 import _ffi
    L = _ffi.lib()
    null = C.c_void_p()
    calls = {
        'eqs_pso_init': (null, None), 'eqs_pso_step': (null, 1, None), 'eqs_pso_sync': (null,),
        'eqs_pso_read_gbest': (null, None, None, None, None, None), 'eqs_pso_read_state': (null, None, None, None, None),
        'eqs_ce_step': (null, 1, None), 'eqs_ce_sync': (null,), 'eqs_ce_read_state': (null, None, None, None, None),
        'eqs_ce_phase': (null, 0, None), 'eqs_ctx_sm_count': (null, None), 'eqs_ctx_rank': (null, None, None),
        'eqs_objective_eval': (null, 0, None, 1, 1, None), 'eqs_mc_integrate': (null, None, None, None, None),
        'eqs_miser_integrate': (null, None, None, None, None), 'eqs_pso_solve': (null, None, None, None, None),
        'eqs_ce_solve': (null, None, None, None, None), 'eqs_pso_create': (null, None, None),
        'eqs_ce_create': (null, None, None, None), 'eqs_measure_copy': (null, 100, None),
        'eqs_measure_pattern': (null, 32, 100, None), 'eqs_measure_fp64': (null, None),
        'eqs_device_cos': (null, None, 1, None), 'eqs_pso_last_step_ms': (null, None, None),
        'eqs_pso_read_ring': (null, None, None), 'eqs_pso_local_range': (null, None, None),
        'eqs_pso_local_indices': (null, None),
        'eqs_pso_init_local': (null, None, 0), 'eqs_pso_init_apply': (null,), 'eqs_pso_step_local': (null, None),
        'eqs_pso_step_apply': (null,), 'eqs_pso_get_record': (null, None), 'eqs_pso_set_records': (null, None),
        'eqs_ce_last_step_ms': (null, None, None), 'eqs_ce_read_elites': (null, None, None),
        'eqs_ce_read_costs': (null, None), 'eqs_ce_local_range': (null, None, None),
        'eqs_ce_get_exchange': (null, 0, None), 'eqs_ce_set_exchange': (null, 0, None),
        'eqs_philox_blocks': (null, None, None, 1, None), 'eqs_draws': (null, 0, 0, 0, 0, 1, 1, None, None),
        'eqs_ctx_comm_init': (null, 0, 1, None), 'eqs_comm_unique_id': (None,), 'eqs_shard_range': (10, 0, 0, None, None),
    }
    for name, args in calls.items():
        assert getattr(L, name)(*args) == _ffi.ERR_INCORRECT_INPUT, name
    for name in ('eqs_ctx_destroy', 'eqs_pso_destroy', 'eqs_ce_destroy'):
        assert getattr(L, name)(null) == _ffi.OK, name
    assert L.eqs_ce_exchange_words(null, 0) == 0
    # and every handle-taking symbol of the header is in the table above (or is a destroy / validate / pure function)
    pure = {'eqs_abi_version', 'eqs_last_error', 'eqs_record_doubles', 'eqs_ce_exchange_words', 'eqs_ctx_create',
            'eqs_ctx_destroy', 'eqs_pso_destroy', 'eqs_ce_destroy', 'eqs_pso_validate', 'eqs_ce_validate',
            'eqs_mc_validate', 'eqs_miser_validate'}
    assert set(_ffi.SYMBOLS) - pure == set(calls), sorted(set(_ffi.SYMBOLS) - pure - set(calls))


def test_product_never_touches_the_oracle():
    """oracle/ is test infrastructure: nothing in the package (Python host layer, CUDA sources, public headers) may
    import, include, load or mention a path under it; bench.py uses it for the cpu_baseline / reference arm only."""
    import re
    from conftest import ROOT
    pat = re.compile(r"from\s+oracle|import\s+oracle|oracle[/.](oracle|eqs_oracle)|liboracle|eqs_oracle")
    offenders = []
    for top in ("eqsolver_b200", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, top)):
            for f in files:
                if not f.endswith((".py", ".cu", ".cuh", ".h", ".hpp")):
                    continue
                path = os.path.join(dirpath, f)
                with open(path, encoding="utf-8", errors="replace") as fh:
                    for i, line in enumerate(fh, 1):
                        if pat.search(line) and "oracle/eqs_oracle.c: orc_objective" not in line:
                            offenders.append(f"{os.path.relpath(path, ROOT)}:{i}: {line.strip()}")
    assert not offenders, offenders
