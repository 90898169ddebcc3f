"""Host-side mirror of `eqsolver::global_optimisers` (reference src/solvers/global_optimisers/mod.rs:1-5).

Same builder surface as the Rust reference -- `ParticleSwarm.new(f, lower, upper).with_particle_count(..)
.with_iter_max(..).solve(x0)` (particle_swarm.rs:119-153, 179-333, 356) and `CrossEntropy.new(f).with_std_dev(..)
.solve(x0)` (cross_entropy.rs:75-85, 108-223, 245) -- with the closure `f` replaced by an `Objective` ID because
a host closure cannot run on the device.  Every call goes through the C ABI (include/eqsolver_b200.h) into the
sm_100a kernels; there is no CPU fallback.  The Rust twin of this file is rust/eqsolver-b200/src/.
"""
from __future__ import annotations

import ctypes as C
import enum
from typing import Optional, Sequence

import numpy as np

from . import _ffi
from ._ffi import CeParams, PsoParams, Report, dp, i64p

# reference defaults: lib.rs:13,16; particle_swarm.rs:8-12; cross_entropy.rs:8-9
DEFAULT_TOL = 1e-6
DEFAULT_ITERMAX = 50
DEFAULT_INERTIA_WEIGHT = 0.5
DEFAULT_COGNITIVE_COEFFICIENT = 1.0
DEFAULT_SOCIAL_COEFFICIENT = 1.0
DEFAULT_PARTICLE_COUNT = 100
DEFAULT_STALL_ITERATIONS = 50
DEFAULT_SAMPLE_SIZE = 100
DEFAULT_IMPORTANCE_SELECTION_SIZE = 10


class Objective(enum.IntEnum):
    """Ahead-of-time compiled device functors standing in for the closure `F: Fn(Vector) -> T`."""
    SPHERE = 0
    ROSENBROCK = 1
    RASTRIGIN = 2
    ACKLEY = 3
    THREE_CIRCLES = 4
    HEAVY = 5
    USER = 6
    GAUSSIAN = 7       # exp(-sum x^2), tests/integrators.rs:11-13
    SINCOS = 8         # sin(cos(x1) + x0), tests/integrators.rs:15-17
    SERIES = 9         # sum_{i<1000} (-1)^i x^i, benchmarks/integrators.rs:69-75


class PsoMode(enum.IntEnum):
    SEQUENTIAL_EXACT = 0  # the reference's semantics: gbest moves inside the particle loop
    SYNCHRONOUS = 1       # one gbest update per iteration (throughput mode)


class MeanDivisor(enum.IntEnum):
    REFERENCE_TEN = 0  # the literal 10 of cross_entropy.rs:287
    ELITE_COUNT = 1


class Exchange(enum.IntEnum):
    AUTO = 0
    MANUAL = 1


class SolverError(Exception):
    """`SolverError` of the reference, src/lib.rs:19-44."""
    status = -1

    def __init__(self, message: str = ""):
        super().__init__(message)
        self.message = message


class MaxIterReached(SolverError):
    status = _ffi.ERR_MAX_ITER


class NotANumber(SolverError):
    status = _ffi.ERR_NAN


class IncorrectInput(SolverError):
    status = _ffi.ERR_INCORRECT_INPUT


class BadJacobian(SolverError):
    status = _ffi.ERR_BAD_JACOBIAN


class TypeConversionError(SolverError):
    status = _ffi.ERR_TYPE_CONVERSION


class ExternalError(SolverError):
    status = _ffi.ERR_EXTERNAL


_ERRORS = {c.status: c for c in (MaxIterReached, NotANumber, IncorrectInput, BadJacobian, TypeConversionError,
                                 ExternalError)}


def _arr(a, n: Optional[int] = None) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.float64).reshape(-1)
    if n is not None and a.size != n:
        raise IncorrectInput(f"expected a vector of {n} elements, got {a.size}")
    return a


def _p(a: Optional[np.ndarray]):
    return a.ctypes.data_as(dp) if a is not None else None


class Context:
    """One GPU (one process per GPU).  Owns the stream and, for world > 1, the NCCL communicator."""

    def __init__(self, device: int = 0):
        self._h = C.c_void_p()
        st = _ffi.lib().eqs_ctx_create(int(device), C.byref(self._h))
        if st != _ffi.OK:
            msg = _ffi.lib().eqs_last_error(None)
            raise _ERRORS.get(st, SolverError)(msg.decode() if msg else "eqs_ctx_create failed")
        self.device = int(device)

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            _ffi.lib().eqs_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self, st: int):
        if st != _ffi.OK:
            msg = _ffi.lib().eqs_last_error(self._h)
            raise _ERRORS.get(st, SolverError)(msg.decode() if msg else f"status {st}")

    @property
    def sm_count(self) -> int:
        v = C.c_int(0)
        self.check(_ffi.lib().eqs_ctx_sm_count(self._h, C.byref(v)))
        return v.value

    @property
    def rank_world(self):
        r, w = C.c_int(0), C.c_int(1)
        self.check(_ffi.lib().eqs_ctx_rank(self._h, C.byref(r), C.byref(w)))
        return r.value, w.value

    def comm_init(self, rank: int, world: int, unique_id: bytes):
        buf = C.create_string_buffer(unique_id, 128)
        self.check(_ffi.lib().eqs_ctx_comm_init(self._h, rank, world, buf))

    # diagnostics
    def objective_eval(self, objective, x) -> np.ndarray:
        x = np.ascontiguousarray(x, dtype=np.float64)
        if x.ndim == 1:
            x = x[None, :]
        out = np.empty(x.shape[0])
        self.check(_ffi.lib().eqs_objective_eval(self._h, int(objective), _p(x), x.shape[1], x.shape[0], _p(out)))
        return out

    def philox_blocks(self, ctr, key) -> np.ndarray:
        ctr = np.ascontiguousarray(ctr, dtype=np.uint32).reshape(-1, 4)
        key = np.ascontiguousarray(key, dtype=np.uint32).reshape(2)
        out = np.empty_like(ctr)
        u32p = C.POINTER(C.c_uint32)
        self.check(_ffi.lib().eqs_philox_blocks(self._h, ctr.ctypes.data_as(u32p), key.ctypes.data_as(u32p),
                                                ctr.shape[0], out.ctypes.data_as(u32p)))
        return out

    def draws(self, kind: int, seed: int, t: int, first: int, count: int, n: int):
        a = np.empty((count, n)); b = np.empty((count, n))
        self.check(_ffi.lib().eqs_draws(self._h, kind, C.c_uint64(seed), t, first, count, n, _p(a), _p(b)))
        return (a,) if kind == 2 else (a, b)

    def device_cos(self, x) -> np.ndarray:
        x = np.ascontiguousarray(x, dtype=np.float64).reshape(-1)
        out = np.empty_like(x)
        self.check(_ffi.lib().eqs_device_cos(self._h, _p(x), x.size, _p(out)))
        return out

    def measure_fp64(self) -> float:
        v = C.c_double(0)
        self.check(_ffi.lib().eqs_measure_fp64(self._h, C.byref(v)))
        return v.value

    def measure_copy(self, nbytes: int = 1 << 30) -> float:
        v = C.c_double(0)
        self.check(_ffi.lib().eqs_measure_copy(self._h, nbytes, C.byref(v)))
        return v.value

    def measure_pattern(self, n: int = 32, particles: int = 1 << 20) -> float:
        """GB/s of the ParticleSwarm step's access pattern with the arithmetic removed (its own ceiling)."""
        v = C.c_double(0)
        self.check(_ffi.lib().eqs_measure_pattern(self._h, n, particles, C.byref(v)))
        return v.value


def unique_id() -> bytes:
    buf = C.create_string_buffer(128)
    st = _ffi.lib().eqs_comm_unique_id(buf)
    if st != _ffi.OK:
        msg = _ffi.lib().eqs_last_error(None)
        raise ExternalError(msg.decode() if msg else "eqs_comm_unique_id failed")
    return buf.raw


def shard_range(count: int, world: int, rank: int):
    f, l = C.c_int64(0), C.c_int64(0)
    st = _ffi.lib().eqs_shard_range(count, world, rank, C.byref(f), C.byref(l))
    if st != _ffi.OK:
        raise IncorrectInput("bad shard arguments")
    return f.value, l.value


_default_ctx: Optional[Context] = None


def default_context() -> Context:
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context(0)
    return _default_ctx


class ParticleSwarm:
    """`ParticleSwarm` of the reference (particle_swarm.rs:90-105), objective by ID."""

    def __init__(self, objective, lower_bounds: Sequence[float], upper_bounds: Sequence[float]):
        self.objective = Objective(int(objective))
        self.lower = _arr(lower_bounds)
        self.upper = _arr(upper_bounds, self.lower.size)
        self.inertia_weight = DEFAULT_INERTIA_WEIGHT
        self.cognitive_coefficient = DEFAULT_COGNITIVE_COEFFICIENT
        self.social_coefficient = DEFAULT_SOCIAL_COEFFICIENT
        self.tolerance = DEFAULT_TOL
        self.iter_max = DEFAULT_ITERMAX
        self.particle_count = DEFAULT_PARTICLE_COUNT
        # what the closure-free, seedable, sharded build adds
        self.seed = 0
        self.mode = PsoMode.SEQUENTIAL_EXACT
        self.exchange = Exchange.AUTO
        self.virtual_rank, self.virtual_world = 0, 1
        self.replay_init: Optional[np.ndarray] = None
        self.context: Optional[Context] = None
        self.last_report: Optional[Report] = None
        # ParticleSwarm::new fails on bad bounds (particle_swarm.rs:124-139) -> SolverError::ExternalError
        st = _ffi.lib().eqs_pso_validate(C.byref(self._params()))
        if st != _ffi.OK:
            raise _ERRORS.get(st, SolverError)("low > high (or non-finite) in uniform distribution"
                                               if st == _ffi.ERR_EXTERNAL else "incorrect input")

    @classmethod
    def new(cls, objective, lower_bounds, upper_bounds) -> "ParticleSwarm":
        return cls(objective, lower_bounds, upper_bounds)

    # --- the reference's setters, particle_swarm.rs:179-333 ---
    def with_inertia_weight(self, v: float): self.inertia_weight = float(v); return self
    def with_cognitive_coefficient(self, v: float): self.cognitive_coefficient = float(v); return self
    def with_social_coefficient(self, v: float): self.social_coefficient = float(v); return self
    def with_tol(self, v: float): self.tolerance = float(v); return self
    def with_particle_count(self, v: int): self.particle_count = int(v); return self
    def with_iter_max(self, v: int): self.iter_max = int(v); return self
    # --- additions ---
    def with_seed(self, v: int): self.seed = int(v); return self
    def with_mode(self, v): self.mode = PsoMode(int(v)); return self
    def with_context(self, ctx: Context): self.context = ctx; return self
    def with_replay_init(self, draws): self.replay_init = None if draws is None else _arr(draws); return self

    def with_virtual_rank(self, rank: int, world: int):
        self.exchange = Exchange.MANUAL; self.virtual_rank, self.virtual_world = int(rank), int(world); return self

    def _params(self) -> PsoParams:
        p = PsoParams()
        p.n = self.lower.size
        p.objective_id = int(self.objective)
        p.mode = int(self.mode)
        p.exchange = int(self.exchange)
        p.particle_count = self.particle_count
        p.iter_max = self.iter_max
        p.tol = self.tolerance
        p.w, p.c1, p.c2 = self.inertia_weight, self.cognitive_coefficient, self.social_coefficient
        p.lower = _p(self.lower)
        p.upper = _p(self.upper)
        p.seed = self.seed
        if self.replay_init is not None:
            if self.replay_init.size != self.particle_count * 2 * self.lower.size:
                raise IncorrectInput("replay_init must hold particle_count * 2n draws")
            p.replay_init = _p(self.replay_init)
        p.virtual_rank, p.virtual_world = self.virtual_rank, self.virtual_world
        return p

    def solve(self, x0) -> np.ndarray:
        """`solve(x0)`, particle_swarm.rs:356-431: returns the global best position."""
        ctx = self.context or default_context()
        x0 = _arr(x0, self.lower.size)
        out = np.empty_like(x0)
        rep = Report()
        ctx.check(_ffi.lib().eqs_pso_solve(ctx._h, C.byref(self._params()), _p(x0), _p(out), C.byref(rep)))
        self.last_report = rep
        return out

    def solve_with_seed(self, x0, seed: int) -> np.ndarray:
        """`solve` with an explicit generator state, the optimisers' twin of
        `OrthotopeRandomIntegrator::integrate_with_rng` (integrators/mod.rs:56-61): same seed, same result."""
        self.seed = int(seed)
        return self.solve(x0)

    def solve_with_draws(self, x0, init_draws, step_draws) -> np.ndarray:
        """`solve` on host-supplied draws instead of a generator (the replay mode of the parity tests as a feature).
        `init_draws` [N][2][n]: u in [0, 1) for the start positions, then for the velocities, which the samplers of
        particle_swarm.rs:367,373 scale into their ranges; `step_draws` [passes][N][n][2]: (r_p, r_g) (:407-408).  Runs
        min(passes, iter_max) passes, or fewer when the stall test (:433-441) fires."""
        init_draws = _arr(init_draws, self.particle_count * 2 * self.lower.size)
        per_pass = self.particle_count * self.lower.size * 2
        step_draws = _arr(step_draws)
        if step_draws.size % per_pass:
            raise IncorrectInput("step_draws must hold whole passes of particle_count * n * 2 draws")
        saved = self.replay_init
        self.replay_init = init_draws
        try:
            run = self.start(x0)
        finally:
            self.replay_init = saved
        try:
            for t in range(min(step_draws.size // per_pass, self.iter_max)):
                run.step(1, step_draws[t * per_pass:(t + 1) * per_pass])
            return run.gbest()[0]
        finally:
            run.close()

    def start(self, x0=None) -> "PsoRun":
        """Step-granular twin of solve() with the population resident in HBM (parity, tracing, benchmarks)."""
        run = PsoRun(self.context or default_context(), self)
        if x0 is not None:
            run.init(x0)
        return run


class PsoRun:
    def __init__(self, ctx: Context, cfg: ParticleSwarm):
        self.ctx, self.cfg = ctx, cfg
        self.n = cfg.lower.size
        self._params = cfg._params()  # keeps the bound arrays alive through cfg
        self._h = C.c_void_p()
        ctx.check(_ffi.lib().eqs_pso_create(ctx._h, C.byref(self._params), C.byref(self._h)))
        f, l = C.c_int64(0), C.c_int64(0)
        ctx.check(_ffi.lib().eqs_pso_local_range(self._h, C.byref(f), C.byref(l)))
        self.first, self.local = f.value, l.value

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            _ffi.lib().eqs_pso_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def init(self, x0):
        self.ctx.check(_ffi.lib().eqs_pso_init(self._h, _p(_arr(x0, self.n))))

    def step(self, iters: int = 1, replay_step=None):
        r = None if replay_step is None else _arr(replay_step, iters * self.cfg.particle_count * self.n * 2)
        self.ctx.check(_ffi.lib().eqs_pso_step(self._h, iters, _p(r)))

    def sync(self):
        self.ctx.check(_ffi.lib().eqs_pso_sync(self._h))

    def last_step_ms(self):
        ms, k = C.c_double(0), C.c_int64(0)
        self.ctx.check(_ffi.lib().eqs_pso_last_step_ms(self._h, C.byref(ms), C.byref(k)))
        return ms.value, k.value

    def gbest(self):
        g = np.empty(self.n)
        cost, iters, stalled, imp = C.c_double(0), C.c_int64(0), C.c_int32(0), C.c_int64(0)
        self.ctx.check(_ffi.lib().eqs_pso_read_gbest(self._h, _p(g), C.byref(cost), C.byref(iters), C.byref(stalled),
                                                     C.byref(imp)))
        return g, cost.value, iters.value, bool(stalled.value), imp.value

    def state(self):
        x = np.empty((self.local, self.n)); v = np.empty_like(x); pb = np.empty_like(x); pbc = np.empty(self.local)
        self.ctx.check(_ffi.lib().eqs_pso_read_state(self._h, _p(x), _p(v), _p(pb), _p(pbc)))
        return x, v, pb, pbc

    def local_indices(self) -> np.ndarray:
        """Global index of every local particle, in the row order of state(): first .. first + local - 1, except in
        sequential-exact mode over several ranks (tiles of 256 particles dealt round-robin over the ranks)."""
        idx = np.empty(self.local, dtype=np.int64)
        self.ctx.check(_ffi.lib().eqs_pso_local_indices(self._h, idx.ctypes.data_as(C.POINTER(C.c_int64))))
        return idx

    def pbest_costs(self) -> np.ndarray:
        """Only the pbest costs of this rank's particles (state() also moves 3 n-vectors per particle)."""
        pbc = np.empty(self.local)
        self.ctx.check(_ffi.lib().eqs_pso_read_state(self._h, None, None, None, _p(pbc)))
        return pbc

    def ring(self):
        r = np.empty(_ffi.STALL_WINDOW); i = C.c_int64(0)
        self.ctx.check(_ffi.lib().eqs_pso_read_ring(self._h, _p(r), C.byref(i)))
        return r, i.value

    # split-phase protocol (virtual ranks)
    def init_local(self, x0, phase: int):
        self.ctx.check(_ffi.lib().eqs_pso_init_local(self._h, _p(_arr(x0, self.n)), phase))

    def init_apply(self):
        self.ctx.check(_ffi.lib().eqs_pso_init_apply(self._h))

    def step_local(self, replay_step=None):
        r = None if replay_step is None else _arr(replay_step, self.cfg.particle_count * self.n * 2)
        self.ctx.check(_ffi.lib().eqs_pso_step_local(self._h, _p(r)))

    def step_apply(self):
        self.ctx.check(_ffi.lib().eqs_pso_step_apply(self._h))

    def get_record(self) -> np.ndarray:
        rec = np.empty(_ffi.lib().eqs_record_doubles(self.n))
        self.ctx.check(_ffi.lib().eqs_pso_get_record(self._h, _p(rec)))
        return rec

    def set_records(self, recs):
        recs = np.ascontiguousarray(recs, dtype=np.float64)
        self.ctx.check(_ffi.lib().eqs_pso_set_records(self._h, _p(recs)))


class CrossEntropy:
    """`CrossEntropy` of the reference (cross_entropy.rs:51-65), objective by ID."""

    def __init__(self, objective):
        self.objective = Objective(int(objective))
        self.std_dev: Optional[np.ndarray] = None
        self.tolerance = DEFAULT_TOL
        self.iter_max = DEFAULT_ITERMAX
        self.sample_size = DEFAULT_SAMPLE_SIZE
        self.importance_selection_size = DEFAULT_IMPORTANCE_SELECTION_SIZE
        self.seed = 0
        self.mean_divisor = MeanDivisor.REFERENCE_TEN
        self.exchange = Exchange.AUTO
        self.virtual_rank, self.virtual_world = 0, 1
        self.context: Optional[Context] = None
        self.last_report: Optional[Report] = None

    @classmethod
    def new(cls, objective) -> "CrossEntropy":
        return cls(objective)

    # --- the reference's setters, cross_entropy.rs:108-223 ---
    def with_tol(self, v: float): self.tolerance = float(v); return self
    def with_iter_max(self, v: int): self.iter_max = int(v); return self
    def with_sample_size(self, v: int): self.sample_size = int(v); return self
    def with_importance_selection_size(self, v: int): self.importance_selection_size = int(v); return self
    def with_std_dev(self, v): self.std_dev = _arr(v); return self
    # --- additions ---
    def with_seed(self, v: int): self.seed = int(v); return self
    def with_mean_divisor(self, v): self.mean_divisor = MeanDivisor(int(v)); return self
    def with_context(self, ctx: Context): self.context = ctx; return self

    def with_virtual_rank(self, rank: int, world: int):
        self.exchange = Exchange.MANUAL; self.virtual_rank, self.virtual_world = int(rank), int(world); return self

    def _params(self, n: int) -> CeParams:
        p = CeParams()
        p.n = n
        p.objective_id = int(self.objective)
        p.mean_divisor = int(self.mean_divisor)
        p.exchange = int(self.exchange)
        p.sample_size = self.sample_size
        p.elite_count = self.importance_selection_size
        p.iter_max = self.iter_max
        p.tol = self.tolerance
        if self.std_dev is not None:
            if self.std_dev.size != n:
                raise IncorrectInput("std_dev and x0 differ in length")
            p.std_dev = _p(self.std_dev)
        p.seed = self.seed
        p.virtual_rank, p.virtual_world = self.virtual_rank, self.virtual_world
        return p

    def solve(self, x0) -> np.ndarray:
        """`solve(x0)`, cross_entropy.rs:245-306: returns the final mean."""
        ctx = self.context or default_context()
        x0 = _arr(x0)
        out = np.empty_like(x0)
        rep = Report()
        ctx.check(_ffi.lib().eqs_ce_solve(ctx._h, C.byref(self._params(x0.size)), _p(x0), _p(out), C.byref(rep)))
        self.last_report = rep
        return out

    def solve_with_seed(self, x0, seed: int) -> np.ndarray:
        """`solve` with an explicit generator state (see ParticleSwarm.solve_with_seed)."""
        self.seed = int(seed)
        return self.solve(x0)

    def solve_with_draws(self, x0, normals) -> np.ndarray:
        """`solve` on host-supplied standard normals `normals` [passes][N][n] (the z of `Normal::sample`,
        cross_entropy.rs:270); runs min(passes, iter_max - 1) passes (`iter` starts at 1, :253) or stops on tol."""
        x0 = _arr(x0)
        per_pass = self.sample_size * x0.size
        normals = _arr(normals)
        if normals.size % per_pass:
            raise IncorrectInput("normals must hold whole passes of sample_size * n draws")
        run = self.start(x0)
        try:
            for t in range(min(normals.size // per_pass, max(0, self.iter_max - 1))):
                run.step(1, normals[t * per_pass:(t + 1) * per_pass])
            return run.state()[0]
        finally:
            run.close()

    def start(self, x0) -> "CeRun":
        return CeRun(self.context or default_context(), self, _arr(x0))


class CeRun:
    def __init__(self, ctx: Context, cfg: CrossEntropy, x0: np.ndarray):
        self.ctx, self.cfg = ctx, cfg
        self.n = x0.size
        self._params = cfg._params(self.n)
        self._h = C.c_void_p()
        ctx.check(_ffi.lib().eqs_ce_create(ctx._h, C.byref(self._params), _p(x0), C.byref(self._h)))
        f, l = C.c_int64(0), C.c_int64(0)
        ctx.check(_ffi.lib().eqs_ce_local_range(self._h, C.byref(f), C.byref(l)))
        self.first, self.local = f.value, l.value

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            _ffi.lib().eqs_ce_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def step(self, iters: int = 1, replay_normals=None):
        r = None if replay_normals is None else _arr(replay_normals, iters * self.cfg.sample_size * self.n)
        self.ctx.check(_ffi.lib().eqs_ce_step(self._h, iters, _p(r)))

    def sync(self):
        self.ctx.check(_ffi.lib().eqs_ce_sync(self._h))

    def last_step_ms(self):
        ms, k = C.c_double(0), C.c_int64(0)
        self.ctx.check(_ffi.lib().eqs_ce_last_step_ms(self._h, C.byref(ms), C.byref(k)))
        return ms.value, k.value

    def state(self):
        mu = np.empty(self.n); sg = np.empty(self.n)
        iters, active = C.c_int64(0), C.c_int32(0)
        self.ctx.check(_ffi.lib().eqs_ce_read_state(self._h, _p(mu), _p(sg), C.byref(iters), C.byref(active)))
        return mu, sg, iters.value, bool(active.value)

    def elites(self) -> np.ndarray:
        idx = np.empty(self.cfg.importance_selection_size, dtype=np.int64)
        cnt = C.c_int64(0)
        self.ctx.check(_ffi.lib().eqs_ce_read_elites(self._h, idx.ctypes.data_as(i64p), C.byref(cnt)))
        return idx[:cnt.value]

    def costs(self) -> np.ndarray:
        c = np.empty(self.local)
        self.ctx.check(_ffi.lib().eqs_ce_read_costs(self._h, _p(c)))
        return c

    # split-phase protocol (virtual ranks): phases 0..9 of one pass
    N_PHASES = 10

    def phase(self, phase: int, replay_normals=None):
        r = None if replay_normals is None else _arr(replay_normals, self.cfg.sample_size * self.n)
        self.ctx.check(_ffi.lib().eqs_ce_phase(self._h, phase, _p(r)))

    def exchange_words(self, phase: int) -> int:
        return int(_ffi.lib().eqs_ce_exchange_words(self._h, phase))

    def get_exchange(self, phase: int) -> np.ndarray:
        w = np.empty(self.exchange_words(phase), dtype=np.uint64)
        self.ctx.check(_ffi.lib().eqs_ce_get_exchange(self._h, phase, w.ctypes.data_as(C.c_void_p)))
        return w

    def set_exchange(self, phase: int, words):
        w = np.ascontiguousarray(words, dtype=np.uint64)
        self.ctx.check(_ffi.lib().eqs_ce_set_exchange(self._h, phase, w.ctypes.data_as(C.c_void_p)))
