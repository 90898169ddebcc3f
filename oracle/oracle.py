"""ctypes front-end of the CPU oracle (oracle/eqs_oracle.c).

TEST INFRASTRUCTURE ONLY -- PARITY UNPINNED (see oracle/eqs_oracle.h).  Imported by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs; never by the product
package eqsolver_b200/.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")

SPHERE, ROSENBROCK, RASTRIGIN, ACKLEY, THREE_CIRCLES, HEAVY, USER, GAUSSIAN, SINCOS, SERIES = range(10)
SEQUENTIAL, SYNCHRONOUS = 0, 1
DIV_REFERENCE_TEN, DIV_ELITE_COUNT = 0, 1
OK, ERR_MAX_ITER, ERR_NAN, ERR_INCORRECT_INPUT, ERR_BAD_JACOBIAN, ERR_TYPE_CONVERSION, ERR_EXTERNAL = range(7)
STALL_WINDOW = 50

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int64)


def build(force: bool = False) -> str:
    """Compile the C restatement (gcc -O2 -ffp-contract=off: rustc never contracts to FMA)."""
    src = os.path.join(_HERE, "eqs_oracle.c")
    hdr = os.path.join(_HERE, "eqs_oracle.h")
    if (not force and os.path.exists(_LIB_PATH)
            and os.path.getmtime(_LIB_PATH) >= max(os.path.getmtime(src), os.path.getmtime(hdr))):
        return _LIB_PATH
    os.makedirs(os.path.dirname(_LIB_PATH), exist_ok=True)
    subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-Wall", "-o", _LIB_PATH,
                           src, "-lm", "-pthread"])
    return _LIB_PATH


class _PsoParams(C.Structure):
    _fields_ = [("n", C.c_int32), ("objective_id", C.c_int32), ("mode", C.c_int32), ("_pad", C.c_int32),
                ("particle_count", C.c_int64), ("first_index", C.c_int64), ("local_count", C.c_int64),
                ("iter_max", C.c_int64),
                ("tol", C.c_double), ("w", C.c_double), ("c1", C.c_double), ("c2", C.c_double),
                ("lower", _dp), ("upper", _dp), ("seed", C.c_uint64),
                ("replay_init", _dp), ("replay_step", _dp), ("replay_step_iters", C.c_int64)]


class _CeParams(C.Structure):
    _fields_ = [("n", C.c_int32), ("objective_id", C.c_int32), ("mean_divisor", C.c_int32), ("_pad", C.c_int32),
                ("sample_size", C.c_int64), ("first_index", C.c_int64), ("local_count", C.c_int64),
                ("elite_count", C.c_int64), ("iter_max", C.c_int64), ("tol", C.c_double),
                ("std_dev", _dp), ("seed", C.c_uint64), ("replay_normals", _dp), ("replay_iters", C.c_int64)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.orc_objective.restype = C.c_double
        L.orc_objective.argtypes = [C.c_int, _dp, C.c_int]
        L.orc_u01.restype = C.c_double
        L.orc_u01.argtypes = [C.c_uint32, C.c_uint32]
        L.orc_pso_create.restype = C.c_void_p
        L.orc_pso_create.argtypes = [C.POINTER(_PsoParams)]
        L.orc_pso_gbest_cost.restype = C.c_double
        L.orc_pso_iters.restype = C.c_int64
        L.orc_pso_last_improvements.restype = C.c_int64
        L.orc_ce_create.restype = C.c_void_p
        L.orc_ce_create.argtypes = [C.POINTER(_CeParams), _dp]
        L.orc_ce_iters.restype = C.c_int64
        L.orc_time_pso_steps.restype = C.c_double
        L.orc_time_pso_steps.argtypes = [C.c_void_p, C.c_int64]
        L.orc_time_pso_steps_mt.restype = C.c_double
        L.orc_time_pso_steps_mt.argtypes = [C.c_void_p, C.c_int64, C.c_int]
        L.orc_pso_step_mt.restype = C.c_int
        L.orc_pso_step_mt.argtypes = [C.c_void_p, C.c_int]
        L.orc_time_ce_steps.restype = C.c_double
        L.orc_time_ce_steps.argtypes = [C.c_void_p, C.c_int64]
        for name in ("orc_pso_destroy", "orc_pso_init", "orc_pso_step", "orc_pso_iters", "orc_pso_gbest_cost",
                     "orc_pso_read_gbest", "orc_pso_read_state", "orc_pso_read_ring", "orc_pso_last_improvements",
                     "orc_pso_init_local", "orc_pso_init_ring_local", "orc_pso_init_apply", "orc_pso_step_local",
                     "orc_pso_step_apply", "orc_ce_destroy", "orc_ce_step", "orc_ce_iters", "orc_ce_read_state",
                     "orc_ce_read_elites", "orc_ce_read_costs", "orc_ce_sample_local", "orc_ce_elite_sum_local",
                     "orc_ce_elite_sqsum_local", "orc_ce_apply", "orc_ce_active"):
            fn = getattr(L, name)
            if fn.argtypes is None:
                fn.argtypes = None  # variadic-style: we pass explicit ctypes objects below
        _lib = L
    return _lib


def _d(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(_dp)


def objective(obj_id: int, x) -> float:
    x, px = _d(x)
    return float(lib().orc_objective(int(obj_id), px, int(x.size)))


def philox4x32_10(ctr, key):
    c = (C.c_uint32 * 4)(*[int(v) & 0xFFFFFFFF for v in ctr])
    k = (C.c_uint32 * 2)(*[int(v) & 0xFFFFFFFF for v in key])
    out = (C.c_uint32 * 4)()
    lib().orc_philox4x32_10(c, k, out)
    return [int(v) for v in out]


def u01(w_lo: int, w_hi: int) -> float:
    return float(lib().orc_u01(w_lo, w_hi))


def pso_init_draws(seed, gidx, n):
    pos = np.empty(n); vel = np.empty(n)
    lib().orc_pso_init_draws(C.c_uint64(seed), C.c_int64(gidx), C.c_int(n), pos.ctypes.data_as(_dp),
                             vel.ctypes.data_as(_dp))
    return pos, vel


def pso_step_draws(seed, t, gidx, n):
    rp = np.empty(n); rg = np.empty(n)
    lib().orc_pso_step_draws(C.c_uint64(seed), C.c_int64(t), C.c_int64(gidx), C.c_int(n),
                             rp.ctypes.data_as(_dp), rg.ctypes.data_as(_dp))
    return rp, rg


def ce_normals(seed, t, gidx, n):
    z = np.empty(n)
    lib().orc_ce_normals(C.c_uint64(seed), C.c_int64(t), C.c_int64(gidx), C.c_int(n), z.ctypes.data_as(_dp))
    return z


def mc_draws(seed, index, n):
    u = np.empty(n)
    lib().orc_mc_draws(C.c_uint64(seed), C.c_int64(index), C.c_int(n), u.ctypes.data_as(_dp))
    return u


def mc_integrate(integrand_id, lower, upper, sample_count=1000, seed=0, replay_u=None):
    """MonteCarlo::integrate_with_rng (reference monte_carlo.rs:164-212): returns (status, mean, variance)."""
    lo, plo = _d(lower); hi, phi = _d(upper)
    n = int(lo.size)
    pr = None
    if replay_u is not None:
        r, pr = _d(replay_u); assert r.size == sample_count * n
    m, v = C.c_double(0), C.c_double(0)
    st = lib().orc_mc_integrate(C.c_int(int(integrand_id)), C.c_int(n), C.c_int64(sample_count), plo, phi,
                                C.c_uint64(seed), pr, C.byref(m), C.byref(v))
    return int(st), m.value, v.value


def mc_moments(integrand_id, lower, upper, first, count, seed=0):
    """One rank's share of a sharded MonteCarlo: (count, mean, sum of squared deviations) of points [first, first+count)."""
    lo, plo = _d(lower); hi, phi = _d(upper)
    m = np.empty(3)
    lib().orc_mc_moments(C.c_int(int(integrand_id)), C.c_int(int(lo.size)), C.c_int64(first), C.c_int64(count), plo, phi,
                         C.c_uint64(seed), m.ctypes.data_as(_dp))
    return m


def miser_integrate(integrand_id, lower, upper, sample_count=1000, maximum_depth=5, min_dim_count=32, alpha=2.0,
                    dither=0.05, seed=0):
    """MISER::integrate_with_rng (reference miser.rs:270-443): returns (status, mean, variance, evals)."""
    lo, plo = _d(lower); hi, phi = _d(upper)
    m, v, e = C.c_double(0), C.c_double(0), C.c_int64(0)
    st = lib().orc_miser_integrate(C.c_int(int(integrand_id)), C.c_int(int(lo.size)), C.c_int64(sample_count),
                                   C.c_int(maximum_depth), C.c_int64(min_dim_count), C.c_double(alpha),
                                   C.c_double(dither), plo, phi, C.c_uint64(seed), C.byref(m), C.byref(v), C.byref(e))
    return int(st), m.value, v.value, int(e.value)


class Pso:
    """Step-granular oracle ParticleSwarm (reference particle_swarm.rs:356-431)."""

    def __init__(self, objective_id, lower, upper, particle_count=100, iter_max=50, tol=1e-6, w=0.5, c1=1.0,
                 c2=1.0, mode=SEQUENTIAL, seed=0, replay_init=None, replay_step=None, first_index=0,
                 local_count=None):
        self._keep = []
        lo, plo = _d(lower); hi, phi = _d(upper)
        self._keep += [lo, hi]
        self.n = int(lo.size)
        self.N = int(particle_count)
        self.local = self.N if local_count is None else int(local_count)
        p = _PsoParams()
        p.n = self.n; p.objective_id = int(objective_id); p.mode = int(mode)
        p.particle_count = self.N; p.first_index = int(first_index); p.local_count = self.local
        p.iter_max = int(iter_max); p.tol = tol; p.w = w; p.c1 = c1; p.c2 = c2
        p.lower = plo; p.upper = phi; p.seed = int(seed)
        if replay_init is not None:
            a, pa = _d(replay_init); assert a.size == self.N * 2 * self.n
            self._keep.append(a); p.replay_init = pa
        if replay_step is not None:
            a, pa = _d(replay_step); assert a.size % (self.N * 2 * self.n) == 0
            self._keep.append(a); p.replay_step = pa; p.replay_step_iters = a.size // (self.N * 2 * self.n)
        self.params = p
        self.iter_max = int(iter_max)
        self._h = lib().orc_pso_create(C.byref(p))
        if not self._h:
            raise ValueError("orc_pso_create rejected the parameters")

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_pso_destroy(C.c_void_p(self._h)); self._h = None

    def init(self, x0):
        a, pa = _d(x0)
        lib().orc_pso_init(C.c_void_p(self._h), pa)

    def step(self) -> int:
        return int(lib().orc_pso_step(C.c_void_p(self._h)))

    def solve(self, x0):
        self.init(x0)
        for _ in range(self.iter_max):
            if self.step():
                break
        return self.gbest()

    # sharded protocol
    def init_local(self, x0):
        a, pa = _d(x0)
        rec = np.empty(2 + self.n)
        lib().orc_pso_init_local(C.c_void_p(self._h), pa, rec.ctypes.data_as(_dp))
        return rec

    def init_ring_local(self, seed_cost):
        pushes = C.c_int64(0)
        last50 = np.empty(STALL_WINDOW)
        lib().orc_pso_init_ring_local(C.c_void_p(self._h), C.c_double(seed_cost), C.byref(pushes),
                                      last50.ctypes.data_as(_dp))
        return int(pushes.value), last50

    def init_apply(self, records, pushes, last50):
        r, pr = _d(records); l, pl = _d(last50)
        k = np.ascontiguousarray(pushes, dtype=np.int64)
        lib().orc_pso_init_apply(C.c_void_p(self._h), C.c_int(len(k)), pr, k.ctypes.data_as(_ip), pl)

    def step_local(self):
        rec = np.empty(2 + self.n)
        stalled = int(lib().orc_pso_step_local(C.c_void_p(self._h), rec.ctypes.data_as(_dp)))
        return stalled, rec

    def step_apply(self, records):
        r, pr = _d(records)
        lib().orc_pso_step_apply(C.c_void_p(self._h), C.c_int(r.size // (2 + self.n)), pr)

    # readers
    def iters(self): return int(lib().orc_pso_iters(C.c_void_p(self._h)))
    def gbest_cost(self): return float(lib().orc_pso_gbest_cost(C.c_void_p(self._h)))
    def last_improvements(self): return int(lib().orc_pso_last_improvements(C.c_void_p(self._h)))

    def gbest(self):
        g = np.empty(self.n)
        lib().orc_pso_read_gbest(C.c_void_p(self._h), g.ctypes.data_as(_dp))
        return g

    def state(self):
        x = np.empty((self.local, self.n)); v = np.empty_like(x); pb = np.empty_like(x); pbc = np.empty(self.local)
        lib().orc_pso_read_state(C.c_void_p(self._h), x.ctypes.data_as(_dp), v.ctypes.data_as(_dp),
                                 pb.ctypes.data_as(_dp), pbc.ctypes.data_as(_dp))
        return x, v, pb, pbc

    def ring(self):
        r = np.empty(STALL_WINDOW); i = C.c_int64(0)
        lib().orc_pso_read_ring(C.c_void_p(self._h), r.ctypes.data_as(_dp), C.byref(i))
        return r, int(i.value)

    def time_steps(self, iters) -> float:
        return float(lib().orc_time_pso_steps(C.c_void_p(self._h), C.c_int64(iters)))

    def step_mt(self, threads: int) -> int:
        """One synchronous pass with `threads` host threads over the particles (bit-identical to step())."""
        return int(lib().orc_pso_step_mt(C.c_void_p(self._h), int(threads)))

    def time_steps_mt(self, iters, threads: int) -> float:
        return float(lib().orc_time_pso_steps_mt(C.c_void_p(self._h), C.c_int64(iters), int(threads)))


class Ce:
    """Step-granular oracle CrossEntropy (reference cross_entropy.rs:245-306)."""

    def __init__(self, objective_id, x0, sample_size=100, elite_count=10, iter_max=50, tol=1e-6, std_dev=None,
                 mean_divisor=DIV_REFERENCE_TEN, seed=0, replay_normals=None, first_index=0, local_count=None):
        self._keep = []
        x0a, px0 = _d(x0)
        self.n = int(x0a.size)
        self.N = int(sample_size)
        self.k = int(elite_count)
        self.local = self.N if local_count is None else int(local_count)
        p = _CeParams()
        p.n = self.n; p.objective_id = int(objective_id); p.mean_divisor = int(mean_divisor)
        p.sample_size = self.N; p.first_index = int(first_index); p.local_count = self.local
        p.elite_count = self.k; p.iter_max = int(iter_max); p.tol = tol; p.seed = int(seed)
        if std_dev is not None:
            a, pa = _d(std_dev); self._keep.append(a); p.std_dev = pa
        if replay_normals is not None:
            a, pa = _d(replay_normals); assert a.size % (self.N * self.n) == 0
            self._keep.append(a); p.replay_normals = pa; p.replay_iters = a.size // (self.N * self.n)
        self.params = p
        self._h = lib().orc_ce_create(C.byref(p), px0)
        if not self._h:
            raise ValueError("orc_ce_create rejected the parameters")

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_ce_destroy(C.c_void_p(self._h)); self._h = None

    def step(self) -> int:
        return int(lib().orc_ce_step(C.c_void_p(self._h)))

    def solve(self):
        r = 0
        while r == 0:
            r = self.step()
        if r < 0:
            raise FloatingPointError(f"oracle CE: the reference would panic here (status {-r})")
        return self.state()[0]

    def iters(self): return int(lib().orc_ce_iters(C.c_void_p(self._h)))
    def active(self): return bool(lib().orc_ce_active(C.c_void_p(self._h)))

    def state(self):
        mu = np.empty(self.n); sg = np.empty(self.n)
        lib().orc_ce_read_state(C.c_void_p(self._h), mu.ctypes.data_as(_dp), sg.ctypes.data_as(_dp))
        return mu, sg

    def elites(self):
        e = np.empty(self.k, dtype=np.int64)
        lib().orc_ce_read_elites(C.c_void_p(self._h), e.ctypes.data_as(_ip))
        return e

    def costs(self):
        c = np.empty(self.local)
        lib().orc_ce_read_costs(C.c_void_p(self._h), c.ctypes.data_as(_dp))
        return c

    # sharded protocol
    def sample_local(self): lib().orc_ce_sample_local(C.c_void_p(self._h))

    def elite_sum_local(self, T, tie_take):
        s = np.empty(self.n)
        lib().orc_ce_elite_sum_local(C.c_void_p(self._h), C.c_double(T), C.c_int64(tie_take), s.ctypes.data_as(_dp))
        return s

    def elite_sqsum_local(self, mu_new):
        m, pm = _d(mu_new); s = np.empty(self.n)
        lib().orc_ce_elite_sqsum_local(C.c_void_p(self._h), pm, s.ctypes.data_as(_dp))
        return s

    def apply(self, mu_new, sigma_new):
        m, pm = _d(mu_new); s, ps = _d(sigma_new)
        lib().orc_ce_apply(C.c_void_p(self._h), pm, ps)

    def time_steps(self, iters) -> float:
        return float(lib().orc_time_ce_steps(C.c_void_p(self._h), C.c_int64(iters)))
