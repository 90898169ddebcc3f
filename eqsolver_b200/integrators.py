"""Host-side mirror of `eqsolver::integrators::{MonteCarlo, MISER}` (reference src/integrators/monte_carlo.rs:85-212,
src/integrators/miser.rs:93-293) and of the
`OrthotopeRandomIntegrator` trait (src/integrators/mod.rs:26-99), for the integrands compiled into the kernel library.

`integrate_with_rng(from, to, rng)` becomes `integrate_with_seed(from, to, seed)`: the `rng` argument is a Philox key.
"""
from __future__ import annotations

import ctypes as C
from typing import NamedTuple, Optional

import numpy as np

from . import _ffi
from ._ffi import McParams, MiserParams, Report
from .global_optimisers import Context, IncorrectInput, Objective, _arr, _p, default_context

DEFAULT_SAMPLE_COUNT = 1000  # integrators/mod.rs:18
DEFAULT_MAXIMUM_DEPTH = 5                       # miser.rs:13
DEFAULT_MINIMUM_DIMENSIONAL_SAMPLE_COUNT = 32   # miser.rs:17
DEFAULT_ALPHA = 2.0                             # miser.rs:20
DEFAULT_DITHER = 0.05                           # miser.rs:23


class MeanVariance(NamedTuple):
    """`MeanVariance<T>` of the reference, src/lib.rs:56-60."""
    mean: float
    variance: float

    def standard_deviation(self) -> float:  # lib.rs:133-135
        return float(np.sqrt(self.variance))


class MonteCarlo:
    def __init__(self, integrand):
        self.integrand = Objective(int(integrand))
        self.sample_count = DEFAULT_SAMPLE_COUNT
        self.context: Optional[Context] = None
        self.last_report: Optional[Report] = None

    @classmethod
    def new(cls, integrand) -> "MonteCarlo":  # monte_carlo.rs:85-91
        return cls(integrand)

    def with_sample_count(self, v: int):  # :111-114
        self.sample_count = int(v)
        return self

    def with_context(self, ctx: Context):
        self.context = ctx
        return self

    def integrate_with_seed(self, lower, upper, seed: int = 0, replay_u=None) -> MeanVariance:
        """`integrate_with_rng`, monte_carlo.rs:164-212 (scalars or vectors for the bounds)."""
        ctx = self.context or default_context()
        lo = _arr(np.atleast_1d(lower))
        hi = _arr(np.atleast_1d(upper), lo.size)
        p = McParams()
        p.n = lo.size
        p.integrand_id = int(self.integrand)
        p.sample_count = self.sample_count
        p.from_ = _p(lo)
        p.to = _p(hi)
        p.seed = int(seed)
        r = None
        if replay_u is not None:
            r = _arr(replay_u, self.sample_count * lo.size)
            p.replay_u = _p(r)
        mean, var = C.c_double(0), C.c_double(0)
        rep = Report()
        ctx.check(_ffi.lib().eqs_mc_integrate(ctx._h, C.byref(p), C.byref(mean), C.byref(var), C.byref(rep)))
        self.last_report = rep
        return MeanVariance(mean.value, var.value)

    def integrate_to_mean_variance(self, lower, upper) -> MeanVariance:  # integrators/mod.rs:75-77
        return self.integrate_with_seed(lower, upper, 0)

    def integrate(self, lower, upper) -> float:  # integrators/mod.rs:95-98
        return self.integrate_to_mean_variance(lower, upper).mean


class MISER:
    """Recursive stratified sampling, `MISER::new(f).with_*(..)` of the reference (miser.rs:93-268)."""

    def __init__(self, integrand):
        self.integrand = Objective(int(integrand))
        self.sample_count = DEFAULT_SAMPLE_COUNT
        self.maximum_depth = DEFAULT_MAXIMUM_DEPTH
        self.minimum_dimensional_sample_count = DEFAULT_MINIMUM_DIMENSIONAL_SAMPLE_COUNT
        self.alpha = DEFAULT_ALPHA
        self.dither = DEFAULT_DITHER
        self.context: Optional[Context] = None
        self.last_report: Optional[Report] = None

    @classmethod
    def new(cls, integrand) -> "MISER":  # miser.rs:109-120
        return cls(integrand)

    def with_sample_count(self, v: int): self.sample_count = int(v); return self                       # :141-144
    def with_minimum_dimensional_sample_count(self, v: int):                                            # :181-187
        self.minimum_dimensional_sample_count = int(v); return self
    def with_maximum_depth(self, v: int): self.maximum_depth = int(v); return self                     # :206-209
    def with_alpha(self, v: float): self.alpha = float(v); return self                                 # :236-239
    def with_dither(self, v: float): self.dither = float(v); return self                               # :264-267
    def with_context(self, ctx: Context): self.context = ctx; return self

    def integrate_with_seed(self, lower, upper, seed: int = 0) -> MeanVariance:
        """`integrate_with_rng`, miser.rs:277-292."""
        ctx = self.context or default_context()
        lo = _arr(np.atleast_1d(lower))
        hi = _arr(np.atleast_1d(upper), lo.size)
        p = MiserParams()
        p.n = lo.size
        p.integrand_id = int(self.integrand)
        p.sample_count = self.sample_count
        p.minimum_dimensional_sample_count = self.minimum_dimensional_sample_count
        p.maximum_depth = self.maximum_depth
        p.alpha, p.dither = self.alpha, self.dither
        p.from_, p.to = _p(lo), _p(hi)
        p.seed = int(seed)
        mean, var = C.c_double(0), C.c_double(0)
        rep = Report()
        ctx.check(_ffi.lib().eqs_miser_integrate(ctx._h, C.byref(p), C.byref(mean), C.byref(var), C.byref(rep)))
        self.last_report = rep
        return MeanVariance(mean.value, var.value)

    def integrate_to_mean_variance(self, lower, upper) -> MeanVariance:  # integrators/mod.rs:75-77
        return self.integrate_with_seed(lower, upper, 0)

    def integrate(self, lower, upper) -> float:  # integrators/mod.rs:95-98
        return self.integrate_to_mean_variance(lower, upper).mean
