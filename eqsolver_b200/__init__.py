"""eqsolver_b200 -- B200-native (sm_100a, f64) population step of eqsolver's global optimisers.

Drop-in for ONE path of AzeezDa/eqsolver: `eqsolver::global_optimisers::{ParticleSwarm, CrossEntropy}`
(reference src/solvers/global_optimisers/).  Everything runs in hand-written CUDA kernels behind the C ABI of
include/eqsolver_b200.h; this package is the thin host mirror of the reference's builder API.
"""
from .global_optimisers import (  # noqa: F401
    DEFAULT_ITERMAX, DEFAULT_TOL, BadJacobian, CeRun, Context, CrossEntropy, Exchange, ExternalError, IncorrectInput,
    MaxIterReached, MeanDivisor, NotANumber, Objective, ParticleSwarm, PsoMode, PsoRun, SolverError,
    TypeConversionError, default_context, shard_range, unique_id,
)
from ._ffi import KernelLibraryMissing  # noqa: F401
from .integrators import MISER, MeanVariance, MonteCarlo  # noqa: F401

__all__ = ["ParticleSwarm", "CrossEntropy", "Objective", "PsoMode", "MeanDivisor", "Exchange", "Context",
           "SolverError", "MaxIterReached", "NotANumber", "IncorrectInput", "BadJacobian", "TypeConversionError",
           "ExternalError", "KernelLibraryMissing", "DEFAULT_TOL", "DEFAULT_ITERMAX", "default_context",
           "shard_range", "unique_id", "PsoRun", "CeRun", "MonteCarlo", "MISER", "MeanVariance"]
