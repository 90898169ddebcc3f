//! `MonteCarlo` with the surface of eqsolver v0.4.1 (`src/integrators/monte_carlo.rs:85-212`,
//! `OrthotopeRandomIntegrator` `src/integrators/mod.rs:26-99`): the integrand is an [`Objective`] ID and the `rng`
//! argument of `integrate_with_rng` is a Philox seed.
use crate::{Context, Objective, SolverError, SolverResult, VectorType};
use eqsolver_cuda_sys as sys;
use nalgebra::{allocator::Allocator, DefaultAllocator, Dim};
use std::ptr;

/// Default number of samples, eqsolver `src/integrators/mod.rs:18`
pub const DEFAULT_SAMPLE_COUNT: usize = 1000;

/// `MeanVariance<T>` of eqsolver (`src/lib.rs:56-60`)
#[derive(Debug, Copy, Clone)]
pub struct MeanVariance {
    pub mean: f64,
    pub variance: f64,
}

impl MeanVariance {
    pub fn standard_deviation(&self) -> f64 {
        self.variance.sqrt()
    }
}

pub struct MonteCarlo {
    integrand: Objective,
    sample_count: usize,
    context: Option<Context>,
}

impl MonteCarlo {
    pub fn new(integrand: Objective) -> Self {
        Self { integrand, sample_count: DEFAULT_SAMPLE_COUNT, context: None }
    }
    pub fn with_sample_count(&mut self, sample_count: usize) -> &mut Self {
        self.sample_count = sample_count;
        self
    }
    pub fn with_context(&mut self, context: Context) -> &mut Self {
        self.context = Some(context);
        self
    }

    /// `integrate_with_rng(from, to, rng)`, `monte_carlo.rs:181-212`, with a seed in place of the generator.
    pub fn integrate_with_seed<D: Dim>(&self, from: VectorType<D>, to: VectorType<D>, seed: u64) -> SolverResult<MeanVariance>
    where
        DefaultAllocator: Allocator<D>,
    {
        if from.len() != to.len() {
            return Err(SolverError::IncorrectInput { details: "bounds differ in length" });
        }
        let ctx = match &self.context {
            Some(c) => c.clone(),
            None => Context::new(0)?,
        };
        let p = sys::eqs_mc_params {
            n: from.len() as i32,
            integrand_id: self.integrand as i32,
            sample_count: self.sample_count as i64,
            from: from.as_ptr(),
            to: to.as_ptr(),
            seed,
            replay_u: ptr::null(),
        };
        let (mut mean, mut variance) = (0.0f64, 0.0f64);
        let mut report = sys::eqs_report::default();
        let st = unsafe { sys::eqs_mc_integrate(ctx.0 .0, &p, &mut mean, &mut variance, &mut report) };
        ctx.check(st)?;
        Ok(MeanVariance { mean, variance })
    }

    /// `integrate(from, to)`, `integrators/mod.rs:95-98`
    pub fn integrate<D: Dim>(&self, from: VectorType<D>, to: VectorType<D>) -> SolverResult<f64>
    where
        DefaultAllocator: Allocator<D>,
    {
        self.integrate_with_seed(from, to, 0).map(|r| r.mean)
    }
}

/// Defaults of eqsolver's MISER, `src/integrators/miser.rs:13-23`
pub const DEFAULT_MAXIMUM_DEPTH: usize = 5;
pub const DEFAULT_MINIMUM_DIMENSIONAL_SAMPLE_COUNT: usize = 32;
pub const DEFAULT_ALPHA: f64 = 2.;
pub const DEFAULT_DITHER: f64 = 0.05;

/// `MISER` with the surface of eqsolver v0.4.1 (`src/integrators/miser.rs:93-293`); the recursion
/// (`miser_recurse`, `:301-443`) is evaluated level by level by the kernel library.
pub struct MISER {
    integrand: Objective,
    sample_count: usize,
    maximum_depth: usize,
    minimum_dimensional_sample_count: usize,
    alpha: f64,
    dither: f64,
    context: Option<Context>,
}

impl MISER {
    pub fn new(integrand: Objective) -> Self {
        Self {
            integrand,
            sample_count: DEFAULT_SAMPLE_COUNT,
            maximum_depth: DEFAULT_MAXIMUM_DEPTH,
            minimum_dimensional_sample_count: DEFAULT_MINIMUM_DIMENSIONAL_SAMPLE_COUNT,
            alpha: DEFAULT_ALPHA,
            dither: DEFAULT_DITHER,
            context: None,
        }
    }
    pub fn with_sample_count(&mut self, sample_count: usize) -> &mut Self {
        self.sample_count = sample_count;
        self
    }
    pub fn with_minimum_dimensional_sample_count(&mut self, count: usize) -> &mut Self {
        self.minimum_dimensional_sample_count = count;
        self
    }
    pub fn with_maximum_depth(&mut self, maximum_depth: usize) -> &mut Self {
        self.maximum_depth = maximum_depth;
        self
    }
    pub fn with_alpha(&mut self, alpha: f64) -> &mut Self {
        self.alpha = alpha;
        self
    }
    pub fn with_dither(&mut self, dither: f64) -> &mut Self {
        self.dither = dither;
        self
    }
    pub fn with_context(&mut self, context: Context) -> &mut Self {
        self.context = Some(context);
        self
    }

    /// `integrate_with_rng(from, to, rng)`, `miser.rs:277-292`, with a seed in place of the generator.
    pub fn integrate_with_seed<D: Dim>(&self, from: VectorType<D>, to: VectorType<D>, seed: u64) -> SolverResult<MeanVariance>
    where
        DefaultAllocator: Allocator<D>,
    {
        if from.len() != to.len() {
            return Err(SolverError::IncorrectInput { details: "bounds differ in length" });
        }
        let ctx = match &self.context {
            Some(c) => c.clone(),
            None => Context::new(0)?,
        };
        let p = sys::eqs_miser_params {
            n: from.len() as i32,
            integrand_id: self.integrand as i32,
            sample_count: self.sample_count as i64,
            minimum_dimensional_sample_count: self.minimum_dimensional_sample_count as i64,
            maximum_depth: self.maximum_depth as i32,
            reserved: 0,
            alpha: self.alpha,
            dither: self.dither,
            from: from.as_ptr(),
            to: to.as_ptr(),
            seed,
        };
        let (mut mean, mut variance) = (0.0f64, 0.0f64);
        let mut report = sys::eqs_report::default();
        let st = unsafe { sys::eqs_miser_integrate(ctx.0 .0, &p, &mut mean, &mut variance, &mut report) };
        ctx.check(st)?;
        Ok(MeanVariance { mean, variance })
    }

    /// `integrate(from, to)`, `integrators/mod.rs:95-98`
    pub fn integrate<D: Dim>(&self, from: VectorType<D>, to: VectorType<D>) -> SolverResult<f64>
    where
        DefaultAllocator: Allocator<D>,
    {
        self.integrate_with_seed(from, to, 0).map(|r| r.mean)
    }
}
