//! `ParticleSwarm` and `CrossEntropy` with the signatures of eqsolver v0.4.1
//! (`particle_swarm.rs:119-123,179-333,356`; `cross_entropy.rs:75,108-223,245`).
use crate::{Context, Objective, SolverError, SolverResult, VectorType, DEFAULT_ITERMAX, DEFAULT_TOL};
use eqsolver_cuda_sys as sys;
use nalgebra::{allocator::Allocator, DefaultAllocator, Dim};
use std::ptr;

const DEFAULT_INERTIA_WEIGHT: f64 = 0.5; // particle_swarm.rs:8
const DEFAULT_COGNITIVE_COEFFICIENT: f64 = 1.0; // :9
const DEFAULT_SOCIAL_COEFFICIENT: f64 = 1.0; // :10
const DEFAULT_PARTICLE_COUNT: usize = 100; // :11
const DEFAULT_SAMPLE_SIZE: usize = 100; // cross_entropy.rs:8
const DEFAULT_IMPORTANCE_SELECTION_SIZE: usize = 10; // :9

/// How the swarm's global best moves inside one iteration.
#[derive(Clone, Copy, Debug, PartialEq, Eq)]
pub enum PsoMode {
    /// the reference: gbest is updated inside the particle loop (`particle_swarm.rs:420-425`)
    SequentialExact = 0,
    /// one arg-min and gbest update per iteration (the throughput mode)
    Synchronous = 1,
}

/// Divisor of the elite mean in `CrossEntropy`.
#[derive(Clone, Copy, Debug, PartialEq, Eq)]
pub enum MeanDivisor {
    /// the literal 10 of `cross_entropy.rs:287`
    ReferenceTen = 0,
    /// the importance selection size k
    EliteCount = 1,
}

pub struct ParticleSwarm<D: Dim>
where
    DefaultAllocator: Allocator<D>,
{
    objective: Objective,
    lower_bounds: VectorType<D>,
    upper_bounds: VectorType<D>,
    inertia_weight: f64,
    cognitive_coefficient: f64,
    social_coefficient: f64,
    tolerance: f64,
    iter_max: usize,
    particle_count: usize,
    seed: u64,
    mode: PsoMode,
    context: Option<Context>,
}

impl<D: Dim> ParticleSwarm<D>
where
    DefaultAllocator: Allocator<D>,
{
    /// `ParticleSwarm::new(f, lower_bounds, upper_bounds)`, `particle_swarm.rs:119-153`, with `f` an [`Objective`].
    /// Fails like the reference (`SolverError::ExternalError`) when a bounds pair does not fulfil `from <= to`.
    pub fn new(objective: Objective, lower_bounds: VectorType<D>, upper_bounds: VectorType<D>) -> SolverResult<Self> {
        let s = Self {
            objective,
            lower_bounds,
            upper_bounds,
            inertia_weight: DEFAULT_INERTIA_WEIGHT,
            cognitive_coefficient: DEFAULT_COGNITIVE_COEFFICIENT,
            social_coefficient: DEFAULT_SOCIAL_COEFFICIENT,
            tolerance: DEFAULT_TOL,
            iter_max: DEFAULT_ITERMAX,
            particle_count: DEFAULT_PARTICLE_COUNT,
            seed: 0,
            mode: PsoMode::SequentialExact,
            context: None,
        };
        match unsafe { sys::eqs_pso_validate(&s.params()) } {
            sys::EQS_OK => Ok(s),
            sys::EQS_ERR_EXTERNAL => Err(SolverError::ExternalError(
                "low > high (or equal if exclusive) in uniform distribution".to_string(),
            )),
            _ => Err(SolverError::IncorrectInput { details: "dimension, objective or bounds rejected" }),
        }
    }

    pub fn with_inertia_weight(&mut self, inertia_weight: f64) -> &mut Self {
        self.inertia_weight = inertia_weight;
        self
    }
    pub fn with_cognitive_coefficient(&mut self, cognitive_coefficient: f64) -> &mut Self {
        self.cognitive_coefficient = cognitive_coefficient;
        self
    }
    pub fn with_social_coefficient(&mut self, social_coefficient: f64) -> &mut Self {
        self.social_coefficient = social_coefficient;
        self
    }
    pub fn with_tol(&mut self, tol: f64) -> &mut Self {
        self.tolerance = tol;
        self
    }
    pub fn with_particle_count(&mut self, particle_count: usize) -> &mut Self {
        self.particle_count = particle_count;
        self
    }
    pub fn with_iter_max(&mut self, iter_max: usize) -> &mut Self {
        self.iter_max = iter_max;
        self
    }
    /// Philox key (the reference's `rand::rng()` cannot be seeded).
    pub fn with_seed(&mut self, seed: u64) -> &mut Self {
        self.seed = seed;
        self
    }
    pub fn with_mode(&mut self, mode: PsoMode) -> &mut Self {
        self.mode = mode;
        self
    }
    pub fn with_context(&mut self, context: Context) -> &mut Self {
        self.context = Some(context);
        self
    }

    fn params(&self) -> sys::eqs_pso_params {
        sys::eqs_pso_params {
            n: self.lower_bounds.len() as i32,
            objective_id: self.objective as i32,
            mode: self.mode as i32,
            exchange: sys::EQS_EXCHANGE_AUTO,
            particle_count: self.particle_count as i64,
            iter_max: self.iter_max as i64,
            tol: self.tolerance,
            w: self.inertia_weight,
            c1: self.cognitive_coefficient,
            c2: self.social_coefficient,
            lower: self.lower_bounds.as_ptr(),
            upper: self.upper_bounds.as_ptr(),
            seed: self.seed,
            replay_init: ptr::null(),
            virtual_rank: 0,
            virtual_world: 1,
        }
    }

    /// `solve` with an explicit generator state: the optimisers' twin of
    /// `OrthotopeRandomIntegrator::integrate_with_rng` (`integrators/mod.rs:56-61`).
    pub fn solve_with_seed(&mut self, x0: VectorType<D>, seed: u64) -> SolverResult<VectorType<D>> {
        self.seed = seed;
        self.solve(x0)
    }

    /// `solve` on caller-supplied draws instead of a generator (the replay mode of the parity tests as a feature).
    /// `init_draws` `[N][2][n]`: u in [0, 1) for the start positions, then the velocities, which the samplers of
    /// `particle_swarm.rs:367,373` scale into their ranges; `step_draws` `[passes][N][n][2]`: (r_p, r_g) of `:407-408`.
    /// Runs `min(passes, iter_max)` passes, fewer when the stall test (`:433-441`) fires.
    pub fn solve_with_draws(&self, x0: VectorType<D>, init_draws: &[f64], step_draws: &[f64]) -> SolverResult<VectorType<D>> {
        let n = self.lower_bounds.len();
        let per_pass = self.particle_count * n * 2;
        if x0.len() != n || init_draws.len() != per_pass || per_pass == 0 || step_draws.len() % per_pass != 0 {
            return Err(SolverError::IncorrectInput { details: "init_draws: particle_count * 2n draws; step_draws: whole passes of particle_count * n * 2 draws" });
        }
        let ctx = match &self.context {
            Some(c) => c.clone(),
            None => Context::new(0)?,
        };
        let mut p = self.params();
        p.replay_init = init_draws.as_ptr();
        let mut h: *mut sys::eqs_pso = ptr::null_mut();
        ctx.check(unsafe { sys::eqs_pso_create(ctx.0 .0, &p, &mut h) })?;
        let mut out = x0.clone();
        let mut st = unsafe { sys::eqs_pso_init(h, x0.as_ptr()) };
        let passes = (step_draws.len() / per_pass).min(self.iter_max);
        for t in 0..passes {
            if st != sys::EQS_OK {
                break;
            }
            st = unsafe { sys::eqs_pso_step(h, 1, step_draws[t * per_pass..].as_ptr()) };
        }
        if st == sys::EQS_OK {
            st = unsafe {
                sys::eqs_pso_read_gbest(h, out.as_mut_ptr(), ptr::null_mut(), ptr::null_mut(), ptr::null_mut(), ptr::null_mut())
            };
        }
        let res = ctx.check(st); // reads the message before the handle goes away
        unsafe { sys::eqs_pso_destroy(h) };
        res?;
        Ok(out)
    }

    /// `solve(x0)`, `particle_swarm.rs:356-431`: the global best position.
    pub fn solve(&self, x0: VectorType<D>) -> SolverResult<VectorType<D>> {
        let ctx = match &self.context {
            Some(c) => c.clone(),
            None => Context::new(0)?,
        };
        let mut out = x0.clone();
        let mut report = sys::eqs_report::default();
        let st = unsafe { sys::eqs_pso_solve(ctx.0 .0, &self.params(), x0.as_ptr(), out.as_mut_ptr(), &mut report) };
        ctx.check(st)?;
        Ok(out)
    }
}

pub struct CrossEntropy<D: Dim>
where
    DefaultAllocator: Allocator<D>,
{
    objective: Objective,
    std_dev: Option<VectorType<D>>,
    tolerance: f64,
    iter_max: usize,
    sample_size: usize,
    importance_selection_size: usize,
    seed: u64,
    mean_divisor: MeanDivisor,
    context: Option<Context>,
}

impl<D: Dim> CrossEntropy<D>
where
    DefaultAllocator: Allocator<D>,
{
    /// `CrossEntropy::new(f)`, `cross_entropy.rs:75-85`, with `f` an [`Objective`].
    pub fn new(objective: Objective) -> Self {
        Self {
            objective,
            std_dev: None,
            tolerance: DEFAULT_TOL,
            iter_max: DEFAULT_ITERMAX,
            sample_size: DEFAULT_SAMPLE_SIZE,
            importance_selection_size: DEFAULT_IMPORTANCE_SELECTION_SIZE,
            seed: 0,
            mean_divisor: MeanDivisor::ReferenceTen,
            context: None,
        }
    }

    pub fn with_tol(&mut self, tol: f64) -> &mut Self {
        self.tolerance = tol;
        self
    }
    pub fn with_iter_max(&mut self, iter_max: usize) -> &mut Self {
        self.iter_max = iter_max;
        self
    }
    pub fn with_sample_size(&mut self, sample_size: usize) -> &mut Self {
        self.sample_size = sample_size;
        self
    }
    pub fn with_importance_selection_size(&mut self, importance_selection_size: usize) -> &mut Self {
        self.importance_selection_size = importance_selection_size;
        self
    }
    pub fn with_std_dev(&mut self, std_dev: VectorType<D>) -> &mut Self {
        self.std_dev = Some(std_dev);
        self
    }
    pub fn with_seed(&mut self, seed: u64) -> &mut Self {
        self.seed = seed;
        self
    }
    pub fn with_mean_divisor(&mut self, mean_divisor: MeanDivisor) -> &mut Self {
        self.mean_divisor = mean_divisor;
        self
    }
    pub fn with_context(&mut self, context: Context) -> &mut Self {
        self.context = Some(context);
        self
    }

    /// `solve` with an explicit generator state (see `ParticleSwarm::solve_with_seed`).
    pub fn solve_with_seed(&mut self, x0: VectorType<D>, seed: u64) -> SolverResult<VectorType<D>> {
        self.seed = seed;
        self.solve(x0)
    }

    /// `solve` on caller-supplied standard normals `normals` `[passes][N][n]` (the z of `Normal::sample`,
    /// `cross_entropy.rs:270`): runs `min(passes, iter_max - 1)` passes (`iter` starts at 1, `:253`) or stops on `tol`.
    pub fn solve_with_draws(&self, x0: VectorType<D>, normals: &[f64]) -> SolverResult<VectorType<D>> {
        let per_pass = self.sample_size * x0.len();
        if per_pass == 0 || normals.len() % per_pass != 0 {
            return Err(SolverError::IncorrectInput { details: "normals must hold whole passes of sample_size * n draws" });
        }
        let ctx = match &self.context {
            Some(c) => c.clone(),
            None => Context::new(0)?,
        };
        let p = self.params(x0.len())?;
        let mut h: *mut sys::eqs_ce = ptr::null_mut();
        ctx.check(unsafe { sys::eqs_ce_create(ctx.0 .0, &p, x0.as_ptr(), &mut h) })?;
        let mut out = x0.clone();
        let mut st = sys::EQS_OK;
        let passes = (normals.len() / per_pass).min(self.iter_max.saturating_sub(1));
        for t in 0..passes {
            if st != sys::EQS_OK {
                break;
            }
            st = unsafe { sys::eqs_ce_step(h, 1, normals[t * per_pass..].as_ptr()) };
        }
        if st == sys::EQS_OK {
            st = unsafe { sys::eqs_ce_read_state(h, out.as_mut_ptr(), ptr::null_mut(), ptr::null_mut(), ptr::null_mut()) };
        }
        let res = ctx.check(st);
        unsafe { sys::eqs_ce_destroy(h) };
        res?;
        Ok(out)
    }

    fn params(&self, n: usize) -> SolverResult<sys::eqs_ce_params> {
        if let Some(s) = &self.std_dev {
            if s.len() != n {
                return Err(SolverError::IncorrectInput { details: "std_dev and x0 differ in length" });
            }
        }
        Ok(sys::eqs_ce_params {
            n: n as i32,
            objective_id: self.objective as i32,
            mean_divisor: self.mean_divisor as i32,
            exchange: sys::EQS_EXCHANGE_AUTO,
            sample_size: self.sample_size as i64,
            elite_count: self.importance_selection_size as i64,
            iter_max: self.iter_max as i64,
            tol: self.tolerance,
            std_dev: self.std_dev.as_ref().map_or(ptr::null(), |s| s.as_ptr()),
            seed: self.seed,
            virtual_rank: 0,
            virtual_world: 1,
        })
    }

    /// `solve(x0)`, `cross_entropy.rs:245-306`: the final mean.  Where the reference panics (NaN cost, sampling with
    /// a non-finite standard deviation) this returns `Err(SolverError::NotANumber)`.
    pub fn solve(&self, x0: VectorType<D>) -> SolverResult<VectorType<D>> {
        let ctx = match &self.context {
            Some(c) => c.clone(),
            None => Context::new(0)?,
        };
        let p = self.params(x0.len())?;
        let mut out = x0.clone();
        let mut report = sys::eqs_report::default();
        let st = unsafe { sys::eqs_ce_solve(ctx.0 .0, &p, x0.as_ptr(), out.as_mut_ptr(), &mut report) };
        ctx.check(st)?;
        Ok(out)
    }
}
