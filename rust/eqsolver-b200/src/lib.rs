//! `eqsolver::global_optimisers` on a B200: the reference's builder API (eqsolver v0.4.1,
//! `src/solvers/global_optimisers/`), with the population step running in CUDA behind `eqsolver-cuda-sys`.
pub extern crate nalgebra;
use nalgebra::{allocator::Allocator, DefaultAllocator, Vector};
use std::{ffi::CStr, ptr, sync::Arc};
use thiserror::Error;

pub mod global_optimisers;
pub mod integrators;

/// Default tolerance, eqsolver `src/lib.rs:13`
pub const DEFAULT_TOL: f64 = 1e-6;
/// Default max iterations, eqsolver `src/lib.rs:16`
pub const DEFAULT_ITERMAX: usize = 50;

/// `SolverError` of eqsolver (`src/lib.rs:19-44`), variant for variant.
#[derive(Error, Debug, PartialEq)]
pub enum SolverError {
    #[error("maximum number of iteration, {0}, reached")]
    MaxIterReached(usize),
    #[error("the resulting value is not a number (NaN)")]
    NotANumber,
    #[error("an input value is incorrect: {details}")]
    IncorrectInput { details: &'static str },
    #[error("encountered a singular or ill-defined Jacobian matrix")]
    BadJacobian,
    #[error("failed to convert between types (often between usize and T: Float)")]
    TypeConversionError,
    #[error("encountered an external error: {0}")]
    ExternalError(String),
}

pub type SolverResult<T> = Result<T, SolverError>;
pub(crate) type VectorType<D> = Vector<f64, D, <DefaultAllocator as Allocator<D>>::Buffer<f64>>;

/// The closure `F: Fn(Vector) -> T` of the reference cannot run on the device: objectives are ahead-of-time
/// compiled device functors selected by ID (`eqsolver_b200/csrc/objectives.cuh`).
#[derive(Clone, Copy, Debug, PartialEq, Eq)]
#[repr(i32)]
pub enum Objective {
    Sphere = 0,
    Rosenbrock = 1,
    Rastrigin = 2,
    Ackley = 3,
    ThreeCircles = 4,
    Heavy = 5,
    User = 6,
    /// exp(-sum x^2): the integrand of eqsolver's tests/integrators.rs:11-13
    Gaussian = 7,
    /// sin(cos(x1) + x0): tests/integrators.rs:15-17
    SinCos = 8,
    /// sum over i < 1000 of (-1)^i x^i: the integrand of eqsolver's benchmarks/integrators.rs:69-75
    Series = 9,
}

/// One GPU: stream, device memory arena, NCCL communicator.  Cheap to clone (shared).
#[derive(Clone)]
pub struct Context(pub(crate) Arc<RawContext>);

pub(crate) struct RawContext(pub(crate) *mut eqsolver_cuda_sys::eqs_ctx);
// distinct handles may be used from distinct threads; calls on one context are serialised by the library user
unsafe impl Send for RawContext {}
unsafe impl Sync for RawContext {}

impl Drop for RawContext {
    fn drop(&mut self) {
        unsafe { eqsolver_cuda_sys::eqs_ctx_destroy(self.0) };
    }
}

impl Context {
    pub fn new(device: i32) -> SolverResult<Self> {
        let mut raw = ptr::null_mut();
        let st = unsafe { eqsolver_cuda_sys::eqs_ctx_create(device, &mut raw) };
        if st != eqsolver_cuda_sys::EQS_OK {
            return Err(status_to_error(st, ptr::null()));
        }
        Ok(Context(Arc::new(RawContext(raw))))
    }

    /// Joins the per-box communicator (one process per GPU).  `unique_id` comes from rank 0's
    /// [`Context::unique_id`] and is shipped to the other ranks by the application (MPI, a file, ...).
    pub fn comm_init(&self, rank: i32, world: i32, unique_id: &[u8; 128]) -> SolverResult<()> {
        let st = unsafe { eqsolver_cuda_sys::eqs_ctx_comm_init(self.0 .0, rank, world, unique_id.as_ptr().cast()) };
        self.check(st)
    }

    pub fn unique_id() -> SolverResult<[u8; 128]> {
        let mut id = [0u8; 128];
        let st = unsafe { eqsolver_cuda_sys::eqs_comm_unique_id(id.as_mut_ptr().cast()) };
        if st != eqsolver_cuda_sys::EQS_OK {
            return Err(status_to_error(st, ptr::null()));
        }
        Ok(id)
    }

    pub(crate) fn check(&self, st: i32) -> SolverResult<()> {
        if st == eqsolver_cuda_sys::EQS_OK {
            Ok(())
        } else {
            Err(status_to_error(st, self.0 .0))
        }
    }
}

pub(crate) fn status_to_error(st: i32, ctx: *const eqsolver_cuda_sys::eqs_ctx) -> SolverError {
    let msg = unsafe {
        let p = eqsolver_cuda_sys::eqs_last_error(ctx);
        if p.is_null() { String::new() } else { CStr::from_ptr(p).to_string_lossy().into_owned() }
    };
    match st {
        eqsolver_cuda_sys::EQS_ERR_MAX_ITER => SolverError::MaxIterReached(0),
        eqsolver_cuda_sys::EQS_ERR_NAN => SolverError::NotANumber,
        eqsolver_cuda_sys::EQS_ERR_INCORRECT_INPUT => SolverError::IncorrectInput { details: "rejected by the kernel library" },
        eqsolver_cuda_sys::EQS_ERR_BAD_JACOBIAN => SolverError::BadJacobian,
        eqsolver_cuda_sys::EQS_ERR_TYPE_CONVERSION => SolverError::TypeConversionError,
        _ => SolverError::ExternalError(msg),
    }
}
