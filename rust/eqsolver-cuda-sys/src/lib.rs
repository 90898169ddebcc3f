//! Raw bindings of `include/eqsolver_b200.h` (ABI version 1).  One declaration per C symbol, nothing else.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_double, c_int, c_void};

pub const EQS_ABI_VERSION: c_int = 1;
pub const EQS_STALL_WINDOW: usize = 50;

// eqs_status == SolverError variants of eqsolver (src/lib.rs:19-44); 0 = Ok
pub const EQS_OK: c_int = 0;
pub const EQS_ERR_MAX_ITER: c_int = 1;
pub const EQS_ERR_NAN: c_int = 2;
pub const EQS_ERR_INCORRECT_INPUT: c_int = 3;
pub const EQS_ERR_BAD_JACOBIAN: c_int = 4;
pub const EQS_ERR_TYPE_CONVERSION: c_int = 5;
pub const EQS_ERR_EXTERNAL: c_int = 6;

// eqs_objective
pub const EQS_OBJ_SPHERE: i32 = 0;
pub const EQS_OBJ_ROSENBROCK: i32 = 1;
pub const EQS_OBJ_RASTRIGIN: i32 = 2;
pub const EQS_OBJ_ACKLEY: i32 = 3;
pub const EQS_OBJ_THREE_CIRCLES: i32 = 4;
pub const EQS_OBJ_HEAVY: i32 = 5;
pub const EQS_OBJ_USER: i32 = 6;
pub const EQS_OBJ_GAUSSIAN: i32 = 7;
pub const EQS_OBJ_SINCOS: i32 = 8;
pub const EQS_OBJ_SERIES: i32 = 9;

pub const EQS_PSO_SEQUENTIAL_EXACT: i32 = 0;
pub const EQS_PSO_SYNCHRONOUS: i32 = 1;
pub const EQS_CE_DIV_REFERENCE_TEN: i32 = 0;
pub const EQS_CE_DIV_ELITE_COUNT: i32 = 1;
pub const EQS_EXCHANGE_AUTO: i32 = 0;
pub const EQS_EXCHANGE_MANUAL: i32 = 1;

#[repr(C)]
pub struct eqs_ctx {
    _private: [u8; 0],
}
#[repr(C)]
pub struct eqs_pso {
    _private: [u8; 0],
}
#[repr(C)]
pub struct eqs_ce {
    _private: [u8; 0],
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct eqs_pso_params {
    pub n: i32,
    pub objective_id: i32,
    pub mode: i32,
    pub exchange: i32,
    pub particle_count: i64,
    pub iter_max: i64,
    pub tol: c_double,
    pub w: c_double,
    pub c1: c_double,
    pub c2: c_double,
    pub lower: *const c_double,
    pub upper: *const c_double,
    pub seed: u64,
    pub replay_init: *const c_double,
    pub virtual_rank: i32,
    pub virtual_world: i32,
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct eqs_ce_params {
    pub n: i32,
    pub objective_id: i32,
    pub mean_divisor: i32,
    pub exchange: i32,
    pub sample_size: i64,
    pub elite_count: i64,
    pub iter_max: i64,
    pub tol: c_double,
    pub std_dev: *const c_double,
    pub seed: u64,
    pub virtual_rank: i32,
    pub virtual_world: i32,
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct eqs_mc_params {
    pub n: i32,
    pub integrand_id: i32,
    pub sample_count: i64,
    pub from: *const c_double,
    pub to: *const c_double,
    pub seed: u64,
    pub replay_u: *const c_double,
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct eqs_miser_params {
    pub n: i32,
    pub integrand_id: i32,
    pub sample_count: i64,
    pub minimum_dimensional_sample_count: i64,
    pub maximum_depth: i32,
    pub reserved: i32,
    pub alpha: c_double,
    pub dither: c_double,
    pub from: *const c_double,
    pub to: *const c_double,
    pub seed: u64,
}

#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct eqs_report {
    pub iters_run: i64,
    pub evals: i64,
    pub best_cost: c_double,
    pub stalled: i32,
    pub world: i32,
    pub kernel_launches: i64,
    pub init_ms: c_double,
    pub loop_ms: c_double,
}

extern "C" {
    pub fn eqs_abi_version() -> c_int;
    pub fn eqs_ctx_create(device: c_int, out: *mut *mut eqs_ctx) -> c_int;
    pub fn eqs_ctx_destroy(ctx: *mut eqs_ctx) -> c_int;
    pub fn eqs_last_error(ctx: *const eqs_ctx) -> *const c_char;
    pub fn eqs_ctx_sm_count(ctx: *const eqs_ctx, out: *mut c_int) -> c_int;
    pub fn eqs_comm_unique_id(out128: *mut c_void) -> c_int;
    pub fn eqs_ctx_comm_init(ctx: *mut eqs_ctx, rank: c_int, world: c_int, id128: *const c_void) -> c_int;
    pub fn eqs_ctx_rank(ctx: *const eqs_ctx, rank: *mut c_int, world: *mut c_int) -> c_int;
    pub fn eqs_shard_range(count: i64, world: c_int, rank: c_int, first: *mut i64, local: *mut i64) -> c_int;
    pub fn eqs_record_doubles(n: i32) -> i64;

    pub fn eqs_objective_eval(ctx: *mut eqs_ctx, objective_id: i32, x: *const c_double, n: i32, count: i64,
                              out: *mut c_double) -> c_int;
    pub fn eqs_philox_blocks(ctx: *mut eqs_ctx, ctr: *const u32, key: *const u32, count: i64, out: *mut u32) -> c_int;
    pub fn eqs_draws(ctx: *mut eqs_ctx, kind: i32, seed: u64, t: i64, first: i64, count: i64, n: i32,
                     a: *mut c_double, b: *mut c_double) -> c_int;
    pub fn eqs_device_cos(ctx: *mut eqs_ctx, x: *const c_double, count: i64, out: *mut c_double) -> c_int;
    pub fn eqs_measure_fp64(ctx: *mut eqs_ctx, tflops: *mut c_double) -> c_int;
    pub fn eqs_measure_pattern(ctx: *mut eqs_ctx, n: i32, particles: i64, gbps: *mut c_double) -> c_int;
    pub fn eqs_measure_copy(ctx: *mut eqs_ctx, bytes: i64, gbps: *mut c_double) -> c_int;

    pub fn eqs_pso_validate(p: *const eqs_pso_params) -> c_int;
    pub fn eqs_pso_solve(ctx: *mut eqs_ctx, p: *const eqs_pso_params, x0: *const c_double, out_x: *mut c_double,
                         report: *mut eqs_report) -> c_int;
    pub fn eqs_pso_create(ctx: *mut eqs_ctx, p: *const eqs_pso_params, out: *mut *mut eqs_pso) -> c_int;
    pub fn eqs_pso_destroy(h: *mut eqs_pso) -> c_int;
    pub fn eqs_pso_init(h: *mut eqs_pso, x0: *const c_double) -> c_int;
    pub fn eqs_pso_step(h: *mut eqs_pso, iters: i64, replay_step: *const c_double) -> c_int;
    pub fn eqs_pso_sync(h: *mut eqs_pso) -> c_int;
    pub fn eqs_pso_last_step_ms(h: *mut eqs_pso, ms: *mut c_double, kernel_launches: *mut i64) -> c_int;
    pub fn eqs_pso_read_gbest(h: *mut eqs_pso, g: *mut c_double, cost: *mut c_double, iters_run: *mut i64,
                              stalled: *mut i32, improvements: *mut i64) -> c_int;
    pub fn eqs_pso_read_state(h: *mut eqs_pso, x: *mut c_double, v: *mut c_double, pbest: *mut c_double,
                              pbest_cost: *mut c_double) -> c_int;
    pub fn eqs_pso_read_ring(h: *mut eqs_pso, ring: *mut c_double, ring_i: *mut i64) -> c_int;
    pub fn eqs_pso_local_range(h: *mut eqs_pso, first: *mut i64, local: *mut i64) -> c_int;
    pub fn eqs_pso_local_indices(h: *mut eqs_pso, gidx: *mut i64) -> c_int;
    pub fn eqs_pso_init_local(h: *mut eqs_pso, x0: *const c_double, phase: i32) -> c_int;
    pub fn eqs_pso_init_apply(h: *mut eqs_pso) -> c_int;
    pub fn eqs_pso_step_local(h: *mut eqs_pso, replay_step: *const c_double) -> c_int;
    pub fn eqs_pso_step_apply(h: *mut eqs_pso) -> c_int;
    pub fn eqs_pso_get_record(h: *mut eqs_pso, record: *mut c_double) -> c_int;
    pub fn eqs_pso_set_records(h: *mut eqs_pso, records: *const c_double) -> c_int;

    pub fn eqs_ce_validate(p: *const eqs_ce_params) -> c_int;
    pub fn eqs_ce_solve(ctx: *mut eqs_ctx, p: *const eqs_ce_params, x0: *const c_double, out_mu: *mut c_double,
                        report: *mut eqs_report) -> c_int;
    pub fn eqs_ce_create(ctx: *mut eqs_ctx, p: *const eqs_ce_params, x0: *const c_double, out: *mut *mut eqs_ce) -> c_int;
    pub fn eqs_ce_destroy(h: *mut eqs_ce) -> c_int;
    pub fn eqs_ce_step(h: *mut eqs_ce, iters: i64, replay_normals: *const c_double) -> c_int;
    pub fn eqs_ce_sync(h: *mut eqs_ce) -> c_int;
    pub fn eqs_ce_last_step_ms(h: *mut eqs_ce, ms: *mut c_double, kernel_launches: *mut i64) -> c_int;
    pub fn eqs_ce_read_state(h: *mut eqs_ce, mu: *mut c_double, sigma: *mut c_double, iters_run: *mut i64,
                             active: *mut i32) -> c_int;
    pub fn eqs_ce_read_elites(h: *mut eqs_ce, idx: *mut i64, count: *mut i64) -> c_int;
    pub fn eqs_ce_read_costs(h: *mut eqs_ce, costs: *mut c_double) -> c_int;
    pub fn eqs_ce_local_range(h: *mut eqs_ce, first: *mut i64, local: *mut i64) -> c_int;
    pub fn eqs_ce_phase(h: *mut eqs_ce, phase: i32, replay_normals: *const c_double) -> c_int;
    pub fn eqs_ce_exchange_words(h: *mut eqs_ce, phase: i32) -> i64;
    pub fn eqs_ce_get_exchange(h: *mut eqs_ce, phase: i32, words: *mut c_void) -> c_int;
    pub fn eqs_ce_set_exchange(h: *mut eqs_ce, phase: i32, words: *const c_void) -> c_int;

    pub fn eqs_mc_validate(p: *const eqs_mc_params) -> c_int;
    pub fn eqs_mc_integrate(ctx: *mut eqs_ctx, p: *const eqs_mc_params, mean: *mut c_double, variance: *mut c_double,
                            report: *mut eqs_report) -> c_int;

    pub fn eqs_miser_validate(p: *const eqs_miser_params) -> c_int;
    pub fn eqs_miser_integrate(ctx: *mut eqs_ctx, p: *const eqs_miser_params, mean: *mut c_double, variance: *mut c_double,
                               report: *mut eqs_report) -> c_int;
}
