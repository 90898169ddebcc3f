//! The integrator doctests of eqsolver v0.4.1 (`src/integrators/monte_carlo.rs:60-80`, `miser.rs:72-83`) on the B200
//! path: the closure becomes an [`Objective`] ID and the `rng` a Philox seed.
use eqsolver_b200::integrators::{MonteCarlo, MISER};
use eqsolver_b200::nalgebra::vector;
use eqsolver_b200::Objective;

fn main() {
    // f(x, y) = sin(cos(y) + x) over the unit square, exact value 0.924846...
    let lower = vector![0., 0.];
    let upper = vector![1., 1.];

    let plain = MonteCarlo::new(Objective::SinCos)
        .with_sample_count(1 << 24)
        .integrate_with_seed(lower, upper, 1729)
        .unwrap();
    println!("Monte Carlo: {} +- {}", plain.mean, (plain.variance / (1u64 << 24) as f64).sqrt());

    let stratified = MISER::new(Objective::SinCos).integrate_with_seed(lower, upper, 1729).unwrap();
    assert!((stratified.mean - 0.924846).abs() <= 0.001); // the bound of eqsolver's own test, tests/integrators.rs:108-116
    println!("MISER: {} (variance {})", stratified.mean, stratified.variance);
}
