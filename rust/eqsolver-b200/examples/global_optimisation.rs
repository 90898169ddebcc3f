// Port of eqsolver's examples/global_optimisation.rs:1-45: only the closure argument changes.
use eqsolver_b200::global_optimisers::{CrossEntropy, ParticleSwarm};
use eqsolver_b200::nalgebra::{vector, SVector};
use eqsolver_b200::Objective;

fn main() {
    const SIZE: usize = 10;
    let guess = vector![80., -81., 82., -83., 84., -85., 86., -87., 88., -89.];

    let lower_bounds = SVector::<f64, SIZE>::repeat(-100.);
    let upper_bounds = SVector::<f64, SIZE>::repeat(100.);
    let solution_pso = ParticleSwarm::new(Objective::Rastrigin, lower_bounds, upper_bounds)
        .unwrap()
        .solve(guess)
        .unwrap();

    let standard_deviations = SVector::<f64, SIZE>::repeat(100.);
    let solution_ce = CrossEntropy::new(Objective::Rastrigin)
        .with_std_dev(standard_deviations)
        .solve(guess)
        .unwrap();

    println!("Guess:    {guess:?}");
    println!("PSO:      {solution_pso:?}");
    println!("CE:       {solution_ce:?}");
}
