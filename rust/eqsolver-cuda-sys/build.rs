// Links the in-tree kernel library built by `python -m eqsolver_b200.build` (nvcc, sm_100a).
use std::{env, path::PathBuf};

fn main() {
    let dir = env::var("EQSOLVER_B200_LIB_DIR").map(PathBuf::from).unwrap_or_else(|_| {
        PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../../eqsolver_b200/lib")
    });
    println!("cargo:rustc-link-search=native={}", dir.display());
    println!("cargo:rustc-link-lib=dylib=eqsolver_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir.display());
    println!("cargo:rerun-if-env-changed=EQSOLVER_B200_LIB_DIR");
}
