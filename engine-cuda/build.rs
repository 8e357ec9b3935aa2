//! Link libmcpower_b200.so.  MCPOWER_B200_LIB_DIR points at the directory holding the library built by
//! `python -m mcpower_b200.build` (nvcc, -gencode arch=compute_100a,code=sm_100a); it defaults to the in-tree location
//! of this repository when the crate is vendored next to it.
use std::env;
use std::path::PathBuf;

fn main() {
    let dir = env::var("MCPOWER_B200_LIB_DIR")
        .map(PathBuf::from)
        .unwrap_or_else(|_| PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../mcpower_b200"));
    println!("cargo:rustc-link-search=native={}", dir.display());
    println!("cargo:rustc-link-lib=dylib=mcpower_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir.display());
    println!("cargo:rerun-if-env-changed=MCPOWER_B200_LIB_DIR");
    println!("cargo:rerun-if-changed=../include/mcpower_b200.h");
}
