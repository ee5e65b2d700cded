// build.rs of pandrs with the B200 path (not compiled in this repository: the image has no cargo / rustc).
// Gated by pandrs's existing `cuda` feature (Cargo.toml); replaces the cudarc shell of src/gpu for the groupby / join path.
//   PANDRS_B200_ROOT = checkout of this repository; `make -C $PANDRS_B200_ROOT/pandrs_b200/csrc` builds the library
//   (nvcc -gencode arch=compute_100a,code=sm_100a, no other architecture, no CPU fallback).
use std::{env, path::PathBuf, process::Command};

fn main() {
    println!("cargo:rerun-if-env-changed=PANDRS_B200_ROOT");
    if env::var("CARGO_FEATURE_CUDA").is_err() {
        return;
    }
    let root = PathBuf::from(env::var("PANDRS_B200_ROOT").expect("PANDRS_B200_ROOT must point at the pandrs_b200 checkout"));
    let csrc = root.join("pandrs_b200").join("csrc");
    let status = Command::new("make").arg("-C").arg(&csrc).arg("-j8").status().expect("make (nvcc) not found");
    assert!(status.success(), "building libpandrs_b200.so failed");
    println!("cargo:rustc-link-search=native={}", root.join("pandrs_b200").join("lib").display());
    println!("cargo:rustc-link-lib=dylib=pandrs_b200");
    println!("cargo:rustc-cfg=cuda_available"); // the cfg pandrs already tests (src/lib.rs:162-163)
    println!("cargo:rerun-if-changed={}", root.join("include").join("pandrs_b200.h").display());
    // src/gpu/b200_ffi.rs = rust/ffi.rs of the checkout (generated from the header by tools/gen_rust_ffi.py)
}
