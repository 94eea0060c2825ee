// build.rs — compile the CUDA engine for sm_100a with nvcc and link it (BASELINE.json north_star: "a thin extern "C" FFI
// layer whose kernels build.rs compiles with nvcc for sm_100a").  SOURCE ONLY: no rustc/cargo in this image.
use std::{env, path::PathBuf, process::Command};

fn main() {
    let root = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("..");
    let csrc = root.join("collider-rs_b200").join("csrc");
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let so = out.join("libcollider_b200.so");
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "nvcc".into());
    // -fmad=false: every f64 product and sum is rounded separately, like the reference's SSE2 code (DESIGN.md "f64 contract")
    let status = Command::new(&nvcc)
        .args(["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-fmad=false",
               "-Xcompiler", "-fPIC", "-shared", "-cudart", "static", "-o"])
        .arg(&so)
        .arg(csrc.join("engine.cu"))
        .status()
        .expect("nvcc not found (set NVCC)");
    assert!(status.success(), "nvcc failed");
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=collider_b200");
    println!("cargo:rerun-if-changed={}", csrc.display());
    println!("cargo:rerun-if-changed={}", root.join("include").join("collider_b200.h").display());
}
