// build.rs -- compiles the CUDA sources of this repository into libfacgpu.so for sm_100a and links it.
//
// FAC_CSRC      directory holding fac_api.cu / fac_builder.cpp (default: ../../fuzzy-aho-corasick-rs_b200/csrc)
// FAC_PREBUILT  directory holding an already built libfacgpu.so (skips nvcc)
// NVCC          compiler (default: nvcc on PATH)
//
// --fmad=false / -ffp-contract=off are part of the contract: the reference's f32 arithmetic is never fused
// (rustc does not contract), and similarity bits must match.  The SHA-256 stamp the Python loader checks
// (fac_build_source_hash) is left "unstamped" here: cargo's rerun-if-changed tracking does that job.
use std::{env, path::PathBuf, process::Command};

fn main() {
    let out = PathBuf::from(env::var("OUT_DIR").expect("OUT_DIR"));
    println!("cargo:rerun-if-env-changed=FAC_CSRC");
    println!("cargo:rerun-if-env-changed=FAC_PREBUILT");
    println!("cargo:rerun-if-env-changed=NVCC");
    if let Ok(dir) = env::var("FAC_PREBUILT") {
        println!("cargo:rustc-link-search=native={dir}");
        println!("cargo:rustc-link-lib=dylib=facgpu");
        return;
    }
    let manifest = PathBuf::from(env::var("CARGO_MANIFEST_DIR").expect("CARGO_MANIFEST_DIR"));
    let csrc = env::var("FAC_CSRC")
        .map(PathBuf::from)
        .unwrap_or_else(|_| manifest.join("../../fuzzy-aho-corasick-rs_b200/csrc"));
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "nvcc".to_string());
    let lib = out.join("libfacgpu.so");
    let status = Command::new(&nvcc)
        .args([
            "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-O3", "-lineinfo", "--fmad=false",
            "-Xcompiler", "-fPIC,-ffp-contract=off,-fvisibility=default", "--shared", "-cudart", "shared",
        ])
        .arg(csrc.join("fac_api.cu"))
        .arg(csrc.join("fac_builder.cpp"))
        .arg("-o")
        .arg(&lib)
        .status()
        .unwrap_or_else(|e| panic!("cannot run {nvcc}: {e}"));
    assert!(status.success(), "nvcc failed building libfacgpu.so");
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=facgpu");
    println!("cargo:rustc-link-lib=dylib=cudart");
    println!("cargo:rerun-if-changed={}", csrc.display());
    println!("cargo:rerun-if-changed={}", manifest.join("../../include/fac.h").display());
}
