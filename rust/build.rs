// build.rs for the `gpu` feature of krabmaga: compiles the CUDA library behind
// include/krabgpu.h with nvcc for B200 (sm_100a) and links it.
// NOT COMPILED IN THIS REPOSITORY'S CI: the build image has no cargo/rustc.
use std::{env, path::PathBuf, process::Command};

fn main() {
    if env::var("CARGO_FEATURE_GPU").is_err() {
        return;
    }
    let root = PathBuf::from(env::var("KRABGPU_ROOT").unwrap_or_else(|_| "../".into()));
    let csrc = root.join("krabmaga_b200/csrc");
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let lib = out.join("libkrabgpu.so");
    let srcs = ["field2d.cu", "grid.cu", "strip.cu", "batch.cu", "gridstrip.cu"];
    let mut cmd = Command::new(env::var("NVCC").unwrap_or_else(|_| "nvcc".into()));
    cmd.args([
        "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
        // f32 arithmetic must round exactly like rustc's: no FMA contraction, IEEE div/sqrt
        "-fmad=false", "-prec-div=true", "-prec-sqrt=true",
        "-Xcompiler", "-fPIC", "--expt-extended-lambda", "-shared", "-o",
    ]);
    cmd.arg(&lib);
    for s in srcs {
        cmd.arg(csrc.join(s));
        println!("cargo:rerun-if-changed={}", csrc.join(s).display());
    }
    cmd.args(["-lcudart_static", "-lpthread", "-ldl", "-lrt"]);
    let status = cmd.status().expect("nvcc not found");
    assert!(status.success(), "nvcc failed");
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=krabgpu");
    println!("cargo:rerun-if-changed={}", root.join("include/krabgpu.h").display());
}
