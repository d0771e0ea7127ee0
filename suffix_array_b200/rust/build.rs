//! build.rs for the crate once `cdivsufsort` is replaced by libsab200 (NOT RUN HERE: no cargo).
//! Compiles the single CUDA translation unit for sm_100a only -- no other arch, no runtime backend
//! dispatch, no CPU fallback -- exactly as suffix_array_b200/csrc/Makefile does.
use std::env;
use std::path::PathBuf;
use std::process::Command;

fn main() {
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let csrc = PathBuf::from(env::var("SAB200_CSRC").unwrap_or_else(|_| "sab200/csrc".into()));
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "/usr/local/cuda/bin/nvcc".into());
    let lib = out.join("libsab200.so");
    let status = Command::new(nvcc)
        .args(&["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17"])
        .args(&["-Xcompiler", "-fPIC", "-shared", "-o"])
        .arg(&lib)
        .arg(csrc.join("sab_api.cu"))
        .status()
        .expect("nvcc not found");
    assert!(status.success(), "nvcc failed");
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=sab200");
    println!("cargo:rerun-if-changed={}", csrc.display());
}
