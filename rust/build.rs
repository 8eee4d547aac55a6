// build.rs of the reference crate with the `gpu` feature: builds libjjschnorr_b200.so with nvcc for sm_100a and links it.
// Place this repository under `cuda/` of the crate (or point JJS_B200_SRC at it).
// NOT compiled in this repository's environment (no cargo / rustc in the image).
use std::{env, path::PathBuf, process::Command};

fn main() {
    if env::var_os("CARGO_FEATURE_GPU").is_none() {
        return;
    }
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let root = PathBuf::from(env::var("JJS_B200_SRC").unwrap_or_else(|_| "cuda".into()));
    let src = root.join("jubjub_schnorr_b200/csrc/kernels.cu");
    let lib = out.join("libjjschnorr_b200.so");
    let status = Command::new(env::var("NVCC").unwrap_or_else(|_| "nvcc".into()))
        .args(["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-shared",
               "-Xcompiler", "-fPIC,-fvisibility=hidden,-pthread", "-o"])
        .arg(&lib)
        .arg(&src)
        .status()
        .expect("nvcc not found");
    assert!(status.success(), "nvcc failed");
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=jjschnorr_b200");
    println!("cargo:rerun-if-changed={}", root.display());
}
