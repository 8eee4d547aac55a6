// Extracted from INTEGRATION.md by tools/extract_rust_shim.py -- edit the markdown, not this file.
// NOT compiled in this repository's environment (no cargo/rustc in the image).

// build.rs — builds the CUDA library with nvcc for sm_100a and links it.
use std::{env, path::PathBuf, process::Command};

fn main() {
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let src = PathBuf::from("cuda/jubjub_schnorr_b200/csrc/kernels.cu"); // this repo vendored under cuda/
    let lib = out.join("libjjschnorr_b200.so");
    let status = Command::new(env::var("NVCC").unwrap_or_else(|_| "nvcc".into()))
        .args(["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-shared",
               "-Xcompiler", "-fPIC,-fvisibility=hidden", "-o"])
        .arg(&lib)
        .arg(&src)
        .status()
        .expect("nvcc not found");
    assert!(status.success(), "nvcc failed");
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=jjschnorr_b200");
    println!("cargo:rerun-if-changed=cuda");
}
