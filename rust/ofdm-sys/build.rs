// build.rs -- compiles the engine's .cu sources with nvcc for sm_100a and links them into the `ofdm-sys` crate.
// No `cc` crate, no Triton, no multi-backend dispatch: one nvcc invocation, one static library.
// NOTE: the build image of this repository has no Rust toolchain, so this file is NOT compiled or tested there;
// the identical C ABI is exercised through C++ (ofdm_b200/host) and Python ctypes (ofdm_b200/engine.py).
use std::{env, path::PathBuf, process::Command};

fn main() {
    let root = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../..");
    let csrc = root.join("ofdm_b200/csrc");
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "/usr/local/cuda/bin/nvcc".into());
    let obj = out.join("ofdm_engine.o");
    let status = Command::new(&nvcc)
        .args(["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "-c"])
        .arg(csrc.join("ofdm_engine.cu"))
        .arg("-o")
        .arg(&obj)
        .status()
        .expect("nvcc not found: the engine has no CPU fallback");
    assert!(status.success(), "nvcc failed");
    let lib = out.join("libofdm_b200.a");
    assert!(Command::new("ar").args(["crs"]).arg(&lib).arg(&obj).status().unwrap().success());
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=static=ofdm_b200");
    println!("cargo:rustc-link-search=native=/usr/local/cuda/lib64");
    println!("cargo:rustc-link-lib=dylib=cudart");
    println!("cargo:rustc-link-lib=dylib=stdc++");
    for f in ["ofdm_engine.cu", "common.cuh", "rx_kernels.cuh", "tx_kernels.cuh", "tables.h"] {
        println!("cargo:rerun-if-changed={}", csrc.join(f).display());
    }
    println!("cargo:rerun-if-changed={}", root.join("include/ofdm_engine.h").display());
}
