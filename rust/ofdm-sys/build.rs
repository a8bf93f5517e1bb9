// build.rs -- compiles the engine's .cu translation units with nvcc for sm_100a and links them into the `ofdm-sys` crate.
// No `cc` crate, no Triton, no multi-backend dispatch: one nvcc invocation per unit (the same list as ofdm_b200/_build.py),
// one static library. NOTE: the build image of this repository has no Rust toolchain, so this file is NOT compiled or tested
// there; the identical C ABI is exercised through C++ (ofdm_b200/host) and Python ctypes (ofdm_b200/engine.py).
use std::{env, fs, path::PathBuf, process::Command};

const UNITS: [&str; 16] = ["rx64_m0", "rx64_m1", "rx64_m2", "rx64", "tx64", "wide_rx_m0", "wide_rx_m1", "wide_rx_m2", "wide_rx", "wide_tx", "wide_txr", "tx64r", "tx64w", "sync", "rs", "ofdm_engine"];

fn main() {
    let root = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../..");
    let csrc = root.join("ofdm_b200/csrc");
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "/usr/local/cuda/bin/nvcc".into());
    let mut objs = Vec::new();
    let children: Vec<_> = UNITS
        .iter()
        .map(|u| {
            let obj = out.join(format!("{u}.o"));
            objs.push(obj.clone());
            Command::new(&nvcc)
                .args(["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "-c"])
                .arg(csrc.join(format!("{u}.cu")))
                .arg("-o")
                .arg(&obj)
                .spawn()
                .expect("nvcc not found: the engine has no CPU fallback")
        })
        .collect();
    for mut c in children {
        assert!(c.wait().unwrap().success(), "nvcc failed");
    }
    let lib = out.join("libofdm_b200.a");
    assert!(Command::new("ar").args(["crs"]).arg(&lib).args(&objs).status().unwrap().success());
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=static=ofdm_b200");
    println!("cargo:rustc-link-search=native=/usr/local/cuda/lib64");
    println!("cargo:rustc-link-lib=dylib=cudart");
    println!("cargo:rustc-link-lib=dylib=stdc++");
    println!("cargo:rustc-link-lib=dylib=dl");
    // every source the units include: all of csrc/ (kernels.h pulls in every .cuh) and the public header
    for e in fs::read_dir(&csrc).unwrap() {
        println!("cargo:rerun-if-changed={}", e.unwrap().path().display());
    }
    println!("cargo:rerun-if-changed={}", root.join("include/ofdm_engine.h").display());
}
