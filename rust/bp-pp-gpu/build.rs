//! Builds libbppp.so from the CUDA sources with nvcc for sm_100a (B200) and links it.
//!
//!   BPPP_PREBUILT=/path/to/dir   skip nvcc and link the libbppp.so found there (e.g. bp_pp_b200/ after `make`)
//!   NVCC=/usr/local/cuda/bin/nvcc, BPPP_CSRC=<repo>/bp_pp_b200/csrc   override the tool / source locations
use std::env;
use std::path::PathBuf;
use std::process::Command;

fn main() {
    let manifest = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap());
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    println!("cargo:rerun-if-env-changed=BPPP_PREBUILT");
    if let Ok(dir) = env::var("BPPP_PREBUILT") {
        println!("cargo:rustc-link-search=native={dir}");
        println!("cargo:rustc-link-lib=dylib=bppp");
        return;
    }
    let csrc = env::var("BPPP_CSRC").map(PathBuf::from).unwrap_or_else(|_| manifest.join("../../bp_pp_b200/csrc"));
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "nvcc".into());
    let units = ["engine_core", "engine_verify", "engine_prove", "engine_var", "engine_var_lat", "engine_bench", "engine_msm", "engine_wnla", "engine_circuit", "engine_multi", "engine_peer"];
    let mut objects = Vec::new();
    for u in units {
        let src = csrc.join(format!("{u}.cu"));
        println!("cargo:rerun-if-changed={}", src.display());
        let obj = out.join(format!("{u}.o"));
        let ok = Command::new(&nvcc)
            .args(["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "-diag-suppress", "550", "-c", "-o"])
            .arg(&obj).arg(&src).status().expect("nvcc not found: set NVCC or BPPP_PREBUILT").success();
        assert!(ok, "nvcc failed on {}", src.display());
        objects.push(obj);
    }
    for h in std::fs::read_dir(&csrc).unwrap().flatten() {
        if h.path().extension().map_or(false, |e| e == "cuh") { println!("cargo:rerun-if-changed={}", h.path().display()); }
    }
    let so = out.join("libbppp.so");
    let ok = Command::new(&nvcc).args(["-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o"]).arg(&so).args(&objects).status().unwrap().success();
    assert!(ok, "nvcc -shared failed");
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=bppp");
    println!("cargo:rustc-link-lib=dylib=cudart");
}
