// Builds libtray_cuda.so with nvcc for sm_100a and links it.  Same command line as tray_racing_b200/csrc/Makefile.
use std::{env, path::PathBuf, process::Command};

fn main() {
    let root = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("..");
    let src = root.join("tray_racing_b200/csrc/tray_cuda.cu");
    let builder = root.join("tray_racing_b200/csrc/build_gpu.cu");
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let lib = out.join("libtray_cuda.so");
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "nvcc".into());
    let status = Command::new(nvcc)
        .args(["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-fmad=false"])
        .args(["-Xcompiler", "-fPIC", "-shared"])
        .arg(format!("-I{}", root.join("include").display()))
        .arg("-o").arg(&lib).arg(&src).arg(&builder)
        .status()
        .expect("failed to run nvcc");
    assert!(status.success(), "nvcc failed");
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=tray_cuda");
    println!("cargo:rerun-if-changed={}", src.display());
    println!("cargo:rerun-if-changed={}", builder.display());
    println!("cargo:rerun-if-changed={}", root.join("tray_racing_b200/csrc/traverse.cuh").display());
    println!("cargo:rerun-if-changed={}", root.join("include/tray_cuda.h").display());
}
