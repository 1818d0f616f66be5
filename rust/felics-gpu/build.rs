// build.rs -- compiles the CUDA sources of the engine for sm_100a and links the result.
//
// Mirrors __graft_entry__.py (NVCC_FLAGS, SOURCES): one `nvcc -c` per source, one `nvcc -shared` link with the version
// script that exports only `felics_*`.  NOT RUN IN THIS REPOSITORY (no Rust toolchain in the build image).
use std::env;
use std::path::PathBuf;
use std::process::Command;

const SOURCES: &[&str] = &["api.cu", "encode.cu", "encode16.cu", "decode.cu", "stream.cu", "synth.cu"];

fn main() {
    let root = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../..");
    let csrc = root.join("felics_b200/csrc");
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "/usr/local/cuda/bin/nvcc".into());
    let mut objects = Vec::new();
    for src in SOURCES {
        let obj = out.join(format!("{src}.o"));
        let status = Command::new(&nvcc)
            .args(["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17"])
            .args(["-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr"])
            .arg("-c").arg(csrc.join(src)).arg("-o").arg(&obj)
            .status().expect("nvcc not found (set NVCC)");
        assert!(status.success(), "nvcc failed on {src}");
        objects.push(obj);
        println!("cargo:rerun-if-changed={}", csrc.join(src).display());
    }
    let lib = out.join("libfelics_b200.so");
    let status = Command::new(&nvcc)
        .arg("-shared").arg("-o").arg(&lib).args(&objects)
        .args(["-gencode", "arch=compute_100a,code=sm_100a"])
        .arg("-Xlinker").arg(format!("--version-script={}", csrc.join("exports.map").display()))
        .status().expect("nvcc not found (set NVCC)");
    assert!(status.success(), "link failed");
    for header in ["felics_b200.h"] {
        println!("cargo:rerun-if-changed={}", root.join("include").join(header).display());
    }
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=felics_b200");
    println!("cargo:rustc-link-lib=dylib=cudart");
}
