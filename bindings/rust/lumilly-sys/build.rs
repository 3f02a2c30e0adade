// UNTESTED SOURCE (no Rust toolchain offline).  Compiles the hand-written sm_100a kernels and the C ABI with
// nvcc — the same command lumillyrender_b200/build.py runs — and links the result.
use std::env;
use std::path::PathBuf;
use std::process::Command;

fn main() {
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let root = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../../..");
    let csrc = root.join("lumillyrender_b200/csrc");
    let lib = out.join("liblumilly_b200.so");
    let sources = ["kernels.cu", "api.cpp", "bvh_build.cpp", "toml_obj.cpp", "host_scene.cpp", "image_io.cpp"];
    let mut cmd = Command::new(env::var("NVCC").unwrap_or_else(|_| "nvcc".into()));
    cmd.args(&["-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-O3", "-lineinfo",
               "-fmad=false", "-prec-div=true", "-prec-sqrt=true",
               "-Xcompiler", "-fPIC,-ffp-contract=off", "-shared", "-o"]).arg(&lib);
    for s in sources.iter() {
        cmd.arg(csrc.join(s));
        println!("cargo:rerun-if-changed={}", csrc.join(s).display());
    }
    cmd.arg("-lz");
    let status = cmd.status().expect("nvcc not found: lumilly-sys has no CPU fallback");
    assert!(status.success(), "nvcc failed");
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=lumilly_b200");
}
