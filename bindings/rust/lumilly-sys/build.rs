// UNTESTED SOURCE (no Rust toolchain offline).  Builds the hand-written sm_100a kernels and the C ABI by running
// lumillyrender_b200/build.py — nvcc on csrc/{kernels.cu,api.cpp,bvh_build.cpp,toml_obj.cpp,host_scene.cpp,
// image_io.cpp} plus the eight (integrator x BVH x GGX) instantiation units of the render kernel
// (csrc/persistent_inst.cu), all with -gencode arch=compute_100a,code=sm_100a -fmad=false -prec-div=true
// -prec-sqrt=true — and links the resulting liblumilly_b200.so.  There is no CPU fallback: without nvcc this fails.
use std::env;
use std::path::PathBuf;
use std::process::Command;

fn main() {
    let root = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../../..");
    let pkg = root.join("lumillyrender_b200");
    let python = env::var("PYTHON").unwrap_or_else(|_| "python3".into());
    let status = Command::new(python)
        .arg(pkg.join("build.py"))
        .status()
        .expect("python3 not found: lumilly-sys builds through lumillyrender_b200/build.py");
    assert!(status.success(), "lumillyrender_b200/build.py failed (nvcc missing? there is no CPU fallback)");
    for f in ["build.py", "csrc/kernels.cu", "csrc/api.cpp", "csrc/persistent_inst.cu", "csrc/persistent.cuh", "csrc/pool.cuh", "csrc/path_vertex.inc",
              "csrc/device_path.cuh", "csrc/device_scene.h", "csrc/bvh_build.cpp", "csrc/toml_obj.cpp", "csrc/host_scene.cpp",
              "csrc/image_io.cpp"].iter() {
        println!("cargo:rerun-if-changed={}", pkg.join(f).display());
    }
    println!("cargo:rustc-link-search=native={}", pkg.display());
    println!("cargo:rustc-link-lib=dylib=lumilly_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", pkg.display());
}
