// build.rs of easy-ml with the `cuda` feature (reference crate root).  Builds libeasyml_b200.a -- a STATIC library of
// CUDA objects compiled by nvcc for sm_100a (`-gencode arch=compute_100a,code=sm_100a`, through the library's own
// Makefile) -- and links it, together with the CUDA runtime, into the crate.  Without the feature this script does
// nothing, so the default build stays dependency-free and wasm32-clean (.github/workflows/tests.yml:39-42).
fn main() {
    println!("cargo:rerun-if-env-changed=EASYML_B200_DIR");
    println!("cargo:rerun-if-env-changed=CUDA_HOME");
    if std::env::var_os("CARGO_FEATURE_CUDA").is_none() {
        return;
    }
    let root = std::env::var("EASYML_B200_DIR")
        .expect("the `cuda` feature needs EASYML_B200_DIR = a checkout of easy-ml_b200");
    let cuda = std::env::var("CUDA_HOME").unwrap_or_else(|_| "/usr/local/cuda".to_string());
    let csrc = format!("{root}/easy-ml_b200/csrc");
    let jobs = std::env::var("NUM_JOBS").unwrap_or_else(|_| "8".to_string());
    let status = std::process::Command::new("make")
        .args(["-C", &csrc, &format!("-j{jobs}"), "../lib/libeasyml_b200.a"])
        .env("NVCC", format!("{cuda}/bin/nvcc"))
        .status()
        .expect("could not run make (nvcc 12.9+ with sm_100a support is required)");
    assert!(status.success(), "building libeasyml_b200.a (nvcc -gencode arch=compute_100a,code=sm_100a) failed");
    // the archive holds host objects with embedded sm_100a cubins (no relocatable device code, so no device link step);
    // they need the CUDA runtime and the C++ runtime the host side of the .cu files was compiled against
    println!("cargo:rustc-link-search=native={root}/easy-ml_b200/lib");
    println!("cargo:rustc-link-lib=static=easyml_b200");
    println!("cargo:rustc-link-search=native={cuda}/lib64");
    println!("cargo:rustc-link-lib=static=cudart_static");
    for lib in ["stdc++", "dl", "pthread", "rt"] {
        println!("cargo:rustc-link-lib=dylib={lib}");
    }
    println!("cargo:rerun-if-changed={csrc}");
    println!("cargo:rerun-if-changed={root}/include/easyml_b200.h");
}
