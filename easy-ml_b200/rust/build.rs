// build.rs of easy-ml with the `cuda` feature (reference crate root).  Builds libeasyml_b200 for sm_100a with nvcc
// (through the library's own Makefile) and links it.  Without the feature this script does nothing, so the
// default build stays dependency-free and wasm32-clean (.github/workflows/tests.yml:39-42).
fn main() {
    println!("cargo:rerun-if-env-changed=EASYML_B200_DIR");
    if std::env::var_os("CARGO_FEATURE_CUDA").is_none() {
        return;
    }
    let root = std::env::var("EASYML_B200_DIR")
        .expect("the `cuda` feature needs EASYML_B200_DIR = a checkout of easy-ml_b200");
    let csrc = format!("{root}/easy-ml_b200/csrc");
    let status = std::process::Command::new("make")
        .args(["-C", &csrc, "-j8"])
        .status()
        .expect("could not run make (nvcc 12.9+ with sm_100a support is required)");
    assert!(status.success(), "building libeasyml_b200 (nvcc -gencode arch=compute_100a,code=sm_100a) failed");
    println!("cargo:rustc-link-search=native={root}/easy-ml_b200/lib");
    println!("cargo:rustc-link-lib=dylib=easyml_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{root}/easy-ml_b200/lib");
    println!("cargo:rerun-if-changed={csrc}");
    println!("cargo:rerun-if-changed={root}/include/easyml_b200.h");
}
