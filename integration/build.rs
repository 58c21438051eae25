// build.rs for shimmer with the `gpu` feature (new file at the crate root; add `build = "build.rs"` and
// `[features] gpu = []` to Cargo.toml).  libshimmer_gpu.so is built from this repository by
// `python -c "import __graft_entry__ as g; g.build()"` (nvcc, sm_100a) and found through SHIMMER_GPU_LIB_DIR.
fn main() {
    println!("cargo:rerun-if-env-changed=SHIMMER_GPU_LIB_DIR");
    if std::env::var("CARGO_FEATURE_GPU").is_ok() {
        let dir = std::env::var("SHIMMER_GPU_LIB_DIR")
            .expect("set SHIMMER_GPU_LIB_DIR to the directory that holds libshimmer_gpu.so");
        println!("cargo:rustc-link-search=native={}", dir);
        println!("cargo:rustc-link-lib=dylib=shimmer_gpu");
        println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir);
    }
}
