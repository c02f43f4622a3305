// Locates libaqua_cuda.so (built by aqua-engine_b200/build.py).
// AQUA_CUDA_LIB_DIR overrides the default in-tree location.
use std::env;
use std::path::PathBuf;

fn main() {
    let dir = env::var("AQUA_CUDA_LIB_DIR").map(PathBuf::from).unwrap_or_else(|_| {
        PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../../aqua-engine_b200")
    });
    println!("cargo:rustc-link-search=native={}", dir.display());
    println!("cargo:rustc-link-lib=dylib=aqua_cuda");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir.display());
    println!("cargo:rerun-if-env-changed=AQUA_CUDA_LIB_DIR");
}
