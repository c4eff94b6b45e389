// Tells cargo where libsgx.so lives: SGX_LIB_DIR, else <repo>/multi-spectrogram-viewer_b200 next to this crate.
use std::env;
use std::path::PathBuf;

fn main() {
    let dir = env::var("SGX_LIB_DIR").map(PathBuf::from).unwrap_or_else(|_| {
        PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../../../multi-spectrogram-viewer_b200")
    });
    println!("cargo:rustc-link-search=native={}", dir.display());
    println!("cargo:rustc-link-lib=dylib=sgx");
    println!("cargo:rerun-if-env-changed=SGX_LIB_DIR");
}
