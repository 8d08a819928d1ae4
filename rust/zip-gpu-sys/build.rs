// UNCOMPILED (no Rust toolchain in this image) -- see ../README.md
//
// libzipgpu.so is built by `python -m zinc_b200.build` (nvcc, sm_100a) into zinc_b200/; point ZIPGPU_LIB_DIR at it.
fn main() {
    let dir = std::env::var("ZIPGPU_LIB_DIR").unwrap_or_else(|_| "../../zinc_b200".to_string());
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=zipgpu");
    println!("cargo:rerun-if-env-changed=ZIPGPU_LIB_DIR");
}
