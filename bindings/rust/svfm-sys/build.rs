// Links libsvfm.so, built by `make -C sview_fmindex_b200/csrc` (nvcc, sm_100a).
fn main() {
    let dir = std::env::var("SVFM_LIB_DIR").unwrap_or_else(|_| "../../../sview_fmindex_b200".to_string());
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=svfm");
    println!("cargo:rerun-if-env-changed=SVFM_LIB_DIR");
}
