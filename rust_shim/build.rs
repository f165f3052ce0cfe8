// Links libbhray.so.  BHRAY_LIB_DIR points at the directory holding it (bhusie_b200/lib after `python -m bhusie_b200.build`).
fn main() {
    if let Ok(dir) = std::env::var("BHRAY_LIB_DIR") {
        println!("cargo:rustc-link-search=native={dir}");
        println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    }
    println!("cargo:rustc-link-lib=dylib=bhray");
    println!("cargo:rerun-if-env-changed=BHRAY_LIB_DIR");
}
