// Links libivpb.so.  IVPB_LIB_DIR points at the directory holding it (default: ../../ivp_b200/lib).
fn main() {
    let dir = std::env::var("IVPB_LIB_DIR").unwrap_or_else(|_| {
        let here = std::env::var("CARGO_MANIFEST_DIR").unwrap();
        format!("{here}/../../ivp_b200/lib")
    });
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=ivpb");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    println!("cargo:rerun-if-env-changed=IVPB_LIB_DIR");
}
