// Links the C-ABI library built by `make lib` (quantr_b200/libqsv.so).
fn main() {
    let dir = std::env::var("QSV_LIB_DIR").unwrap_or_else(|_| "../../quantr_b200".to_string());
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=qsv");
    println!("cargo:rerun-if-env-changed=QSV_LIB_DIR");
}
