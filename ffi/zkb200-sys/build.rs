// Links libzkb200.so.  Point ZKB200_LIB_DIR at the directory holding it
// (zk-research-implementations_b200/ after `make -C zk-research-implementations_b200/csrc`).
fn main() {
    let dir = std::env::var("ZKB200_LIB_DIR").unwrap_or_else(|_| "../../zk-research-implementations_b200".into());
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=zkb200");
    println!("cargo:rerun-if-env-changed=ZKB200_LIB_DIR");
}
