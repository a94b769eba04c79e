// Links liblasso_b200.so.  LASSO_B200_LIB_DIR = <lasso_b200 checkout>/lasso_b200 (where `make -C lasso_b200/csrc`
// or `python -c "import __graft_entry__ as g; g.build()"` puts the library).
fn main() {
    let dir = std::env::var("LASSO_B200_LIB_DIR").unwrap_or_else(|_| "../../../lasso_b200".to_string());
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=lasso_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir);
    println!("cargo:rerun-if-env-changed=LASSO_B200_LIB_DIR");
}
