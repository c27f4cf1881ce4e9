// Links the in-tree C-ABI library.  H2B200_LIB_DIR must point at halo2-prover_b200/csrc/.
fn main() {
    let dir = std::env::var("H2B200_LIB_DIR").expect("set H2B200_LIB_DIR to the directory holding libh2b200.so");
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=h2b200");
    println!("cargo:rerun-if-env-changed=H2B200_LIB_DIR");
}
