// Link against the in-tree shared library: set TRIPLE_ACCEL_B200_LIB_DIR to <repo>/triple_accel_b200
// (built by `make -C triple_accel_b200/csrc` or `python -c "import __graft_entry__ as g; g.build()"`).
fn main() {
    if let Ok(dir) = std::env::var("TRIPLE_ACCEL_B200_LIB_DIR") {
        println!("cargo:rustc-link-search=native={}", dir);
        println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir);
    }
    println!("cargo:rustc-link-lib=dylib=triple_accel_b200");
    println!("cargo:rerun-if-env-changed=TRIPLE_ACCEL_B200_LIB_DIR");
}
