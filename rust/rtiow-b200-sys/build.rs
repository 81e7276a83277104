// UNCOMPILED (see ../README.md).  Links the in-tree build of the C ABI library:
//   RTIOW_B200_LIB_DIR=/path/to/repo/rtiow-rust_b200/_build cargo build
// (`_build_fast` for the tolerance build: same symbols, FMA contraction + approximate division.)
fn main() {
    let dir = std::env::var("RTIOW_B200_LIB_DIR").expect("set RTIOW_B200_LIB_DIR to the directory holding librtiow_b200.so");
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=rtiow_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir);
    println!("cargo:rerun-if-env-changed=RTIOW_B200_LIB_DIR");
}
