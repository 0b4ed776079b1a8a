// Links libcrane_b200.so.  CRANE_B200_LIB_DIR = the directory holding it (crane_b200/ of this repository after `python -m
// crane_b200.build`); the library has no dependencies besides the CUDA runtime it was linked against.
fn main() {
    let dir = std::env::var("CRANE_B200_LIB_DIR").unwrap_or_else(|_| "../../crane_b200".to_string());
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=crane_b200");
    println!("cargo:rerun-if-env-changed=CRANE_B200_LIB_DIR");
}
