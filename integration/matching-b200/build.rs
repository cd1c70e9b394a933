// Links libslideo_b200.so (built by `make -C slideo_b200/csrc` of the slideo_b200 repository).  Point SLIDEO_B200_LIB_DIR at the
// directory that holds the shared object; at run time the loader must find it too (LD_LIBRARY_PATH or an rpath, added below).
fn main() {
    println!("cargo:rerun-if-env-changed=SLIDEO_B200_LIB_DIR");
    let dir = std::env::var("SLIDEO_B200_LIB_DIR").expect("set SLIDEO_B200_LIB_DIR to the directory of libslideo_b200.so");
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=slideo_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir);
}
