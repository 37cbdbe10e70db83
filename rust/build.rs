fn main() {
    // point BSHARK_LIB_DIR at the directory holding libbshark_cuda.so (baby_shark_b200/ in this repository)
    if let Ok(dir) = std::env::var("BSHARK_LIB_DIR") {
        println!("cargo:rustc-link-search=native={dir}");
        println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    }
    println!("cargo:rustc-link-lib=dylib=bshark_cuda");
}
