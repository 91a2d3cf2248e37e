fn main() {
    let dir = std::env::var("HGI_B200_LIB_DIR").expect("set HGI_B200_LIB_DIR to the directory of libhgi_b200.so");
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=hgi_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir);
}
