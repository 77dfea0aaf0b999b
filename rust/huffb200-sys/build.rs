// Links libhuffb200.so, built by `python -m huff_encoding_b200.build` (nvcc, sm_100a).
fn main() {
    let dir = std::env::var("HUFFB200_LIB_DIR").unwrap_or_else(|_| "../../huff_encoding_b200".to_string());
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=huffb200");
}
