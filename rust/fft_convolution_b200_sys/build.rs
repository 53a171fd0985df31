// Links the prebuilt C-ABI library (fft_convolution_b200/libfftconv_b200.so, built by nvcc for
// sm_100a).  bindgen is not used: the extern block in src/lib.rs is written by hand from
// include/fftconv_b200.h.
fn main() {
    let dir = std::env::var("FFTCONV_B200_LIB_DIR")
        .unwrap_or_else(|_| "../../fft_convolution_b200".to_string());
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=fftconv_b200");
    println!("cargo:rerun-if-env-changed=FFTCONV_B200_LIB_DIR");
}
