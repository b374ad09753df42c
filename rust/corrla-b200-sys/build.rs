// Link against the prebuilt shared library: CORRLA_B200_LIB_DIR=/path/to/corrla_rs_b200/lib
fn main() {
    if let Ok(dir) = std::env::var("CORRLA_B200_LIB_DIR") {
        println!("cargo:rustc-link-search=native={dir}");
    }
    println!("cargo:rustc-link-lib=dylib=corrla_b200");
}
