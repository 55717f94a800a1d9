// Point the linker at the directory holding liblzkp_b200.so (LZKP_B200_LIB_DIR, default ../../../libzkp_b200/_lib).
fn main() {
    let dir = std::env::var("LZKP_B200_LIB_DIR").unwrap_or_else(|_| {
        let manifest = std::env::var("CARGO_MANIFEST_DIR").unwrap();
        format!("{}/../../../libzkp_b200/_lib", manifest)
    });
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=lzkp_b200");
    println!("cargo:rerun-if-env-changed=LZKP_B200_LIB_DIR");
}
