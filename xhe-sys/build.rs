// build.rs -- link (default) or build (`--features build-cuda`) libxhe_cuda for sm_100a.  No CPU fallback is compiled.
use std::{env, path::PathBuf};

fn main() {
    let root = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("..");
    println!("cargo:rerun-if-changed={}", root.join("include/xhe.h").display());
    if env::var("CARGO_FEATURE_BUILD_CUDA").is_ok() {
        let csrc = root.join("xelis_he_b200/csrc");
        let host = root.join("xelis_he_b200/host");
        let mut b = cc::Build::new();
        b.cuda(true)
            .cudart("shared")
            .flag("-gencode").flag("arch=compute_100a,code=sm_100a")
            .flag("-O3").flag("-lineinfo").flag("-std=c++17")
            .include(root.join("include"));
        for f in ["capi.cu", "fiat_shamir.cu", "kernels_point.cu", "ledger.cu", "msm.cu", "verify.cu"] {
            b.file(csrc.join(f));
            println!("cargo:rerun-if-changed={}", csrc.join(f).display());
        }
        b.file(host.join("verifier.cpp"));     // the C++ mirror of the host layer; a Rust host does not need its symbols
        b.compile("xhe_cuda");
        println!("cargo:rustc-link-lib=dylib=cudart");
        println!("cargo:rustc-link-lib=dylib=stdc++");
    } else {
        let dir = env::var("XHE_LIB_DIR").unwrap_or_else(|_| root.join("xelis_he_b200").display().to_string());
        println!("cargo:rustc-link-search=native={dir}");
        println!("cargo:rustc-link-lib=dylib=xhe_cuda");
        println!("cargo:rerun-if-env-changed=XHE_LIB_DIR");
    }
}
