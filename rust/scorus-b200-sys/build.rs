// Links libscorus_b200.so.  With SCORUS_B200_LIB_DIR set, the prebuilt library (make -C
// scorus_b200/csrc) is used; otherwise the .cu sources are compiled with cc's CUDA mode for
// sm_100a -- no Triton, no multi-backend dispatch, no CPU fallback.
use std::{env, path::PathBuf};

fn main() {
    let root = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../..");
    if let Ok(dir) = env::var("SCORUS_B200_LIB_DIR") {
        println!("cargo:rustc-link-search=native={dir}");
        println!("cargo:rustc-link-lib=dylib=scorus_b200");
        return;
    }
    let csrc = root.join("scorus_b200/csrc");
    let plugins = [
        "LpGaussian", "LpBoundedGaussian", "LpRosenbrock", "LpRastrigin", "LpBimodal2d", "LpStep2d",
        "LpCorrGaussian", "LpMixture",
    ];
    let base = |b: &mut cc::Build| {
        b.cuda(true)
            .flag("-gencode").flag("arch=compute_100a,code=sm_100a")
            .flag("-std=c++17").flag("-O3").flag("-lineinfo").flag("-fmad=false")
            .flag("-Xcompiler").flag("-ffp-contract=off")
            .include(&csrc).include(root.join("include"));
    };
    let mut main = cc::Build::new();
    base(&mut main);
    main.file(csrc.join("scorus_b200.cu")).file(csrc.join("launch_dispatch.cu")).compile("scorus_b200_core");
    for p in plugins {
        let mut b = cc::Build::new();
        base(&mut b);
        b.define("SCORUS_PLUGIN", p).file(csrc.join("launch_plugin.cu")).compile(&format!("scorus_b200_{p}"));
    }
    println!("cargo:rustc-link-lib=dylib=cudart");
    println!("cargo:rerun-if-changed={}", csrc.display());
}
