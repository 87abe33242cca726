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
    for api in ["api_core.cu", "api_shard.cu", "api_stats.cu", "api_moves.cu", "api_host.cu", "launch_dispatch.cu"] {
        main.file(csrc.join(api));
    }
    main.compile("scorus_b200_core");
    // one translation unit per (kernel family, log-density plugin), as scorus_b200/csrc/Makefile does
    for unit in ["launch", "twalk", "hmc"] {
        for p in plugins {
            let mut b = cc::Build::new();
            base(&mut b);
            b.define("SCORUS_PLUGIN", p).file(csrc.join(format!("{unit}_plugin.cu"))).compile(&format!("scorus_b200_{unit}_{p}"));
        }
    }
    println!("cargo:rustc-link-lib=dylib=cudart");
    println!("cargo:rerun-if-changed={}", csrc.display());
}
