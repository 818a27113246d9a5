// Compiles the CUDA sources of the C-ABI library for sm_100a and links them into the crate.
// The translation units come from halo2-svd041_b200/csrc/SOURCES.txt -- the same list halo2-svd041_b200/build.py reads,
// so the two builds cannot drift apart (tests/test_abi.py checks that the list names every .cu in csrc/).
// (Unverified here: no Rust toolchain in the build container -- see INTEGRATION.md.)
use std::{env, fs, path::PathBuf, process::Command};

fn main() {
    let root = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../..");
    let csrc = root.join("halo2-svd041_b200/csrc");
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "/usr/local/cuda/bin/nvcc".into());
    let list = csrc.join("SOURCES.txt");
    println!("cargo:rerun-if-changed={}", list.display());
    let sources: Vec<String> = fs::read_to_string(&list)
        .expect("csrc/SOURCES.txt")
        .lines()
        .map(str::trim)
        .filter(|l| !l.is_empty() && !l.starts_with('#'))
        .map(String::from)
        .collect();
    // headers: any change rebuilds everything
    for e in fs::read_dir(&csrc).unwrap().flatten() {
        println!("cargo:rerun-if-changed={}", e.path().display());
    }
    println!("cargo:rerun-if-changed={}", root.join("include/h2svd_b200.h").display());
    let mut objs = Vec::new();
    for src in &sources {
        let obj = out.join(src.replace(".cu", ".o"));
        let st = Command::new(&nvcc)
            .args(["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
                   "-Xcompiler", "-fPIC", "-c"])
            .arg(csrc.join(src)).arg("-o").arg(&obj)
            .status().expect("nvcc not found");
        assert!(st.success(), "nvcc failed on {src}");
        objs.push(obj);
    }
    let lib = out.join("libh2svd_b200.a");
    let st = Command::new("ar").arg("crs").arg(&lib).args(&objs).status().unwrap();
    assert!(st.success());
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=static=h2svd_b200");
    println!("cargo:rustc-link-search=native=/usr/local/cuda/lib64");
    println!("cargo:rustc-link-lib=dylib=cudart");
    println!("cargo:rustc-link-lib=dylib=stdc++");
}
