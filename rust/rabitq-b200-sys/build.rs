// build.rs -- compiles the CUDA sources of librbq with nvcc for sm_100a and links the result.
// (Equivalent of rabitq_rs_b200/build.py; untested here: no Rust toolchain in the build image.)
use std::{env, path::PathBuf, process::Command};

fn main() {
    let root = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../..");
    let csrc = root.join("rabitq_rs_b200/csrc");
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "/usr/local/cuda/bin/nvcc".into());
    let sources = ["format.cc", "query_prep.cu", "coarse.cu", "coarse_tc.cu", "scan.cu", "scan_tail.cu", "tail_tc.cu", "resolve.cu",
                   "fetch.cu", "build.cu", "api.cu"];  // keep in sync with rabitq_rs_b200/build.py::SOURCES
    let mut objs = Vec::new();
    for s in sources {
        let obj = out.join(format!("{s}.o"));
        let ok = Command::new(&nvcc)
            .args(["-x", "cu", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
                   "-fmad=false", "-Xcompiler", "-fPIC", "-c"])
            .arg(csrc.join(s)).arg("-o").arg(&obj)
            .status().expect("nvcc not found").success();
        assert!(ok, "nvcc failed on {s}");
        objs.push(obj);
        println!("cargo:rerun-if-changed={}", csrc.join(s).display());
    }
    let lib = out.join("librbq.so");
    let ok = Command::new(&nvcc)
        .args(["-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-cudart", "static", "-o"]).arg(&lib).args(&objs)
        .status().unwrap().success();
    assert!(ok, "link failed");
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=rbq");
}
