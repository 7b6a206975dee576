// Compiles the hand-written CUDA of scan_rs_b200/csrc for sm_100a and links it statically.  No multi-backend dispatch and no
// CPU fallback: without nvcc the build fails, without a B200 `Context::new` returns the library's SB_ERR_CUDA.
// Layout expected next to this file: csrc/ (a copy or symlink of scan_rs_b200/csrc) and include/ (include/scanb200.h).
use std::{env, fs, path::PathBuf, process::Command};

fn main() {
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let src = PathBuf::from(env::var("SCANB200_CSRC").unwrap_or_else(|_| "csrc".into()));
    let cuda = env::var("CUDA_HOME").unwrap_or_else(|_| "/usr/local/cuda".into());
    // every translation unit of the library: the list is the directory, so it cannot fall behind the Makefile
    let mut files: Vec<PathBuf> = fs::read_dir(&src)
        .expect("csrc/ not found: copy or symlink scan_rs_b200/csrc next to build.rs (or set SCANB200_CSRC)")
        .filter_map(|e| e.ok().map(|e| e.path()))
        .filter(|p| p.extension().map_or(false, |x| x == "cu"))
        .collect();
    files.sort();
    let mut objs = vec![];
    for f in &files {
        let o = out.join(f.file_stem().unwrap()).with_extension("o");
        let mut c = Command::new(format!("{cuda}/bin/nvcc"));
        c.args(["-O3", "-std=c++17", "-lineinfo", "-gencode", "arch=compute_100a,code=sm_100a",
                "-Xcompiler", "-fPIC,-fvisibility=hidden", "--expt-relaxed-constexpr", "-c"]);
        if f.file_name().unwrap() == "synth.cu" {
            c.arg("-fmad=false"); // bit-exact twin of the CPU workload generator
        }
        assert!(c.arg(f).arg("-o").arg(&o).status().expect("nvcc").success(), "nvcc failed on {}", f.display());
        objs.push(o);
        println!("cargo:rerun-if-changed={}", f.display());
    }
    for h in ["common.cuh", "map.cuh", "tc05.cuh", "gather_units.h", "nccl_shim.h", "synth_nb.h", "log_table.h"] {
        println!("cargo:rerun-if-changed={}", src.join(h).display());
    }
    let lib = out.join("libscanb200.a");
    let _ = fs::remove_file(&lib);
    assert!(Command::new("ar").arg("crs").arg(&lib).args(&objs).status().expect("ar").success());
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-search=native={cuda}/lib64");
    println!("cargo:rustc-link-lib=static=scanb200");
    // NCCL is not linked: the library dlopens libnccl.so.2 at the first communicator call (csrc/nccl_shim.h)
    for l in ["cudart", "cublas", "cusolver", "stdc++", "dl", "pthread"] {
        println!("cargo:rustc-link-lib={l}");
    }
}
