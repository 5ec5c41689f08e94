// build.rs — builds libdryv_recon.so with nvcc and links it into the dryv binary.
//
// UNVERIFIED TEXT: this image has no cargo/rustc, so this file has never been compiled. It is the build
// script a dryv maintainer would drop next to Cargo.toml (plus `build = "build.rs"` in [package]);
// the nvcc command line is exactly the one dryv_b200/recon.py::build() runs and that IS exercised.
use std::env;
use std::path::PathBuf;
use std::process::Command;

fn main() {
  // location of this repository's csrc/ (vendored or a git submodule inside the dryv tree)
  let csrc = PathBuf::from(env::var("DRYV_RECON_CSRC").unwrap_or_else(|_| "dryv_recon/dryv_b200/csrc".into()));
  let out = PathBuf::from(env::var("OUT_DIR").unwrap());
  let lib = out.join("libdryv_recon.so");
  let status = Command::new(env::var("NVCC").unwrap_or_else(|_| "nvcc".into()))
    .args(["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17"])
    .args(["-shared", "-Xcompiler", "-fPIC"])
    .arg(csrc.join("recon.cu"))
    .arg(csrc.join("recon_tables.cpp"))
    .arg(csrc.join("levels_pack.cpp"))
    .arg(csrc.join("cabac_host.cpp"))
    .arg("-o")
    .arg(&lib)
    .status()
    .expect("nvcc not found: the reconstruction path has no CPU fallback");
  assert!(status.success(), "nvcc failed");
  println!("cargo:rustc-link-search=native={}", out.display());
  println!("cargo:rustc-link-lib=dylib=dryv_recon");
  println!("cargo:rustc-link-lib=dylib=cudart");
  println!("cargo:rerun-if-changed={}", csrc.join("recon.cu").display());
  println!("cargo:rerun-if-changed={}", csrc.join("recon_kernels.cuh").display());
  println!("cargo:rerun-if-changed={}", csrc.join("recon_tables.cpp").display());
  println!("cargo:rerun-if-changed={}", csrc.join("levels_pack.cpp").display());
}
