// Builds libsvdlas2_b200.so with the repository's Makefile (nvcc, sm_100a) and links it.
use std::{env, path::PathBuf, process::Command};

fn main() {
    let root = PathBuf::from(
        env::var("SVDLAS2_B200_DIR").expect("set SVDLAS2_B200_DIR to the svdlas2-b200 repository root"),
    );
    let status = Command::new("make")
        .arg("-C")
        .arg(root.join("svdlibrs_b200/csrc"))
        .arg("-j8")
        .status()
        .expect("failed to run make");
    assert!(status.success(), "nvcc build of libsvdlas2_b200.so failed");
    let libdir = root.join("svdlibrs_b200");
    println!("cargo:rustc-link-search=native={}", libdir.display());
    println!("cargo:rustc-link-lib=dylib=svdlas2_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", libdir.display());
    println!("cargo:rerun-if-changed={}", root.join("include/svdlas2.h").display());
    println!("cargo:rerun-if-env-changed=SVDLAS2_B200_DIR");
}
