// Links libb381_cuda.so.  B381_LIB_DIR = directory holding it (default: ../midnight_bls12_381_cuda_b200/lib, where
// `python -c "import __graft_entry__ as g; g.build()"` leaves it).
use std::{env, path::PathBuf};

fn main() {
    let dir = env::var("B381_LIB_DIR").map(PathBuf::from).unwrap_or_else(|_| {
        PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../midnight_bls12_381_cuda_b200/lib")
    });
    println!("cargo:rustc-link-search=native={}", dir.display());
    println!("cargo:rustc-link-lib=dylib=b381_cuda");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir.display());
    println!("cargo:rerun-if-env-changed=B381_LIB_DIR");
}
