//! B200-native BLS12-381 proving backend -- Rust host layer (mirror of the reference's `core/`: msm.rs, ntt.rs,
//! vecops.rs, stream.rs, types.rs) over the C ABI of `libb381_cuda.so` (`include/b381.h`).
//!
//! What is deliberately NOT here: `dispatch.rs` and `config.rs` (the hybrid CPU/GPU thresholds and the CPU fallback
//! are removed: if the device call fails the error is returned) and `traits/` (the `MsmBackend` indirection existed to
//! switch between BLST and ICICLE).
//!
//! This crate could not be compiled in the environment it was written in (no Rust toolchain in the build image); the
//! same binding is exercised through ctypes by the Python package `midnight_bls12_381_cuda_b200`, whose tests run
//! against the CPU oracle on the GPU.  See INTEGRATION.md.
pub mod ffi;
pub mod msm;
pub mod ntt;
pub mod stream;
pub mod types;
pub mod vecops;

pub use msm::{BatchMsmHandle, GpuMsmContext, MsmError, MsmHandle, PrecomputedBases};
pub use ntt::{GpuNttContext, NttError, NttHandle};
pub use stream::{ensure_backend_loaded, is_gpu_available, DeviceVec, GpuError, ManagedStream, PinnedVec};
pub use types::{G1Result, G2Result, PodG1Affine, PodG2Affine, PodScalar, TypeConverter};
pub use vecops::VecOpsError;
