//! Element-wise Fr vector operations -- core/vecops.rs:140-570 over `b381_vector_*` / `b381_scalar_*_vec` /
//! `b381_bit_reverse`.  `should_use_gpu_vecops` and the CPU branches are gone with the rest of the dispatcher.
use crate::{ffi, stream::{check, DeviceVec, GpuError}, types::*};

#[derive(Debug)]
pub enum VecOpsError {
    SizeMismatch { expected: usize, got: usize },
    InvalidSize(usize),
    ExecutionFailed(String),
    Gpu(GpuError),
}
impl std::fmt::Display for VecOpsError {
    fn fmt(&self, f: &mut std::fmt::Formatter<'_>) -> std::fmt::Result { write!(f, "{self:?}") }
}
impl std::error::Error for VecOpsError {}
impl From<GpuError> for VecOpsError {
    fn from(e: GpuError) -> Self { VecOpsError::Gpu(e) }
}

type BinOp = unsafe extern "C" fn(*const ffi::Fr, *const ffi::Fr, u64, *const ffi::VecOpsConfig, *mut ffi::Fr) -> i32;

fn binary<S: PodScalar>(op: BinOp, a: &[S], b: &[S], a_is_scalar: bool) -> Result<Vec<S>, VecOpsError> {
    if !a_is_scalar && a.len() != b.len() { return Err(VecOpsError::SizeMismatch { expected: a.len(), got: b.len() }); }
    if b.is_empty() { return Ok(vec![]); }
    let (ia, ib) = (TypeConverter::scalar_slice_as_icicle(a), TypeConverter::scalar_slice_as_icicle(b));
    let cfg = unsafe { ffi::b381_default_vecops_config() };
    let mut out = vec![[0u64; 4]; ib.len()];
    check(unsafe { op(ia.as_ptr(), ib.as_ptr(), ib.len() as u64, &cfg, out.as_mut_ptr()) })?;
    Ok(TypeConverter::icicle_slice_as_scalars::<S>(&out).to_vec())
}

pub fn vector_add<S: PodScalar>(a: &[S], b: &[S]) -> Result<Vec<S>, VecOpsError> { binary(ffi::b381_vector_add, a, b, false) }
pub fn vector_sub<S: PodScalar>(a: &[S], b: &[S]) -> Result<Vec<S>, VecOpsError> { binary(ffi::b381_vector_sub, a, b, false) }
pub fn vector_mul<S: PodScalar>(a: &[S], b: &[S]) -> Result<Vec<S>, VecOpsError> { binary(ffi::b381_vector_mul, a, b, false) }
/// core/vecops.rs:315-365: one scalar broadcast over the vector (no repeated copy of the scalar is built)
pub fn scalar_mul<S: PodScalar>(scalar: S, a: &[S]) -> Result<Vec<S>, VecOpsError> { binary(ffi::b381_scalar_mul_vec, &[scalar], a, true) }
pub fn scalar_add<S: PodScalar>(scalar: S, a: &[S]) -> Result<Vec<S>, VecOpsError> { binary(ffi::b381_scalar_add_vec, &[scalar], a, true) }

/// core/vecops.rs:392-450
pub fn bit_reverse<S: PodScalar>(input: &[S]) -> Result<Vec<S>, VecOpsError> {
    if input.is_empty() { return Ok(vec![]); }
    if !input.len().is_power_of_two() { return Err(VecOpsError::InvalidSize(input.len())); }
    let v = TypeConverter::scalar_slice_as_icicle(input);
    let cfg = unsafe { ffi::b381_default_vecops_config() };
    let mut out = vec![[0u64; 4]; v.len()];
    check(unsafe { ffi::b381_bit_reverse(v.as_ptr(), v.len() as u64, &cfg, out.as_mut_ptr()) })?;
    Ok(TypeConverter::icicle_slice_as_scalars::<S>(&out).to_vec())
}
/// core/vecops.rs:454-535: permutes on the device, in place
pub fn bit_reverse_inplace<S: PodScalar>(input: &mut [S]) -> Result<(), VecOpsError> {
    if input.is_empty() { return Ok(()); }
    if !input.len().is_power_of_two() { return Err(VecOpsError::InvalidSize(input.len())); }
    let v = TypeConverter::scalar_slice_as_icicle_mut(input);
    let mut d = DeviceVec::<ffi::Fr>::from_host(v)?;
    let mut cfg = unsafe { ffi::b381_default_vecops_config() };
    cfg.is_a_on_device = true;
    cfg.is_result_on_device = true;
    let p = d.as_mut_ptr();
    check(unsafe { ffi::b381_bit_reverse(p, v.len() as u64, &cfg, p) })?;
    d.copy_to_host(v)?;
    Ok(())
}
