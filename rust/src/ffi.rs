//! Raw bindings of `include/b381.h`.  Struct layouts are the ICICLE v4 config structs the reference's backend receives
//! (bls12-381/include/icicle_types.cuh:102-201); `tests/test_abi.py` pins the offsets on the C side.
#![allow(non_camel_case_types, dead_code)]
use std::os::raw::{c_int, c_void};

pub type Fr = [u64; 4];
pub type G1Affine = [u64; 12];
pub type G1Projective = [u64; 18];
pub type G2Affine = [u64; 24];
pub type G2Projective = [u64; 36];

pub const SUCCESS: c_int = 0;
pub const NTT_FORWARD: c_int = 0;
pub const NTT_INVERSE: c_int = 1;

/// icicle::Ordering
#[repr(i32)]
#[derive(Clone, Copy, Debug, PartialEq, Eq)]
pub enum Ordering { NN = 0, NR = 1, RN = 2, RR = 3, NM = 4, MN = 5 }

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct MsmConfig {
    pub stream: *mut c_void,
    pub precompute_factor: c_int,
    pub c: c_int,
    pub bitsize: c_int,
    pub batch_size: c_int,
    pub are_points_shared_in_batch: bool,
    pub are_scalars_on_device: bool,
    pub are_scalars_montgomery_form: bool,
    pub are_points_on_device: bool,
    pub are_points_montgomery_form: bool,
    pub are_results_on_device: bool,
    pub is_async: bool,
    pub ext: *mut c_void,
}

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct NttConfig {
    pub stream: *mut c_void,
    pub coset_gen: Fr,
    pub batch_size: c_int,
    pub columns_batch: bool,
    pub ordering: Ordering,
    pub are_inputs_on_device: bool,
    pub are_outputs_on_device: bool,
    pub is_async: bool,
    pub ext: *mut c_void,
}

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct NttInitDomainConfig { pub stream: *mut c_void, pub is_async: bool, pub ext: *mut c_void }

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct VecOpsConfig {
    pub stream: *mut c_void,
    pub is_a_on_device: bool,
    pub is_b_on_device: bool,
    pub is_result_on_device: bool,
    pub is_async: bool,
    pub ext: *mut c_void,
}

extern "C" {
    pub fn b381_default_msm_config() -> MsmConfig;
    pub fn b381_default_ntt_config() -> NttConfig;
    pub fn b381_default_vecops_config() -> VecOpsConfig;

    // MSM (replaces msm_cuda_impl / msm_g2_cuda_impl / *_precompute_bases_*, icicle_curve_api.cu:243-650)
    pub fn b381_g1_msm(scalars: *const Fr, bases: *const G1Affine, msm_size: c_int, cfg: *const MsmConfig, results: *mut G1Projective) -> c_int;
    pub fn b381_g2_msm(scalars: *const Fr, bases: *const G2Affine, msm_size: c_int, cfg: *const MsmConfig, results: *mut G2Projective) -> c_int;
    pub fn b381_g1_msm_precompute_bases(input: *const G1Affine, n: c_int, cfg: *const MsmConfig, output: *mut G1Affine) -> c_int;
    pub fn b381_g2_msm_precompute_bases(input: *const G2Affine, n: c_int, cfg: *const MsmConfig, output: *mut G2Affine) -> c_int;
    // multi-GPU: one partial per GPU (192 / 384 bytes, device memory), combined after an all-gather
    pub fn b381_g1_msm_partial(scalars: *const Fr, bases: *const G1Affine, n: c_int, cfg: *const MsmConfig, out: *mut c_void) -> c_int;
    pub fn b381_g1_msm_combine(parts: *const c_void, count: c_int, stream: *mut c_void, on_device: bool, result: *mut G1Projective) -> c_int;

    // NTT (replaces ntt_cuda_impl and the domain callbacks, icicle_field_api.cu:97-131)
    pub fn b381_ntt_init_domain(root_of_unity: *const Fr, cfg: *const NttInitDomainConfig) -> c_int;
    pub fn b381_ntt_release_domain() -> c_int;
    pub fn b381_ntt_get_rou_from_domain(log_size: u64, rou: *mut Fr) -> c_int;
    pub fn b381_ntt(input: *const Fr, size: c_int, dir: c_int, cfg: *const NttConfig, output: *mut Fr) -> c_int;

    // vector ops (icicle_field_api.cu:133-334) + the unregistered ones core/vecops.rs uses
    pub fn b381_vector_add(a: *const Fr, b: *const Fr, n: u64, cfg: *const VecOpsConfig, out: *mut Fr) -> c_int;
    pub fn b381_vector_sub(a: *const Fr, b: *const Fr, n: u64, cfg: *const VecOpsConfig, out: *mut Fr) -> c_int;
    pub fn b381_vector_mul(a: *const Fr, b: *const Fr, n: u64, cfg: *const VecOpsConfig, out: *mut Fr) -> c_int;
    pub fn b381_scalar_mul_vec(scalar: *const Fr, b: *const Fr, n: u64, cfg: *const VecOpsConfig, out: *mut Fr) -> c_int;
    pub fn b381_scalar_add_vec(scalar: *const Fr, b: *const Fr, n: u64, cfg: *const VecOpsConfig, out: *mut Fr) -> c_int;
    pub fn b381_bit_reverse(a: *const Fr, n: u64, cfg: *const VecOpsConfig, out: *mut Fr) -> c_int;
    pub fn b381_vector_sum(a: *const Fr, n: u64, cfg: *const VecOpsConfig, out: *mut Fr) -> c_int;
    pub fn b381_vector_inv(a: *const Fr, n: u64, cfg: *const VecOpsConfig, out: *mut Fr) -> c_int;
    pub fn b381_montgomery_convert(a: *const Fr, n: u64, to_montgomery: c_int, cfg: *const VecOpsConfig, out: *mut Fr) -> c_int;

    // point validation on ingest
    pub fn b381_g1_is_on_curve(input: *const G1Affine, n: c_int, cfg: *const VecOpsConfig, flags: *mut u8) -> c_int;
    pub fn b381_g1_is_in_subgroup(input: *const G1Affine, n: c_int, cfg: *const VecOpsConfig, flags: *mut u8) -> c_int;
    pub fn b381_g2_is_on_curve(input: *const G2Affine, n: c_int, cfg: *const VecOpsConfig, flags: *mut u8) -> c_int;
    pub fn b381_g2_is_in_subgroup(input: *const G2Affine, n: c_int, cfg: *const VecOpsConfig, flags: *mut u8) -> c_int;

    // device plumbing (CudaDeviceAPI, src/device/cuda_device_api.cu:38-149)
    pub fn b381_device_count(count: *mut c_int) -> c_int;
    pub fn b381_set_device(device_id: c_int) -> c_int;
    pub fn b381_malloc(ptr: *mut *mut c_void, size: usize) -> c_int;
    pub fn b381_free(ptr: *mut c_void) -> c_int;
    pub fn b381_copy_to_device(dst: *mut c_void, src: *const c_void, size: usize) -> c_int;
    pub fn b381_copy_to_host(dst: *mut c_void, src: *const c_void, size: usize) -> c_int;
    pub fn b381_copy_to_device_async(dst: *mut c_void, src: *const c_void, size: usize, stream: *mut c_void) -> c_int;
    pub fn b381_copy_to_host_async(dst: *mut c_void, src: *const c_void, size: usize, stream: *mut c_void) -> c_int;
    pub fn b381_host_alloc_pinned(ptr: *mut *mut c_void, size: usize) -> c_int;
    pub fn b381_host_free_pinned(ptr: *mut c_void) -> c_int;
    pub fn b381_stream_create(stream: *mut *mut c_void) -> c_int;
    pub fn b381_stream_destroy(stream: *mut c_void) -> c_int;
    pub fn b381_stream_synchronize(stream: *mut c_void) -> c_int;
}
