//! `GpuNttContext` / `NttHandle` -- the API of core/ntt.rs over `b381_ntt*`.  The `*_auto` dispatchers of the reference
//! (core/ntt.rs:1879-1990) chose between CPU and GPU by size; there is no CPU path here, so they are plain aliases.
//! Orderings and coset generators are honoured by the backend (the reference's registered NTT ignores both,
//! icicle_field_api.cu:97-131).
use crate::{ffi, stream::{check, DeviceVec, GpuError, ManagedStream, PinnedVec}, types::*};

#[derive(Debug)]
pub enum NttError {
    InvalidSize(usize),
    SizeExceedsDomain { size: usize, max_log_size: u32 },
    ExecutionFailed(String),
    Gpu(GpuError),
}
impl std::fmt::Display for NttError {
    fn fmt(&self, f: &mut std::fmt::Formatter<'_>) -> std::fmt::Result { write!(f, "{self:?}") }
}
impl std::error::Error for NttError {}
impl From<GpuError> for NttError {
    fn from(e: GpuError) -> Self { NttError::Gpu(e) }
}

pub struct GpuNttContext { max_log_size: u32 }

/// 7^((r-1)/2^32) in STANDARD form: the 2^32-th root of unity every halo2/ICICLE stack uses; the domain root of size
/// 2^k is its 2^(32-k)-th power.  The backend accepts the root in standard or Montgomery form and finds its order.
const ROOT_OF_UNITY_2_32: ffi::Fr = [0x3829971f439f0d2b, 0xb63683508c2280b9, 0xd09b681922c813b4, 0x16a2a19edfe81f20];

impl GpuNttContext {
    /// core/ntt.rs:342-451: initialises the backend's twiddle domain for transforms up to 2^max_log_size.
    pub fn new(max_log_size: u32) -> Result<Self, NttError> {
        crate::stream::ensure_backend_loaded()?;
        if max_log_size == 0 || max_log_size > 27 { return Err(NttError::InvalidSize(max_log_size as usize)); }
        let root = Self::pow2k(ROOT_OF_UNITY_2_32, 32 - max_log_size)?;
        let cfg = ffi::NttInitDomainConfig { stream: std::ptr::null_mut(), is_async: false, ext: std::ptr::null_mut() };
        check(unsafe { ffi::b381_ntt_init_domain(&root, &cfg) })?;
        Ok(Self { max_log_size })
    }
    pub fn max_log_size(&self) -> u32 { self.max_log_size }

    /// x^(2^k) with the device's own multiplier (standard form in, standard form out): convert to Montgomery, square k
    /// times with vector_mul on one element, convert back.  Runs once per context.
    fn pow2k(x: ffi::Fr, k: u32) -> Result<ffi::Fr, NttError> {
        let cfg = unsafe { ffi::b381_default_vecops_config() };
        let mut v = [x];
        let mut t = [[0u64; 4]];
        check(unsafe { ffi::b381_montgomery_convert(v.as_ptr(), 1, 1, &cfg, t.as_mut_ptr()) })?;
        for _ in 0..k {
            check(unsafe { ffi::b381_vector_mul(t.as_ptr(), t.as_ptr(), 1, &cfg, v.as_mut_ptr()) })?;
            t = v;
        }
        check(unsafe { ffi::b381_montgomery_convert(t.as_ptr(), 1, 0, &cfg, v.as_mut_ptr()) })?;
        Ok(v[0])
    }

    fn check_size(&self, size: usize) -> Result<(), NttError> {
        if size == 0 || !size.is_power_of_two() { return Err(NttError::InvalidSize(size)); }
        if size > (1usize << self.max_log_size) {
            return Err(NttError::SizeExceedsDomain { size, max_log_size: self.max_log_size });
        }
        Ok(())
    }

    fn cfg(batch: usize, coset: Option<ffi::Fr>, stream: Option<&ManagedStream>, on_device: bool) -> ffi::NttConfig {
        let mut cfg = unsafe { ffi::b381_default_ntt_config() };
        cfg.batch_size = batch as i32;
        cfg.ordering = ffi::Ordering::NN;
        cfg.are_inputs_on_device = on_device;
        cfg.are_outputs_on_device = on_device;
        if let Some(g) = coset { cfg.coset_gen = g; }       // Montgomery form, like every scalar on this API
        if let Some(s) = stream { cfg.stream = s.handle(); cfg.is_async = true; }
        cfg
    }

    fn run_host(&self, input: &[ffi::Fr], poly_size: usize, dir: i32, coset: Option<ffi::Fr>, out: &mut [ffi::Fr]) -> Result<(), NttError> {
        self.check_size(poly_size)?;
        if input.len() % poly_size != 0 || out.len() != input.len() { return Err(NttError::InvalidSize(input.len())); }
        let cfg = Self::cfg(input.len() / poly_size, coset, None, false);
        check(unsafe { ffi::b381_ntt(input.as_ptr(), poly_size as i32, dir, &cfg, out.as_mut_ptr()) })?;
        Ok(())
    }

    // ---- host vectors (core/ntt.rs:453-489, :1041-1060, :1228-1390)
    pub fn forward_ntt<S: PodScalar>(&self, coefficients: &[S]) -> Result<Vec<S>, NttError> { self.batch(coefficients, coefficients.len(), ffi::NTT_FORWARD, None) }
    pub fn inverse_ntt<S: PodScalar>(&self, evaluations: &[S]) -> Result<Vec<S>, NttError> { self.batch(evaluations, evaluations.len(), ffi::NTT_INVERSE, None) }
    pub fn forward_ntt_batch<S: PodScalar>(&self, batch: &[S], poly_size: usize) -> Result<Vec<S>, NttError> { self.batch(batch, poly_size, ffi::NTT_FORWARD, None) }
    pub fn inverse_ntt_batch<S: PodScalar>(&self, batch: &[S], poly_size: usize) -> Result<Vec<S>, NttError> { self.batch(batch, poly_size, ffi::NTT_INVERSE, None) }
    pub fn forward_coset_ntt<S: PodScalar>(&self, coefficients: &[S], coset_gen: S) -> Result<Vec<S>, NttError> {
        self.batch(coefficients, coefficients.len(), ffi::NTT_FORWARD, Some(TypeConverter::scalar_slice_as_icicle(&[coset_gen])[0]))
    }
    pub fn inverse_coset_ntt<S: PodScalar>(&self, evaluations: &[S], coset_gen: S) -> Result<Vec<S>, NttError> {
        self.batch(evaluations, evaluations.len(), ffi::NTT_INVERSE, Some(TypeConverter::scalar_slice_as_icicle(&[coset_gen])[0]))
    }
    pub fn forward_coset_ntt_batch<S: PodScalar>(&self, batch: &[S], poly_size: usize, coset_gen: S) -> Result<Vec<S>, NttError> {
        self.batch(batch, poly_size, ffi::NTT_FORWARD, Some(TypeConverter::scalar_slice_as_icicle(&[coset_gen])[0]))
    }
    pub fn inverse_coset_ntt_batch<S: PodScalar>(&self, batch: &[S], poly_size: usize, coset_gen: S) -> Result<Vec<S>, NttError> {
        self.batch(batch, poly_size, ffi::NTT_INVERSE, Some(TypeConverter::scalar_slice_as_icicle(&[coset_gen])[0]))
    }
    fn batch<S: PodScalar>(&self, data: &[S], poly_size: usize, dir: i32, coset: Option<ffi::Fr>) -> Result<Vec<S>, NttError> {
        let input = TypeConverter::scalar_slice_as_icicle(data);
        let mut out = vec![[0u64; 4]; input.len()];
        self.run_host(input, poly_size, dir, coset, &mut out)?;
        Ok(TypeConverter::icicle_slice_as_scalars::<S>(&out).to_vec())
    }
    pub fn forward_ntt_inplace<S: PodScalar>(&self, data: &mut [S]) -> Result<(), NttError> { let n = data.len(); self.inplace(data, n, ffi::NTT_FORWARD, None) }
    pub fn inverse_ntt_inplace<S: PodScalar>(&self, data: &mut [S]) -> Result<(), NttError> { let n = data.len(); self.inplace(data, n, ffi::NTT_INVERSE, None) }
    pub fn forward_ntt_batch_inplace<S: PodScalar>(&self, batch: &mut [S], poly_size: usize) -> Result<(), NttError> { self.inplace(batch, poly_size, ffi::NTT_FORWARD, None) }
    pub fn inverse_ntt_batch_inplace<S: PodScalar>(&self, batch: &mut [S], poly_size: usize) -> Result<(), NttError> { self.inplace(batch, poly_size, ffi::NTT_INVERSE, None) }
    fn inplace<S: PodScalar>(&self, data: &mut [S], poly_size: usize, dir: i32, coset: Option<ffi::Fr>) -> Result<(), NttError> {
        self.check_size(poly_size)?;
        let v = TypeConverter::scalar_slice_as_icicle_mut(data);
        if v.len() % poly_size != 0 { return Err(NttError::InvalidSize(v.len())); }
        let cfg = Self::cfg(v.len() / poly_size, coset, None, false);
        let p = v.as_mut_ptr();
        check(unsafe { ffi::b381_ntt(p, poly_size as i32, dir, &cfg, p) })?;
        Ok(())
    }

    // ---- device-resident vectors (core/ntt.rs:610-760): in place, no PCIe traffic
    pub fn ntt_on_device(&self, data: &mut DeviceVec<ffi::Fr>, dir_inverse: bool) -> Result<(), NttError> {
        let n = data.len();
        self.ntt_batch_on_device(data, n, dir_inverse)
    }
    pub fn forward_ntt_on_device(&self, data: &mut DeviceVec<ffi::Fr>) -> Result<(), NttError> { self.ntt_on_device(data, false) }
    pub fn inverse_ntt_on_device(&self, data: &mut DeviceVec<ffi::Fr>) -> Result<(), NttError> { self.ntt_on_device(data, true) }
    pub fn ntt_batch_on_device(&self, data: &mut DeviceVec<ffi::Fr>, poly_size: usize, dir_inverse: bool) -> Result<(), NttError> {
        self.device(data, poly_size, dir_inverse, None, None)
    }
    pub fn coset_ntt_on_device(&self, data: &mut DeviceVec<ffi::Fr>, coset_gen: ffi::Fr, dir_inverse: bool) -> Result<(), NttError> {
        let n = data.len();
        self.device(data, n, dir_inverse, Some(coset_gen), None)
    }
    pub fn coset_ntt_batch_on_device(&self, data: &mut DeviceVec<ffi::Fr>, poly_size: usize, coset_gen: ffi::Fr, dir_inverse: bool) -> Result<(), NttError> {
        self.device(data, poly_size, dir_inverse, Some(coset_gen), None)
    }
    /// core/ntt.rs:831-919: returns at once; synchronise `stream` before reading `data`
    pub fn ntt_on_device_async(&self, data: &mut DeviceVec<ffi::Fr>, dir_inverse: bool, stream: &ManagedStream) -> Result<(), NttError> {
        let n = data.len();
        self.device(data, n, dir_inverse, None, Some(stream))
    }
    pub fn ntt_batch_on_device_async(&self, data: &mut DeviceVec<ffi::Fr>, poly_size: usize, dir_inverse: bool, stream: &ManagedStream) -> Result<(), NttError> {
        self.device(data, poly_size, dir_inverse, None, Some(stream))
    }
    fn device(&self, data: &mut DeviceVec<ffi::Fr>, poly_size: usize, inverse: bool, coset: Option<ffi::Fr>, stream: Option<&ManagedStream>) -> Result<(), NttError> {
        self.check_size(poly_size)?;
        if data.len() % poly_size != 0 { return Err(NttError::InvalidSize(data.len())); }
        let cfg = Self::cfg(data.len() / poly_size, coset, stream, true);
        let p = data.as_mut_ptr();
        check(unsafe { ffi::b381_ntt(p, poly_size as i32, if inverse { ffi::NTT_INVERSE } else { ffi::NTT_FORWARD }, &cfg, p) })?;
        Ok(())
    }

    // ---- async over host vectors (core/ntt.rs:945-1040): result lands in pinned memory
    pub fn forward_ntt_async<'a, S: PodScalar>(&self, coefficients: &'a [S]) -> Result<NttHandle<'a>, NttError> { self.host_async(coefficients, ffi::NTT_FORWARD) }
    pub fn inverse_ntt_async<'a, S: PodScalar>(&self, evaluations: &'a [S]) -> Result<NttHandle<'a>, NttError> { self.host_async(evaluations, ffi::NTT_INVERSE) }
    fn host_async<'a, S: PodScalar>(&self, data: &'a [S], dir: i32) -> Result<NttHandle<'a>, NttError> {
        let input = TypeConverter::scalar_slice_as_icicle(data);
        self.check_size(input.len())?;
        let stream = ManagedStream::create()?;
        let mut result = PinnedVec::<ffi::Fr>::zeroed(input.len())?;
        let cfg = Self::cfg(1, None, Some(&stream), false);
        check(unsafe { ffi::b381_ntt(input.as_ptr(), input.len() as i32, dir, &cfg, result.as_mut_ptr()) })?;
        Ok(NttHandle { stream, result, _borrow: std::marker::PhantomData })
    }
}

impl Drop for GpuNttContext {
    fn drop(&mut self) {
        // the domain is process-wide state of the backend (one per device), like ICICLE's; it is left in place so
        // that other contexts keep working -- call release_domain() explicitly to free the twiddle table
    }
}
pub fn release_domain() -> Result<(), NttError> { Ok(check(unsafe { ffi::b381_ntt_release_domain() })?) }

pub struct NttHandle<'a> { stream: ManagedStream, result: PinnedVec<ffi::Fr>, _borrow: std::marker::PhantomData<&'a ()> }
impl<'a> NttHandle<'a> {
    pub fn size(&self) -> usize { self.result.as_slice().len() }
    pub fn wait<S: PodScalar>(mut self) -> Result<Vec<S>, NttError> {
        self.stream.synchronize()?;
        self.stream.destroy()?;
        Ok(TypeConverter::icicle_slice_as_scalars::<S>(self.result.as_slice()).to_vec())
    }
}

// the reference's size-dispatching entry points (core/ntt.rs:1879-1990), now always the device
pub fn forward_ntt_auto<S: PodScalar>(input: &[S]) -> Result<Vec<S>, NttError> { ctx_for(input.len())?.forward_ntt(input) }
pub fn inverse_ntt_auto<S: PodScalar>(input: &[S]) -> Result<Vec<S>, NttError> { ctx_for(input.len())?.inverse_ntt(input) }
pub fn forward_ntt_inplace_auto<S: PodScalar>(data: &mut [S]) -> Result<(), NttError> { ctx_for(data.len())?.forward_ntt_inplace(data) }
pub fn inverse_ntt_inplace_auto<S: PodScalar>(data: &mut [S]) -> Result<(), NttError> { ctx_for(data.len())?.inverse_ntt_inplace(data) }
pub fn forward_ntt_batch_auto<S: PodScalar>(batch: &[S], poly_size: usize) -> Result<Vec<S>, NttError> { ctx_for(poly_size)?.forward_ntt_batch(batch, poly_size) }
pub fn inverse_ntt_batch_auto<S: PodScalar>(batch: &[S], poly_size: usize) -> Result<Vec<S>, NttError> { ctx_for(poly_size)?.inverse_ntt_batch(batch, poly_size) }

/// process-wide context sized for the largest transform asked for so far (core/ntt.rs:289-293 keeps an OnceLock)
fn ctx_for(size: usize) -> Result<std::sync::Arc<GpuNttContext>, NttError> {
    use std::sync::{Arc, Mutex, OnceLock};
    static CTX: OnceLock<Mutex<Option<Arc<GpuNttContext>>>> = OnceLock::new();
    if size == 0 || !size.is_power_of_two() { return Err(NttError::InvalidSize(size)); }
    let need = size.trailing_zeros().max(20);
    let mut g = CTX.get_or_init(|| Mutex::new(None)).lock().unwrap();
    if let Some(c) = g.as_ref() {
        if c.max_log_size() >= need { return Ok(c.clone()); }
        release_domain()?;
    }
    let c = Arc::new(GpuNttContext::new(need)?);
    *g = Some(c.clone());
    Ok(c)
}
