//! `GpuMsmContext`, `PrecomputedBases`, `MsmHandle`, `G2MsmHandle`, `BatchMsmHandle` -- the API of core/msm.rs, over
//! `b381_g1_msm` / `b381_g2_msm` / `b381_g1_msm_precompute_bases` instead of `icicle_core::msm::*`.
//! Differences from the reference, all in the caller's favour: `precompute_bases` really expands the table (the
//! reference's backend copies, icicle_curve_api.cu:415-440); a batch is ONE pipeline run on the device; `is_async`
//! calls return without synchronising (the reference's MSM syncs in its cleanup path).
use crate::{ffi, stream::{check, DeviceVec, GpuError, ManagedStream, PinnedVec}, types::*};

#[derive(Debug)]
pub enum MsmError {
    EmptyInput,
    LengthMismatch { scalars: usize, points: usize },
    InvalidPrecompute(String),
    ExecutionFailed(String),
    Gpu(GpuError),
}
impl std::fmt::Display for MsmError {
    fn fmt(&self, f: &mut std::fmt::Formatter<'_>) -> std::fmt::Result { write!(f, "{self:?}") }
}
impl std::error::Error for MsmError {}
impl From<GpuError> for MsmError {
    fn from(e: GpuError) -> Self { MsmError::Gpu(e) }
}

/// Device-resident G1 bases, optionally expanded by `precompute_factor` (core/msm.rs:174-264).  The stored multiples
/// are interleaved (`bases[i * factor + k]`), upstream ICICLE's layout, so an MSM over the first n' <= n points
/// uses a prefix of the same buffer; `window` is the window size the table was built for.
pub struct PrecomputedBases { buffer: DeviceVec<ffi::G1Affine>, original: usize, factor: i32, window: i32 }

impl PrecomputedBases {
    pub fn new(buffer: DeviceVec<ffi::G1Affine>, size: usize) -> Self { Self { buffer, original: size, factor: 1, window: 0 } }
    pub fn new_precomputed(buffer: DeviceVec<ffi::G1Affine>, original: usize, factor: i32, window: i32) -> Self {
        Self { buffer, original, factor, window }
    }
    pub fn is_precomputed(&self) -> bool { self.factor > 1 }
    pub fn factor(&self) -> i32 { self.factor }
    pub fn original_size(&self) -> usize { self.original }
    pub fn buffer_size(&self) -> usize { self.buffer.len() }
    pub fn len(&self) -> usize { self.original }
    pub fn is_empty(&self) -> bool { self.original == 0 }
    pub fn device_ptr(&self) -> *const ffi::G1Affine { self.buffer.as_ptr() }
    pub fn required_size_for_scalars(&self, num_scalars: usize) -> Result<usize, MsmError> {
        if num_scalars > self.original {
            return Err(MsmError::LengthMismatch { scalars: num_scalars, points: self.original });
        }
        Ok(num_scalars * self.factor as usize)
    }
}

pub struct GpuMsmContext { device_id: i32, window: i32 }

impl GpuMsmContext {
    pub fn new() -> Result<Self, MsmError> { Self::with_device(0, 0) }
    /// `window` = MIDNIGHT_MSM_WINDOW of core/config.rs as a plain parameter; 0 = the backend's measured table.
    pub fn with_device(device_id: i32, window: i32) -> Result<Self, MsmError> {
        crate::stream::ensure_backend_loaded()?;
        crate::stream::set_device(device_id)?;
        Ok(Self { device_id, window })
    }
    pub fn device_id(&self) -> i32 { self.device_id }

    fn cfg(&self, bases: Option<&PrecomputedBases>, stream: Option<&ManagedStream>) -> ffi::MsmConfig {
        let mut cfg = unsafe { ffi::b381_default_msm_config() };
        cfg.are_scalars_montgomery_form = true;        // the caller's scalars are Montgomery (types.rs)
        cfg.are_points_montgomery_form = true;
        cfg.c = self.window;
        if let Some(b) = bases {
            cfg.are_points_on_device = true;
            cfg.precompute_factor = b.factor;
            if b.is_precomputed() { cfg.c = b.window; }
        }
        if let Some(s) = stream {
            cfg.stream = s.handle();
            cfg.is_async = true;
        }
        cfg
    }

    pub fn upload_g1_bases<P: PodG1Affine>(&self, points: &[P]) -> Result<PrecomputedBases, MsmError> {
        if points.is_empty() { return Err(MsmError::EmptyInput); }
        let pts = TypeConverter::g1_slice_as_icicle(points);
        Ok(PrecomputedBases::new(DeviceVec::from_host(pts)?, pts.len()))
    }
    pub fn upload_g2_bases<P: PodG2Affine>(&self, points: &[P]) -> Result<DeviceVec<ffi::G2Affine>, MsmError> {
        if points.is_empty() { return Err(MsmError::EmptyInput); }
        Ok(DeviceVec::from_host(TypeConverter::g2_slice_as_icicle(points))?)
    }

    /// core/msm.rs:401-492
    pub fn precompute_bases(&self, bases: &PrecomputedBases, factor: i32) -> Result<PrecomputedBases, MsmError> {
        if factor <= 1 || bases.is_precomputed() {
            return Err(MsmError::InvalidPrecompute("factor must be > 1 on plain bases".into()));
        }
        let n = bases.original_size();
        let mut out = DeviceVec::<ffi::G1Affine>::device_malloc(n * factor as usize)?;
        let mut cfg = self.cfg(Some(bases), None);
        cfg.precompute_factor = factor;
        cfg.are_results_on_device = true;
        cfg.c = if self.window > 0 { self.window } else { 16 };     // fixed window: the table must not depend on n
        check(unsafe { ffi::b381_g1_msm_precompute_bases(bases.device_ptr(), n as i32, &cfg, out.as_mut_ptr()) })?;
        Ok(PrecomputedBases::new_precomputed(out, n, factor, cfg.c))
    }
    pub fn upload_g1_bases_with_precompute<P: PodG1Affine>(&self, points: &[P], factor: i32) -> Result<PrecomputedBases, MsmError> {
        let plain = self.upload_g1_bases(points)?;
        if factor > 1 { self.precompute_bases(&plain, factor) } else { Ok(plain) }
    }

    /// core/msm.rs:519-590: host scalars, host points
    pub fn msm<S: PodScalar, P: PodG1Affine>(&self, scalars: &[S], points: &[P]) -> Result<G1Result, MsmError> {
        if scalars.is_empty() { return Err(MsmError::EmptyInput); }
        if scalars.len() != points.len() { return Err(MsmError::LengthMismatch { scalars: scalars.len(), points: points.len() }); }
        let (sc, pts) = (TypeConverter::scalar_slice_as_icicle(scalars), TypeConverter::g1_slice_as_icicle(points));
        let cfg = self.cfg(None, None);
        let mut out = [0u64; 18];
        check(unsafe { ffi::b381_g1_msm(sc.as_ptr(), pts.as_ptr(), sc.len() as i32, &cfg, &mut out) })?;
        Ok(G1Result::from_icicle(&out))
    }

    /// core/msm.rs:594-713: host scalars (copied in pieces under the sort), resident bases
    pub fn msm_with_device_bases<S: PodScalar>(&self, scalars: &[S], bases: &PrecomputedBases) -> Result<G1Result, MsmError> {
        if scalars.is_empty() { return Err(MsmError::EmptyInput); }
        bases.required_size_for_scalars(scalars.len())?;
        let sc = TypeConverter::scalar_slice_as_icicle(scalars);
        let cfg = self.cfg(Some(bases), None);
        let mut out = [0u64; 18];
        check(unsafe { ffi::b381_g1_msm(sc.as_ptr(), bases.device_ptr(), sc.len() as i32, &cfg, &mut out) })?;
        Ok(G1Result::from_icicle(&out))
    }

    /// core/msm.rs:715-798.  The scalars must stay alive and unmodified until `wait()`: the handle borrows them.
    pub fn msm_with_device_bases_async<'a, S: PodScalar>(&self, scalars: &'a [S], bases: &'a PrecomputedBases)
        -> Result<MsmHandle<'a>, MsmError> {
        if scalars.is_empty() { return Err(MsmError::EmptyInput); }
        bases.required_size_for_scalars(scalars.len())?;
        let sc = TypeConverter::scalar_slice_as_icicle(scalars);
        let stream = ManagedStream::create()?;
        let mut result = PinnedVec::<ffi::G1Projective>::zeroed(1)?;
        let cfg = self.cfg(Some(bases), Some(&stream));
        check(unsafe { ffi::b381_g1_msm(sc.as_ptr(), bases.device_ptr(), sc.len() as i32, &cfg, result.as_mut_ptr()) })?;
        Ok(MsmHandle { stream, result, _borrow: std::marker::PhantomData })
    }

    /// core/msm.rs:1179-1295: B scalar vectors of equal length over one base set, one backend call = one pipeline run
    pub fn msm_batch_with_device_bases<S: PodScalar>(&self, scalars_batch: &[&[S]], bases: &PrecomputedBases)
        -> Result<Vec<G1Result>, MsmError> {
        let flat = Self::flatten(scalars_batch)?;
        let n = scalars_batch[0].len();
        bases.required_size_for_scalars(n)?;
        let mut cfg = self.cfg(Some(bases), None);
        cfg.batch_size = scalars_batch.len() as i32;
        cfg.are_points_shared_in_batch = true;
        let mut out = vec![[0u64; 18]; scalars_batch.len()];
        check(unsafe { ffi::b381_g1_msm(flat.as_ptr(), bases.device_ptr(), n as i32, &cfg, out.as_mut_ptr()) })?;
        Ok(out.iter().map(G1Result::from_icicle).collect())
    }

    /// core/msm.rs:1314-1418
    pub fn msm_batch_with_device_bases_async<'a, S: PodScalar>(&self, scalars_batch: &[&[S]], bases: &'a PrecomputedBases)
        -> Result<BatchMsmHandle<'a>, MsmError> {
        let flat = Self::flatten(scalars_batch)?;
        let n = scalars_batch[0].len();
        bases.required_size_for_scalars(n)?;
        let stream = ManagedStream::create()?;
        let mut cfg = self.cfg(Some(bases), Some(&stream));
        cfg.batch_size = scalars_batch.len() as i32;
        cfg.are_points_shared_in_batch = true;
        let mut result = PinnedVec::<ffi::G1Projective>::zeroed(scalars_batch.len())?;
        check(unsafe { ffi::b381_g1_msm(flat.as_ptr(), bases.device_ptr(), n as i32, &cfg, result.as_mut_ptr()) })?;
        Ok(BatchMsmHandle { stream, result, _scalars: flat, _borrow: std::marker::PhantomData })
    }

    fn flatten<S: PodScalar>(batch: &[&[S]]) -> Result<Vec<ffi::Fr>, MsmError> {
        if batch.is_empty() || batch[0].is_empty() { return Err(MsmError::EmptyInput); }
        let n = batch[0].len();
        let mut flat = Vec::with_capacity(n * batch.len());
        for s in batch {
            if s.len() != n { return Err(MsmError::LengthMismatch { scalars: s.len(), points: n }); }
            flat.extend_from_slice(TypeConverter::scalar_slice_as_icicle(s));
        }
        Ok(flat)
    }

    /// core/msm.rs:800-859
    pub fn g2_msm<S: PodScalar, P: PodG2Affine>(&self, scalars: &[S], points: &[P]) -> Result<G2Result, MsmError> {
        if scalars.is_empty() { return Err(MsmError::EmptyInput); }
        if scalars.len() != points.len() { return Err(MsmError::LengthMismatch { scalars: scalars.len(), points: points.len() }); }
        let (sc, pts) = (TypeConverter::scalar_slice_as_icicle(scalars), TypeConverter::g2_slice_as_icicle(points));
        let cfg = self.cfg(None, None);
        let mut out = [0u64; 36];
        check(unsafe { ffi::b381_g2_msm(sc.as_ptr(), pts.as_ptr(), sc.len() as i32, &cfg, &mut out) })?;
        Ok(G2Result::from_icicle(&out))
    }
    /// core/msm.rs:861-929
    pub fn g2_msm_with_device_bases<S: PodScalar>(&self, scalars: &[S], bases: &DeviceVec<ffi::G2Affine>) -> Result<G2Result, MsmError> {
        if scalars.is_empty() { return Err(MsmError::EmptyInput); }
        if scalars.len() > bases.len() { return Err(MsmError::LengthMismatch { scalars: scalars.len(), points: bases.len() }); }
        let sc = TypeConverter::scalar_slice_as_icicle(scalars);
        let mut cfg = self.cfg(None, None);
        cfg.are_points_on_device = true;
        let mut out = [0u64; 36];
        check(unsafe { ffi::b381_g2_msm(sc.as_ptr(), bases.as_ptr(), sc.len() as i32, &cfg, &mut out) })?;
        Ok(G2Result::from_icicle(&out))
    }

    /// core/msm.rs:931-983: one tiny MSM pays context creation and module load once
    pub fn warmup(&self) -> Result<std::time::Duration, MsmError> {
        let t = std::time::Instant::now();
        let one: [ffi::Fr; 1] = [[1, 0, 0, 0]];
        let inf: [ffi::G1Affine; 1] = [[0; 12]];
        let mut cfg = self.cfg(None, None);
        cfg.are_scalars_montgomery_form = false;
        let mut out = [0u64; 18];
        check(unsafe { ffi::b381_g1_msm(one.as_ptr(), inf.as_ptr(), 1, &cfg, &mut out) })?;
        Ok(t.elapsed())
    }
}

pub struct MsmHandle<'a> { stream: ManagedStream, result: PinnedVec<ffi::G1Projective>, _borrow: std::marker::PhantomData<&'a ()> }
impl<'a> MsmHandle<'a> {
    pub fn wait(mut self) -> Result<G1Result, MsmError> {
        self.stream.synchronize()?;
        self.stream.destroy()?;
        Ok(G1Result::from_icicle(&self.result.as_slice()[0]))
    }
}

pub struct BatchMsmHandle<'a> {
    stream: ManagedStream,
    result: PinnedVec<ffi::G1Projective>,
    _scalars: Vec<ffi::Fr>,
    _borrow: std::marker::PhantomData<&'a ()>,
}
impl<'a> BatchMsmHandle<'a> {
    pub fn batch_size(&self) -> usize { self.result.as_slice().len() }
    pub fn wait(mut self) -> Result<Vec<G1Result>, MsmError> {
        self.stream.synchronize()?;
        self.stream.destroy()?;
        Ok(self.result.as_slice().iter().map(G1Result::from_icicle).collect())
    }
}
