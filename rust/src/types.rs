//! Zero-copy views between the caller's curve types and the wire layouts of the C ABI -- the role of
//! `TypeConverter` in core/types.rs:89-368.  Wire layouts (little-endian u64 limbs):
//!   scalar  32 B  Montgomery (R = 2^256)                     == midnight_curves::Fq        (types.rs:148-152)
//!   G1 affine 96 B (x, y) Montgomery, infinity = (0, 0)      == midnight_curves::G1Affine  (types.rs:167-200)
//!   G2 affine 192 B (x.c0, x.c1, y.c0, y.c1)                 == midnight_curves::G2Affine
//!   results: ICICLE projective (x, y, z) in STANDARD form, (x, y, 1) or (0, 1, 0)          (types.rs:353-368)
//! The marker traits below are what a consumer implements (one line each) for its own types; they assert size and
//! alignment at compile time, as the reference does with `static_assertions`.
use crate::ffi;

/// # Safety
/// `Self` must be exactly 32 bytes, 8-byte aligned: four little-endian u64 limbs of a Montgomery-form Fr element.
pub unsafe trait PodScalar: Copy {}
/// # Safety
/// 96 bytes: (x, y) of 6 u64 limbs each, Montgomery form, infinity = all zero.
pub unsafe trait PodG1Affine: Copy {}
/// # Safety
/// 192 bytes: (x.c0, x.c1, y.c0, y.c1).
pub unsafe trait PodG2Affine: Copy {}

unsafe impl PodScalar for ffi::Fr {}
unsafe impl PodG1Affine for ffi::G1Affine {}
unsafe impl PodG2Affine for ffi::G2Affine {}

pub struct TypeConverter;

impl TypeConverter {
    #[inline]
    pub fn scalar_slice_as_icicle<S: PodScalar>(s: &[S]) -> &[ffi::Fr] {
        const { assert!(std::mem::size_of::<S>() == 32 && std::mem::align_of::<S>() <= 8) };
        unsafe { std::slice::from_raw_parts(s.as_ptr() as *const ffi::Fr, s.len()) }
    }
    #[inline]
    pub fn scalar_slice_as_icicle_mut<S: PodScalar>(s: &mut [S]) -> &mut [ffi::Fr] {
        const { assert!(std::mem::size_of::<S>() == 32 && std::mem::align_of::<S>() <= 8) };
        unsafe { std::slice::from_raw_parts_mut(s.as_mut_ptr() as *mut ffi::Fr, s.len()) }
    }
    #[inline]
    pub fn icicle_slice_as_scalars<S: PodScalar>(s: &[ffi::Fr]) -> &[S] {
        const { assert!(std::mem::size_of::<S>() == 32 && std::mem::align_of::<S>() <= 8) };
        unsafe { std::slice::from_raw_parts(s.as_ptr() as *const S, s.len()) }
    }
    #[inline]
    pub fn g1_slice_as_icicle<P: PodG1Affine>(p: &[P]) -> &[ffi::G1Affine] {
        const { assert!(std::mem::size_of::<P>() == 96 && std::mem::align_of::<P>() <= 8) };
        unsafe { std::slice::from_raw_parts(p.as_ptr() as *const ffi::G1Affine, p.len()) }
    }
    #[inline]
    pub fn g2_slice_as_icicle<P: PodG2Affine>(p: &[P]) -> &[ffi::G2Affine] {
        const { assert!(std::mem::size_of::<P>() == 192 && std::mem::align_of::<P>() <= 8) };
        unsafe { std::slice::from_raw_parts(p.as_ptr() as *const ffi::G2Affine, p.len()) }
    }
}

/// MSM result as the backend writes it: standard-form (x, y) with `infinity` for (0, 1, 0).  Converting to the
/// consumer's projective type is `from_raw_unchecked(x.to_montgomery(), ...)` on its side (types.rs:353-368).
#[derive(Clone, Copy, Debug, PartialEq, Eq)]
pub struct G1Result { pub x: [u64; 6], pub y: [u64; 6], pub infinity: bool }
#[derive(Clone, Copy, Debug, PartialEq, Eq)]
pub struct G2Result { pub x: [[u64; 6]; 2], pub y: [[u64; 6]; 2], pub infinity: bool }

impl G1Result {
    pub fn from_icicle(p: &ffi::G1Projective) -> Self {
        let mut x = [0u64; 6];
        let mut y = [0u64; 6];
        x.copy_from_slice(&p[0..6]);
        y.copy_from_slice(&p[6..12]);
        Self { x, y, infinity: p[12..18].iter().all(|&w| w == 0) }
    }
}
impl G2Result {
    pub fn from_icicle(p: &ffi::G2Projective) -> Self {
        let mut v = [[0u64; 6]; 4];
        for (i, c) in v.iter_mut().enumerate() { c.copy_from_slice(&p[6 * i..6 * i + 6]); }
        Self { x: [v[0], v[1]], y: [v[2], v[3]], infinity: p[24..36].iter().all(|&w| w == 0) }
    }
}
