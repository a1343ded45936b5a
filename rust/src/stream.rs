//! `GpuError`, `ManagedStream` (core/stream.rs:96-198), `DeviceVec` (the slice of icicle_runtime::memory the reference's
//! core/ uses) and `PinnedVec` (page-locked host buffer: an `is_async` call must not write its result into pageable
//! memory, or the copy blocks the calling thread until the whole call has run).
use crate::ffi;
use std::{marker::PhantomData, os::raw::c_void, ptr};

/// eIcicleError numbering (include/icicle/errors.h:37-53), as returned by every C-ABI entry point.
#[derive(Clone, Copy, Debug, PartialEq, Eq)]
pub struct GpuError(pub i32);

impl std::fmt::Display for GpuError {
    fn fmt(&self, f: &mut std::fmt::Formatter<'_>) -> std::fmt::Result {
        const NAMES: [&str; 15] = ["SUCCESS", "INVALID_DEVICE", "OUT_OF_MEMORY", "INVALID_POINTER", "ALLOCATION_FAILED",
            "DEALLOCATION_FAILED", "COPY_FAILED", "SYNCHRONIZATION_FAILED", "STREAM_CREATION_FAILED",
            "STREAM_DESTRUCTION_FAILED", "API_NOT_IMPLEMENTED", "INVALID_ARGUMENT", "BACKEND_LOAD_FAILED",
            "LICENSE_CHECK_ERROR", "UNKNOWN_ERROR"];
        match NAMES.get(self.0 as usize) {
            Some(n) => write!(f, "b381: {n}"),
            None => write!(f, "b381: error {}", self.0),
        }
    }
}
impl std::error::Error for GpuError {}

#[inline]
pub(crate) fn check(code: i32) -> Result<(), GpuError> {
    if code == ffi::SUCCESS { Ok(()) } else { Err(GpuError(code)) }
}

/// core/backend.rs:75-97.  The CUDA library is linked at build time; this only verifies that a device answers.
/// There is no CPU path: callers that used `dispatch_*` fallbacks get the error instead.
pub fn ensure_backend_loaded() -> Result<(), GpuError> {
    let mut n = 0;
    check(unsafe { ffi::b381_device_count(&mut n) })?;
    if n > 0 { Ok(()) } else { Err(GpuError(1)) }
}
pub fn is_gpu_available() -> bool { ensure_backend_loaded().is_ok() }
pub fn set_device(device_id: i32) -> Result<(), GpuError> { check(unsafe { ffi::b381_set_device(device_id) }) }

pub struct ManagedStream { handle: *mut c_void, owned: bool, destroyed: bool }
unsafe impl Send for ManagedStream {}

impl ManagedStream {
    pub fn create() -> Result<Self, GpuError> {
        let mut h = ptr::null_mut();
        check(unsafe { ffi::b381_stream_create(&mut h) })?;
        Ok(Self { handle: h, owned: true, destroyed: false })
    }
    pub fn default_stream() -> Self { Self { handle: ptr::null_mut(), owned: false, destroyed: false } }
    pub fn handle(&self) -> *mut c_void { self.handle }
    pub fn synchronize(&mut self) -> Result<(), GpuError> { check(unsafe { ffi::b381_stream_synchronize(self.handle) }) }
    pub fn destroy(&mut self) -> Result<(), GpuError> {
        if self.owned && !self.destroyed {
            self.destroyed = true;
            return check(unsafe { ffi::b381_stream_destroy(self.handle) });
        }
        self.destroyed = true;
        Ok(())
    }
    pub fn is_destroyed(&self) -> bool { self.destroyed }
}
impl Drop for ManagedStream {
    fn drop(&mut self) { let _ = self.destroy(); }
}

/// Owned device buffer of `len` elements of `T` (plain old data).
pub struct DeviceVec<T: Copy> { ptr: *mut c_void, len: usize, _t: PhantomData<T> }
unsafe impl<T: Copy> Send for DeviceVec<T> {}
unsafe impl<T: Copy> Sync for DeviceVec<T> {}

impl<T: Copy> DeviceVec<T> {
    pub fn device_malloc(len: usize) -> Result<Self, GpuError> {
        let mut p = ptr::null_mut();
        check(unsafe { ffi::b381_malloc(&mut p, (len * std::mem::size_of::<T>()).max(1)) })?;
        Ok(Self { ptr: p, len, _t: PhantomData })
    }
    pub fn from_host(src: &[T]) -> Result<Self, GpuError> {
        let mut v = Self::device_malloc(src.len())?;
        v.copy_from_host(src)?;
        Ok(v)
    }
    pub fn len(&self) -> usize { self.len }
    pub fn is_empty(&self) -> bool { self.len == 0 }
    pub fn as_ptr(&self) -> *const T { self.ptr as *const T }
    pub fn as_mut_ptr(&mut self) -> *mut T { self.ptr as *mut T }
    pub fn copy_from_host(&mut self, src: &[T]) -> Result<(), GpuError> {
        assert_eq!(src.len(), self.len);
        check(unsafe { ffi::b381_copy_to_device(self.ptr, src.as_ptr() as *const c_void, std::mem::size_of_val(src)) })
    }
    pub fn copy_to_host(&self, dst: &mut [T]) -> Result<(), GpuError> {
        assert_eq!(dst.len(), self.len);
        check(unsafe { ffi::b381_copy_to_host(dst.as_mut_ptr() as *mut c_void, self.ptr, std::mem::size_of_val(dst)) })
    }
}
impl<T: Copy> Drop for DeviceVec<T> {
    fn drop(&mut self) { if !self.ptr.is_null() { unsafe { ffi::b381_free(self.ptr) }; } }
}

/// Page-locked host buffer.
pub struct PinnedVec<T: Copy> { ptr: *mut T, len: usize }
unsafe impl<T: Copy> Send for PinnedVec<T> {}

impl<T: Copy> PinnedVec<T> {
    /// all-zero bytes: T is a plain array of u64 limbs everywhere in this crate
    pub fn zeroed(len: usize) -> Result<Self, GpuError> {
        let mut p = ptr::null_mut();
        check(unsafe { ffi::b381_host_alloc_pinned(&mut p, (len * std::mem::size_of::<T>()).max(8)) })?;
        unsafe { ptr::write_bytes(p as *mut u8, 0, len * std::mem::size_of::<T>()) };
        Ok(Self { ptr: p as *mut T, len })
    }
}
impl<T: Copy> PinnedVec<T> {
    pub fn as_slice(&self) -> &[T] { unsafe { std::slice::from_raw_parts(self.ptr, self.len) } }
    pub fn as_mut_slice(&mut self) -> &mut [T] { unsafe { std::slice::from_raw_parts_mut(self.ptr, self.len) } }
    pub fn as_mut_ptr(&mut self) -> *mut T { self.ptr }
}
impl<T: Copy> Drop for PinnedVec<T> {
    fn drop(&mut self) { unsafe { ffi::b381_host_free_pinned(self.ptr as *mut c_void) }; }
}
