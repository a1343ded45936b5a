"""ctypes binding of libb381_cuda.so (the C ABI in include/b381.h).

There is no CPU fallback: if the CUDA library is missing or fails to load, importing any
compute entry point raises `BackendNotBuilt` loudly (north_star: "no CPU fallback").
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libb381_cuda.so")


class BackendNotBuilt(RuntimeError):
    pass


class B381Error(RuntimeError):
    def __init__(self, code: int, where: str):
        self.code = code
        super().__init__(f"{where}: {ERROR_NAMES.get(code, code)}")


ERROR_NAMES = {
    0: "SUCCESS", 1: "INVALID_DEVICE", 2: "OUT_OF_MEMORY", 3: "INVALID_POINTER", 4: "ALLOCATION_FAILED",
    5: "DEALLOCATION_FAILED", 6: "COPY_FAILED", 7: "SYNCHRONIZATION_FAILED", 8: "STREAM_CREATION_FAILED",
    9: "STREAM_DESTRUCTION_FAILED", 10: "API_NOT_IMPLEMENTED", 11: "INVALID_ARGUMENT",
    12: "BACKEND_LOAD_FAILED", 13: "LICENSE_CHECK_ERROR", 14: "UNKNOWN_ERROR",
}


class MSMConfig(C.Structure):
    """byte-compatible with icicle::MSMConfig (bls12-381/include/icicle_types.cuh:155-169)."""
    _fields_ = [
        ("stream", C.c_void_p), ("precompute_factor", C.c_int), ("c", C.c_int), ("bitsize", C.c_int),
        ("batch_size", C.c_int), ("are_points_shared_in_batch", C.c_bool), ("are_scalars_on_device", C.c_bool),
        ("are_scalars_montgomery_form", C.c_bool), ("are_points_on_device", C.c_bool),
        ("are_points_montgomery_form", C.c_bool), ("are_results_on_device", C.c_bool), ("is_async", C.c_bool),
        ("ext", C.c_void_p),
    ]


class Fr(C.Structure):
    _fields_ = [("l", C.c_uint64 * 4)]


class NTTConfig(C.Structure):
    """icicle::NTTConfig<Fr> (icicle_types.cuh:102-113)."""
    _fields_ = [
        ("stream", C.c_void_p), ("coset_gen", Fr), ("batch_size", C.c_int), ("columns_batch", C.c_bool),
        ("ordering", C.c_int), ("are_inputs_on_device", C.c_bool), ("are_outputs_on_device", C.c_bool),
        ("is_async", C.c_bool), ("ext", C.c_void_p),
    ]


class NTTInitDomainConfig(C.Structure):
    _fields_ = [("stream", C.c_void_p), ("is_async", C.c_bool), ("ext", C.c_void_p)]


class VecOpsConfig(C.Structure):
    _fields_ = [("stream", C.c_void_p), ("is_a_on_device", C.c_bool), ("is_b_on_device", C.c_bool),
                ("is_result_on_device", C.c_bool), ("is_async", C.c_bool), ("ext", C.c_void_p)]


_lib = None

# every symbol include/b381.h declares (tests/test_abi.py checks the .so exports each one)
EXPORTS = [
    "b381_default_msm_config", "b381_default_ntt_config", "b381_default_vecops_config",
    "b381_g1_msm", "b381_g2_msm", "b381_g1_msm_precompute_bases", "b381_g2_msm_precompute_bases",
    "b381_ntt_init_domain", "b381_ntt_release_domain", "b381_ntt", "b381_ntt_get_rou_from_domain",
    "b381_vector_add", "b381_vector_sub", "b381_vector_mul", "b381_scalar_mul_vec", "b381_scalar_add_vec", "b381_scalar_mul_vec_batch", "b381_scalar_add_vec_batch",
    "bls12_381_g1_msm_cuda", "bls12_381_g2_msm_cuda", "bls12_381_ntt_cuda", "bls12_381_ntt_init_domain_cuda",
    "bls12_381_ntt_release_domain_cuda", "bls12_381_coset_ntt_cuda", "bls12_381_field_ntt_cuda",
    "bls12_381_field_ntt_init_domain_cuda", "bls12_381_field_ntt_release_domain_cuda",
    "bls12_381_vector_add", "bls12_381_vector_sub", "bls12_381_vector_mul",
    "bls12_381_g1_affine_to_projective", "bls12_381_g1_projective_to_affine", "bls12_381_g2_affine_to_projective",
    "bls12_381_g2_projective_to_affine", "b381_g1_is_on_curve", "b381_g2_is_on_curve",
    "b381_g1_is_in_subgroup", "b381_g2_is_in_subgroup", "bls12_381_g1_scalar_mul_glv", "bls12_381_g1_scalar_mul",
    "b381_vector_sum", "b381_vector_inv", "b381_bit_reverse", "b381_montgomery_convert",
    "vec_add_cuda", "vec_sub_cuda", "vec_mul_cuda", "scalar_mul_vec_cuda", "scalar_add_vec_cuda", "vec_sum_cuda",
    "b381_device_count", "b381_set_device", "b381_malloc", "b381_malloc_async", "b381_free", "b381_free_async",
    "b381_memset", "b381_copy_to_device", "b381_copy_to_host", "b381_copy_to_device_async",
    "b381_copy_to_host_async", "b381_copy_device_to_device", "b381_host_alloc_pinned", "b381_host_free_pinned",
    "b381_stream_create", "b381_stream_destroy", "b381_stream_synchronize", "b381_device_synchronize",
    "b381_ntt_dist_columns", "b381_ntt_dist_columns_p2p", "b381_ipc_alloc", "b381_ipc_open", "b381_ipc_close", "b381_g1_msm_partial", "b381_g2_msm_partial", "b381_g1_msm_combine", "b381_g2_msm_combine",
    "b381_g1_point_series", "b381_g2_point_series",
    "b381_bench_imad_peak", "b381_bench_field_mul", "b381_msm_last_timings", "b381_msm_last_info", "b381_msm_last_level0_ms", "b381_ntt_last_info", "b381_version",
]


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise BackendNotBuilt(
                f"{LIB_PATH} not found: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback)")
        try:
            _lib = C.CDLL(LIB_PATH)
        except OSError as e:  # pragma: no cover
            raise BackendNotBuilt(f"cannot load {LIB_PATH}: {e}") from e
        _lib.b381_version.restype = C.c_char_p
        _lib.b381_default_msm_config.restype = MSMConfig
        if hasattr(_lib, "b381_default_ntt_config"):
            _lib.b381_default_ntt_config.restype = NTTConfig
        _lib.b381_default_vecops_config.restype = VecOpsConfig
    return _lib


def check(code: int, where: str) -> None:
    if code != 0:
        raise B381Error(code, where)


def ptr(x) -> C.c_void_p:
    """address of a torch tensor / numpy array / bytes-like / int."""
    if x is None:
        return C.c_void_p(0)
    if isinstance(x, int):
        return C.c_void_p(x)
    if hasattr(x, "data_ptr"):
        return C.c_void_p(x.data_ptr())
    if hasattr(x, "ctypes"):
        return C.c_void_p(x.ctypes.data)
    return C.cast(x, C.c_void_p)
