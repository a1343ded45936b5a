"""Driver for tests/test_icicle_dispatch.py, run in a FRESH process: loads the mock ICICLE frontend
(oracle/_ref/libicicle_mock.so, built from the reference's own register_* / DeviceAPI declarations) with RTLD_GLOBAL,
then the three backend libraries, whose static initialisers now register "CUDA" with it; then invokes what they
registered.  Prints one JSON object.  modes: registration (no GPU needed) | compute (GPU)."""
import ctypes as C
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
MOCK = os.path.join(ROOT, "oracle", "_ref", "libicicle_mock.so")
LIBDIR = os.path.join(ROOT, "midnight_bls12_381_cuda_b200", "lib")
BACKENDS = ["libicicle_backend_cuda_device.so", "libicicle_backend_cuda_field_bls12_381.so",
            "libicicle_backend_cuda_curve_bls12_381.so"]


def load(field_dir=LIBDIR):
    mock = C.CDLL(MOCK, mode=C.RTLD_GLOBAL)
    libs = [C.CDLL(os.path.join(field_dir if "field" in b else LIBDIR, b), mode=C.RTLD_GLOBAL) for b in BACKENDS]
    return mock, libs


class VecOpsConfigV4(C.Structure):
    """upstream ICICLE v4.0.0 VecOpsConfig: the reference's struct (icicle_types.cuh:194-201) + batch_size, columns_batch"""
    _fields_ = [("stream", C.c_void_p), ("is_a_on_device", C.c_bool), ("is_b_on_device", C.c_bool),
                ("is_result_on_device", C.c_bool), ("is_async", C.c_bool), ("batch_size", C.c_int),
                ("columns_batch", C.c_bool), ("ext", C.c_void_p)]


def vecops_v4():
    """lib/upstream_v4/ field library: callbacks read the LONG VecOpsConfig.  The mock forwards `const VecOpsConfig&`
    (a pointer) untouched, so handing it the upstream struct exercises exactly what a real ICICLE v4 frontend passes,
    including core/vecops.rs:345-346's batch_size."""
    import numpy as np

    from midnight_bls12_381_cuda_b200 import _lib as L
    from oracle import cref as O
    mock, _ = load(os.path.join(LIBDIR, "upstream_v4"))
    res = {"mask": mock.mock_registered_mask()}
    assert C.sizeof(VecOpsConfigV4) == 32 and VecOpsConfigV4.ext.offset == 24 and VecOpsConfigV4.batch_size.offset == 12
    n, batch = 1 << 10, 5
    a = O.random_fr(51, n * batch)
    b = O.random_fr(52, n * batch)
    s = O.random_fr(53, batch)
    ptr = L.ptr
    ok = {}
    for columns in (False, True):
        cfg = VecOpsConfigV4(batch_size=batch, columns_batch=columns, ext=0xDEAD0000)     # ext must never be dereferenced
        for which, op in ((0, 0), (1, 1), (2, 2)):
            o = np.empty_like(b)
            rc = mock.mock_vecop(which, ptr(a), ptr(b), C.c_uint64(n), C.byref(cfg), ptr(o))
            ok[f"vec{which}_{int(columns)}"] = rc == 0 and bool((o == O.vecop(op, a, b)).all())
        for which, op in ((3, 2), (4, 0)):
            o = np.empty_like(b)
            rc = mock.mock_vecop(which, ptr(s), ptr(b), C.c_uint64(n), C.byref(cfg), ptr(o))
            exp = np.empty_like(b)
            for k in range(batch):
                if columns:
                    exp[k::batch] = O.vecop(op, s[k:k + 1], np.ascontiguousarray(b[k::batch]), a_scalar=True)
                else:
                    exp[k * n:(k + 1) * n] = O.vecop(op, s[k:k + 1], b[k * n:(k + 1) * n], a_scalar=True)
            ok[f"scalar{which}_{int(columns)}"] = rc == 0 and bool((o == exp).all())
    # batch_size = 0 / 1 behave like the single-vector call
    cfg = VecOpsConfigV4(batch_size=0)
    o = np.empty_like(b[:n])
    rc = mock.mock_vecop(3, ptr(s), ptr(b), C.c_uint64(n), C.byref(cfg), ptr(o))
    ok["scalar_single"] = rc == 0 and bool((o == O.vecop(2, s[:1], b[:n], a_scalar=True)).all())
    res["cases"] = ok
    return res


def g2_getters(curve_lib):
    """addresses of icicle::get_g2_msm_backend / get_g2_msm_precompute_bases_backend (src/backend/g2_registry.cu:84-101)"""
    out = subprocess.run(["nm", "-D", os.path.join(LIBDIR, BACKENDS[2])], capture_output=True, text=True, check=True).stdout
    names = {}
    for line in out.splitlines():
        parts = line.split()
        if len(parts) == 3 and "get_g2_msm" in parts[2]:
            names["pre" if "precompute" in parts[2] else "msm"] = parts[2]
    return [C.cast(getattr(curve_lib, names[k]), C.c_void_p) for k in ("msm", "pre")]


def registration():
    mock, libs = load()
    get_msm, get_pre = g2_getters(libs[2])
    return {"mask": mock.mock_registered_mask(), "g2": mock.mock_fetch_g2(get_msm, get_pre)}


def compute():
    import numpy as np

    from midnight_bls12_381_cuda_b200 import _lib as L
    from oracle import cref as O
    from oracle import pyref as P
    from vectors import fr_array
    mock, libs = load()
    get_msm, get_pre = g2_getters(libs[2])
    res = {"mask": mock.mock_registered_mask(), "g2": mock.mock_fetch_g2(get_msm, get_pre)}
    ptr = L.ptr
    lib = L.lib()

    # ---- G1 MSM + precompute through the registered callbacks (host buffers, Montgomery scalars and points)
    n = 1000
    bases = O.gen_series(1, [7, 0, 0, 0], [11, 0, 0, 0], n)
    sc_int = O.random_fr(42, n)
    sc = np.stack([O.unop("fr_to_mont", s, 4) for s in sc_int])
    exp = O.msm(1, sc_int, bases).tobytes()
    cfg = lib.b381_default_msm_config()
    cfg.are_scalars_montgomery_form = cfg.are_points_montgomery_form = True
    out = np.zeros(18, dtype=np.uint64)
    rc = mock.mock_msm(ptr(sc), ptr(bases), n, C.byref(cfg), ptr(out))
    res["g1_msm"] = rc == 0 and out.tobytes() == exp
    cfg.precompute_factor = 2
    table = np.zeros((2 * n, 12), dtype=np.uint64)
    rc = mock.mock_msm_precompute(ptr(bases), n, C.byref(cfg), ptr(table))
    out[:] = 0
    rc2 = mock.mock_msm(ptr(sc), ptr(table), n, C.byref(cfg), ptr(out))
    res["g1_precompute"] = rc == 0 and rc2 == 0 and out.tobytes() == exp
    # ---- G2 through the backend's own registry getters
    n2 = 300
    bases2 = O.gen_series(2, [3, 0, 0, 0], [5, 0, 0, 0], n2)
    exp2 = O.msm(2, sc_int[:n2], bases2).tobytes()
    cfg = lib.b381_default_msm_config()
    cfg.are_scalars_montgomery_form = cfg.are_points_montgomery_form = True
    out2 = np.zeros(36, dtype=np.uint64)
    rc = mock.mock_g2_msm(ptr(sc), ptr(bases2), n2, C.byref(cfg), ptr(out2))
    res["g2_msm"] = rc == 0 and out2.tobytes() == exp2
    cfg.precompute_factor = 2
    table2 = np.zeros((2 * n2, 24), dtype=np.uint64)
    rc = mock.mock_g2_msm_precompute(ptr(bases2), n2, C.byref(cfg), ptr(table2))
    out2[:] = 0
    rc2 = mock.mock_g2_msm(ptr(sc), ptr(table2), n2, C.byref(cfg), ptr(out2))
    res["g2_precompute"] = rc == 0 and rc2 == 0 and out2.tobytes() == exp2

    # ---- NTT domain + transform (root in STANDARD form, as upstream ICICLE passes it: core/ntt.rs:412-413)
    root = np.array(P.to_limbs(P.fr_omega(16), 4), dtype=np.uint64)
    res["ntt_init"] = mock.mock_ntt_init_domain(ptr(root), C.byref(L.NTTInitDomainConfig())) == 0
    a = O.random_fr(43, 1 << 12)
    y = np.empty_like(a)
    ncfg = lib.b381_default_ntt_config()
    res["ntt_forward"] = mock.mock_ntt(ptr(a), 1 << 12, 0, C.byref(ncfg), ptr(y)) == 0 and bool((y == O.ntt(a)).all())
    back = np.empty_like(a)
    res["ntt_inverse"] = mock.mock_ntt(ptr(y), 1 << 12, 1, C.byref(ncfg), ptr(back)) == 0 and bool((back == a).all())
    rou = np.zeros(4, dtype=np.uint64)
    res["ntt_rou"] = mock.mock_ntt_get_rou(C.c_uint64(12), ptr(rou)) == 0 and P.from_limbs(rou) == P.fr_to_mont(P.fr_omega(12))
    res["ntt_release"] = mock.mock_ntt_release_domain() == 0
    res["ntt_after_release"] = mock.mock_ntt(ptr(a), 1 << 12, 0, C.byref(ncfg), ptr(y))     # INVALID_ARGUMENT = 11

    # ---- the five vector ops
    b = O.random_fr(44, 1 << 12)
    vcfg = lib.b381_default_vecops_config()
    ok = True
    for which, op, scalar in ((0, 0, False), (1, 1, False), (2, 2, False), (3, 2, True), (4, 0, True)):
        o = np.empty_like(b)
        lhs = a[:1] if scalar else a
        rc = mock.mock_vecop(which, ptr(lhs), ptr(b), C.c_uint64(1 << 12), C.byref(vcfg), ptr(o))
        ok = ok and rc == 0 and bool((o == O.vecop(op, lhs, b, a_scalar=scalar)).all())
    res["vecops"] = ok

    # ---- DeviceAPI vtable
    cnt = C.c_int(0)
    res["dev_count"] = mock.mock_dev_count(C.byref(cnt)) == 0 and cnt.value >= 1
    res["dev_set"] = mock.mock_dev_set_device(0) == 0
    p, st = C.c_void_p(), C.c_void_p()
    host = np.arange(4096, dtype=np.uint64)
    back = np.zeros_like(host)
    ok = mock.mock_dev_malloc(C.byref(p), C.c_size_t(host.nbytes)) == 0
    ok = ok and mock.mock_dev_copy(p, ptr(host), C.c_size_t(host.nbytes), 0) == 0
    ok = ok and mock.mock_dev_copy(ptr(back), p, C.c_size_t(host.nbytes), 1) == 0 and bool((back == host).all())
    ok = ok and mock.mock_dev_memset(p, 0, C.c_size_t(host.nbytes)) == 0
    ok = ok and mock.mock_dev_create_stream(C.byref(st)) == 0
    ok = ok and mock.mock_dev_copy_async(ptr(back), p, C.c_size_t(host.nbytes), 1, st) == 0
    ok = ok and mock.mock_dev_synchronize(st) == 0 and not back.any()
    p2 = C.c_void_p()
    ok = ok and mock.mock_dev_malloc_async(C.byref(p2), C.c_size_t(1 << 20), st) == 0
    ok = ok and mock.mock_dev_memset_async(p2, 0xFF, C.c_size_t(1 << 20), st) == 0
    ok = ok and mock.mock_dev_free_async(p2, st) == 0 and mock.mock_dev_synchronize(st) == 0
    ok = ok and mock.mock_dev_destroy_stream(st) == 0 and mock.mock_dev_free(p) == 0
    total, free = C.c_size_t(0), C.c_size_t(0)
    ok = ok and mock.mock_dev_mem(C.byref(total), C.byref(free)) == 0 and 0 < free.value <= total.value
    res["dev_memory_and_streams"] = bool(ok)
    u, r, pin = C.c_int(-1), C.c_int(-1), C.c_int(-1)
    res["dev_properties"] = [mock.mock_dev_properties(C.byref(u), C.byref(r), C.byref(pin)), u.value, r.value, pin.value]
    return res


if __name__ == "__main__":
    print(json.dumps({"registration": registration, "compute": compute, "vecops_v4": vecops_v4}[sys.argv[1]]()))
