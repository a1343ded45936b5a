"""B200-native BLS12-381 proving backend (G1/G2 MSM, Fr NTT, Fr vecops) -- host layer.

The product is the CUDA library (csrc/, C ABI in include/b381.h); this package mirrors the
reference's Rust `core/` API (msm.rs, ntt.rs, vecops.rs, stream.rs, types.rs) over that ABI with
ctypes, because the image has no Rust toolchain.  There is no CPU fallback anywhere in here.
"""
from . import _lib  # noqa: F401
from .msm import BatchMsmHandle, G2MsmHandle, GpuMsmContext, MsmError, MsmHandle, PrecomputedBases  # noqa: F401
from .ntt import GpuNttContext, NttError, NttHandle, get_root_of_unity  # noqa: F401
from .stream import DeviceVec, GpuError, ManagedStream, ensure_backend_loaded, is_gpu_available, set_device  # noqa: F401
from .types import TypeConverter  # noqa: F401
from . import points, vecops  # noqa: F401
