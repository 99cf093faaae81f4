"""Loader of the CUDA library `fountain_b200/csrc/libfountain_gpu.so` (built in-tree by
`__graft_entry__.build()` / `make -C fountain_b200/csrc`).  Fails loudly: a missing extension,
a missing symbol or a missing GPU is an error, never a silent fallback.
"""
import ctypes as C
import os

from . import _abi as A

_HERE = os.path.dirname(os.path.abspath(__file__))
GPU_LIB_PATH = os.path.join(_HERE, "csrc", "libfountain_gpu.so")


def load_gpu_library(path=None):
    path = path or os.environ.get("FTN_GPU_LIB") or GPU_LIB_PATH   # FTN_GPU_LIB: A/B experiments with variant builds
    if not os.path.exists(path):
        raise ImportError("CUDA extension not built: %s is missing. Run `python -c 'import __graft_entry__ as g; "
                          "g.build()'` (nvcc, sm_100a). There is no CPU fallback." % path)
    return C.CDLL(path, mode=C.RTLD_GLOBAL)


def load_gpu_backend(path=None, require_device=True):
    from .api import Backend, FountainError
    lib = load_gpu_library(path)
    be = Backend(lib, "ftn_", list(A.PROTOTYPES.keys()), "cuda")
    if be.fn["abi_version"]() != A.FTN_ABI_VERSION:
        raise ImportError("libfountain_gpu.so ABI version mismatch")
    if require_device:
        n = C.c_int(0)
        rc = be.fn["device_count"](C.byref(n))
        if rc != A.FTN_OK or n.value < 1:
            msg = be.fn["last_error"]() or b""
            raise FountainError(A.FTN_ERR_NO_DEVICE, "no CUDA device visible (%s); the hot path has no CPU fallback"
                                % msg.decode("utf-8", "replace"))
    return be
