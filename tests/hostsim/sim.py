"""TEST HARNESS ONLY -- binding of tests/hostsim/libfountain_hostsim.so, the FTN_HD device
functions of fountain_b200/csrc compiled for the CPU (see hostsim.cpp).  Lets the CPU-only test
tier check the kernels' per-thread logic against the oracle.  Never imported by the product."""
import ctypes as C
import os
import subprocess

from fountain_b200 import _abi as A
from fountain_b200.api import Backend

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libfountain_hostsim.so")
_lib = None
_backend = None


def build():
    csrc = os.path.join(_HERE, "..", "..", "fountain_b200", "csrc")
    srcs = [os.path.join(_HERE, "hostsim.cpp")] + [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith((".cuh", ".h"))]
    if os.path.exists(LIB_PATH) and all(os.path.getmtime(s) <= os.path.getmtime(LIB_PATH) for s in srcs):
        return LIB_PATH
    subprocess.check_call(["make", "-C", _HERE], stdout=subprocess.DEVNULL)
    return LIB_PATH


def library():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB_PATH)
        f32, u32, u64, P = A.f32, A.u32, A.u64, C.POINTER
        protos = {
            "sim_intersect_count": (C.c_int, [C.c_void_p, C.c_size_t, P(A.FtnRay), P(A.FtnHit), P(u64)]),
            "sim_kat_triangle_intersect": (C.c_int, [P(f32), P(f32), P(f32), P(A.FtnRay), P(f32)]),
            "sim_kat_sphere_intersect": (C.c_int, [P(A.FtnSphere), P(A.FtnRay), P(f32)]),
            "sim_kat_camera_ray": (None, [P(A.FtnCamera), f32, f32, f32, f32, f32, P(A.FtnRay)]),
            "sim_kat_offset_ray_origin": (None, [P(f32), P(f32), P(f32), P(f32), P(f32)]),
            "sim_kat_bsdf": (None, [P(A.FtnMaterial), P(f32), P(f32), P(f32), P(f32)]),
            "sim_kat_mipmap_lookup": (C.c_int, [P(f32), C.c_int, C.c_int, C.c_int, C.c_int, f32, f32, f32, P(f32)]),
            "sim_kat_env": (C.c_int, [C.c_void_p, P(f32), P(f32)]),
            "sim_kat_counter_uniform": (f32, [u64, u64, u32]),
            "sim_kat_gamma": (f32, [C.c_int]),
            "sim_kat_slab_test": (C.c_int, [P(f32), P(f32), P(A.FtnRay), P(f32), C.c_int]),
        }
        for name, (res, args) in protos.items():
            fn = getattr(_lib, name)
            fn.restype, fn.argtypes = res, args
    return _lib


def backend():
    global _backend
    if _backend is None:
        _backend = Backend(library(), "sim_", list(A.ORACLE_SUBSET), "hostsim")
    return _backend
