"""TEST INFRASTRUCTURE ONLY -- ctypes binding of the CPU oracle (oracle/libfountain_oracle.so).

Import this from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs only.  The product package (fountain_b200/) never imports it.
"""
import ctypes as C
import os
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(_HERE))
from fountain_b200 import _abi as A          # noqa: E402  (struct layouts only)
from fountain_b200.api import Backend        # noqa: E402

LIB_PATH = os.path.join(_HERE, "libfountain_oracle.so")


def build(force=False):
    """`make -C oracle`; a no-op when the library is already there and g++ is absent."""
    if os.path.exists(LIB_PATH) and not force:
        srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".h", ".cpp"))]
        srcs.append(os.path.join(os.path.dirname(_HERE), "include", "fountain_gpu.h"))
        if all(os.path.getmtime(s) <= os.path.getmtime(LIB_PATH) for s in srcs):
            return LIB_PATH
    subprocess.check_call(["make", "-C", _HERE] + (["-B"] if force else []), stdout=subprocess.DEVNULL)
    return LIB_PATH


_backend = None
_lib = None


def library():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB_PATH)
        f32, u32, i32, u64 = A.f32, A.u32, A.i32, A.u64
        P = C.POINTER
        protos = {
            "orc_set_threads": (C.c_int, [C.c_int]),
            "orc_hardware_threads": (C.c_int, []),
            "orc_intersect_count": (C.c_int, [C.c_void_p, C.c_size_t, P(A.FtnRay), P(A.FtnHit), P(u64)]),
            "orc_intersect_brute": (C.c_int, [C.c_void_p, C.c_size_t, P(A.FtnRay), P(A.FtnHit)]),
            "orc_film_to_rgb": (C.c_int, [C.c_size_t, P(A.FtnPixel), P(f32)]),
            "orc_debug_light_prims": (C.c_int, [C.c_void_p, P(i32), u32, P(u32)]),
            "orc_kat_morton3": (u32, [f32, f32, f32]),
            "orc_kat_expand_bits": (u32, [u32]),
            "orc_kat_to_fixed_point": (u32, [f32]),
            "orc_kat_sign_differs": (C.c_int, [f32, f32, f32]),
            "orc_kat_gamma": (f32, [C.c_int]),
            "orc_kat_next_float_up": (f32, [f32]),
            "orc_kat_next_float_down": (f32, [f32]),
            "orc_kat_bounds_intersect": (C.c_int, [P(f32), P(f32), P(A.FtnRay), P(f32)]),
            "orc_kat_fresnel_dielectric": (f32, [f32, f32, f32]),
            "orc_kat_fresnel_conductor": (None, [f32, P(f32), P(f32), P(f32)]),
            "orc_kat_distribution1d_sample": (None, [P(f32), C.c_int, f32, P(f32), P(f32), P(C.c_int)]),
            "orc_kat_concentric_sample_disk": (None, [f32, f32, P(f32)]),
            "orc_kat_offset_ray_origin": (None, [P(f32), P(f32), P(f32), P(f32), P(f32)]),
            "orc_kat_triangle_intersect": (C.c_int, [P(f32), P(f32), P(f32), P(A.FtnRay), P(f32)]),
            "orc_kat_sphere_intersect": (C.c_int, [P(A.FtnSphere), P(A.FtnRay), P(f32)]),
            "orc_kat_camera_ray": (None, [P(A.FtnCamera), f32, f32, f32, f32, f32, P(A.FtnRay)]),
            "orc_kat_bsdf": (None, [P(A.FtnMaterial), P(f32), P(f32), P(f32), P(f32)]),
            "orc_kat_env": (C.c_int, [C.c_void_p, P(f32), P(f32)]),
            "orc_kat_mipmap_lookup": (C.c_int, [P(f32), C.c_int, C.c_int, C.c_int, C.c_int, f32, f32, f32, P(f32)]),
            "orc_kat_counter_uniform": (f32, [u64, u64, u32]),
            "orc_kat_reference_stream": (None, [u64, C.c_int, P(f32)]),
        }
        for name, (res, args) in protos.items():
            fn = getattr(_lib, name)
            fn.restype, fn.argtypes = res, args
    return _lib


def backend():
    """A `Backend` the host mirror (fountain_b200.api) can run against -- the oracle arm of
    every parity test."""
    global _backend
    if _backend is None:
        _backend = Backend(library(), "orc_", list(A.ORACLE_SUBSET), "oracle")
    return _backend


def set_threads(n):
    library().orc_set_threads(int(n))


def hardware_threads():
    return int(library().orc_hardware_threads())


def light_prims(scene):
    """Primitive id behind every light of the oracle scene's light list (-1: not an area light)."""
    n = C.c_uint32(0)
    out = (A.i32 * 65536)()
    assert library().orc_debug_light_prims(scene.handle, out, 65536, C.byref(n)) == 0
    return list(out[: n.value])
