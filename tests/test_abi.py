"""CPU tier: the C-ABI library loads and exports every symbol include/fountain_gpu.h declares
(no compute calls without a GPU), struct layouts match the header, and the product refuses to
run without a device instead of falling back."""
import ctypes as C
import os
import re

import pytest

from fountain_b200 import _abi as A
from fountain_b200 import lib as gpulib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "fountain_gpu.h")


def _declared_symbols():
    src = open(HEADER).read()
    return sorted(set(re.findall(r"FTN_API\s+[\w\s\*]+?\b(ftn_\w+)\s*\(", src)))


def test_header_symbols_are_bound_in_python():
    assert _declared_symbols() == sorted("ftn_" + n for n in A.PROTOTYPES)


def test_library_exports_every_declared_symbol():
    if not os.path.exists(gpulib.GPU_LIB_PATH):
        pytest.fail("CUDA extension not built: run __graft_entry__.build()")
    lib = gpulib.load_gpu_library()
    for name in _declared_symbols():
        assert hasattr(lib, name), name
    lib.ftn_abi_version.restype = C.c_uint32
    assert lib.ftn_abi_version() == A.FTN_ABI_VERSION


def test_struct_sizes_match_the_header():
    # sizes implied by the header's field lists (LP64)
    assert C.sizeof(A.FtnRay) == 32 and C.sizeof(A.FtnHit) == 16 and C.sizeof(A.FtnPixel) == 16
    assert C.sizeof(A.FtnMeshDesc) == 16 + 4 + 12
    assert C.sizeof(A.FtnSphere) == 2 * 64 + 4 * 4 + 3 * 4 + 12
    assert C.sizeof(A.FtnCamera) == 2 * 64 + 16
    assert C.sizeof(A.FtnFilm) == 8 + 16 + 8
    assert C.sizeof(A.FtnIntegrator) == 12
    assert C.sizeof(A.FtnStats) == 6 * 8 + 2 * 8 + 4 * 4 + 5 * 24 + 16 + 8


def test_no_cpu_fallback_without_a_device():
    """Without a GPU the product raises; with one this test is vacuous."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from fountain_b200.api import FountainError
    with pytest.raises((FountainError, ImportError)):
        gpulib.load_gpu_backend()


def test_product_does_not_import_the_oracle():
    """The product never imports, links or loads anything under oracle/ or tests/."""
    pkg = os.path.join(ROOT, "fountain_b200")
    bad = re.compile(r"^\s*(import|from)\s+(oracle|tests)\b|libfountain_oracle|libfountain_hostsim|#include\s+\"[^\"]*(oracle|hostsim)", re.M)
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")) or f == "Makefile":
                src = open(os.path.join(dirpath, f)).read()
                assert not bad.search(src), os.path.join(dirpath, f)


def test_ctypes_structs_match_the_compiled_header(tmp_path):
    """sizeof / offsetof of every ABI struct as gcc lays the header out == the ctypes mirror in
    fountain_b200/_abi.py (a stale mirror would silently shift array strides)."""
    import subprocess
    structs = ["FtnRay", "FtnHit", "FtnMeshDesc", "FtnMaterial", "FtnTexture", "FtnSphere", "FtnLight", "FtnSceneDesc", "FtnCamera", "FtnFilm",
               "FtnSampler", "FtnIntegrator", "FtnPixel", "FtnStats"]
    lines = []
    for s in structs:
        lines.append('printf("%s %%zu", sizeof(%s));' % (s, s))
        for name, _ in getattr(A, s)._fields_:
            lines.append('printf(" %%zu", offsetof(%s, %s));' % (s, name))
        lines.append('printf("\\n");')
    src = tmp_path / "abi.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "fountain_gpu.h"\nint main(void) {\n' + "\n".join(lines) + "\nreturn 0; }\n")
    exe = tmp_path / "abi"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.strip().splitlines()
    for line in out:
        parts = line.split()
        cls = getattr(A, parts[0])
        assert C.sizeof(cls) == int(parts[1]), parts[0]
        for (name, _), off in zip(cls._fields_, parts[2:]):
            assert getattr(cls, name).offset == int(off), (parts[0], name)
