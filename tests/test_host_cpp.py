"""The C++ host side (include/fountain_host.hpp) above the C ABI: tests/cpp/host_tests.cpp holds
the reference's integration tests (tests/furnace.rs, tests/tri_watertight.rs) restated in C++.
CPU tier: the host code drives the checker libraries (host logic only).  GPU tier: the same
binary drives libfountain_gpu.so."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CPP = os.path.join(ROOT, "tests", "cpp")
PLY = os.path.join(ROOT, "data", "rounded_cube.ply")


@pytest.fixture(scope="module")
def host_tests():
    subprocess.run(["make", "-s", "-C", CPP], check=True)
    return os.path.join(CPP, "host_tests")


def run(binary, lib, prefix, *names):
    p = subprocess.run([binary, lib, prefix, PLY, *names], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout + p.stderr
    return p.stdout


def test_host_cpp_against_oracle(host_tests, oracle):
    out = run(host_tests, oracle.build(), "orc_")
    assert "11 run, 0 failed" in out


def test_host_cpp_against_hostsim(host_tests):
    from tests.hostsim import sim
    lib = sim.build()
    out = run(host_tests, lib, "sim_", "furnace_test_path_no_rr", "furnace_test_directlighting", "test_rounded_cube",
              "world_bound_and_morton_order", "image_texture_closed_forms", "texture_table_parameters")
    assert "6 run, 0 failed" in out


def test_host_cpp_fails_loudly_without_a_library(host_tests):
    p = subprocess.run([host_tests, "/nonexistent/libfountain_gpu.so", "ftn_", PLY], capture_output=True, text=True)
    assert p.returncode == 3 and "cannot load" in p.stderr


@pytest.mark.gpu
def test_host_cpp_on_gpu(host_tests, gpu_backend):
    from fountain_b200 import lib as gpulib
    out = run(host_tests, os.environ.get("FTN_GPU_LIB") or gpulib.GPU_LIB_PATH, "ftn_")
    assert "11 run, 0 failed" in out
