"""The reference's integration tests (tests/furnace.rs) run through the CPU oracle, in both
sampler modes -- this pins the oracle's whole integrator stack (sphere + EFloat, area light,
MIS, Lambert, film) to the reference's own acceptance thresholds.  CPU only.
"""
import numpy as np
import pytest

from fountain_b200 import _abi as A
from fountain_b200 import api, scenes


def _render(backend, integrator, mode, spp=128, seed=0):
    scene, camera, film = scenes.furnace_scene(backend=backend)
    sampler = api.RandomSampler.new_with_seed(spp, seed, mode=mode)     # pbrt.rs:475: seed 0
    stats = api.SamplerIntegrator(camera, integrator).render_parallel(scene, film, sampler)
    rgb, (w, h) = film.into_spectrum_buffer()
    assert (w, h) == (16, 16)
    return rgb, stats


MODES = [A.FTN_SAMPLER_REFERENCE_TILE_STREAM, A.FTN_SAMPLER_COUNTER]


# tests/furnace.rs:11-25
@pytest.mark.parametrize("mode", MODES)
def test_furnace_path(orc_backend, mode):
    rgb, _ = _render(orc_backend, api.PathIntegrator(10, 1.0), mode)
    assert np.all(np.abs(rgb - 2.0) <= 0.1)


# tests/furnace.rs:28-41
@pytest.mark.parametrize("mode", MODES)
def test_furnace_path_no_rr(orc_backend, mode):
    rgb, stats = _render(orc_backend, api.PathIntegrator(10, 0.0), mode)
    assert np.all(np.abs(rgb - 2.0) <= 0.001)
    # 1 + sum_{k=1..10} .5^k exactly
    assert np.all(np.abs(rgb - 1.9990234) <= 2e-4)
    assert stats["camera_samples"] == 16 * 16 * 128
    # primary + per bounce (1 MIS closest + 1 continuation closest), 1 shadow any-hit
    assert stats["rays_closest"] == 16 * 16 * 128 * (1 + 2 * 10)
    assert stats["rays_any"] == 16 * 16 * 128 * 10


# tests/furnace.rs:44-60
@pytest.mark.parametrize("mode", MODES)
def test_furnace_directlighting(orc_backend, mode):
    rgb, _ = _render(orc_backend, api.DirectLightingIntegrator(3), mode)
    assert np.all(np.abs(rgb - 1.5) <= 0.00001)


def test_sample_sharding_partitions_the_render(orc_backend):
    """Multi-GPU contract: the sum of the per-rank partial films over sample shards equals the
    single render (same samples, same per-sample values; only the f32 summation order differs)."""
    scene, camera, film = scenes.furnace_scene(backend=orc_backend)
    integ = api.SamplerIntegrator(camera, api.PathIntegrator(5, 1.0))
    sampler = api.RandomSampler.new_with_seed(16, 5)
    integ.render_parallel(scene, film, sampler)
    full = film.pixels.copy()
    acc = np.zeros_like(full)
    for rank in range(4):
        integ.render_parallel(scene, film, sampler, sample_begin=rank, sample_stride=4)
        assert np.all(film.pixels[..., 3] == 4.0)
        acc += film.pixels
    assert np.allclose(acc, full, rtol=1e-5, atol=1e-6)


def test_rounded_cube_render_sane(orc_backend):
    """C2 at reduced size: closed white-ish cube under a uniform unit environment -- every pixel
    is in (0, 1], background is exactly the env radiance, mean energy is conserved."""
    scene, camera, film = scenes.rounded_cube_scene(backend=orc_backend, resolution=(48, 48))
    stats = api.SamplerIntegrator(camera, api.PathIntegrator(5, 1.0)).render_parallel(
        scene, film, api.RandomSampler.new_with_seed(8, 0))
    rgb, _ = film.into_spectrum_buffer()
    img = rgb.reshape(48, 48, 3)
    assert np.all(np.isfinite(img)) and img.min() >= 0.0 and img.max() <= 1.0 + 1e-4
    assert np.allclose(img[0, 0], 1.0, atol=1e-5)           # corner pixel sees only the environment
    assert 0.2 < img[24, 24].mean() < 1.0                    # the cube is darker than the sky
    assert stats["rays_any"] > 0 and stats["rays_closest"] > stats["camera_samples"]
