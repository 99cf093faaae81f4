"""The reference's integration tests (tests/furnace.rs) run through the CPU oracle, in both
sampler modes -- this pins the oracle's whole integrator stack (sphere + EFloat, area light,
MIS, Lambert, film) to the reference's own acceptance thresholds.  CPU only.
"""
import numpy as np
import pytest

from fountain_b200 import _abi as A
from fountain_b200 import api
from workloads import scenes


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


# ---- delta lights (light/point.rs, light/distant.rs) --------------------------------
# The reference holds no test for them, so the restatement is pinned against the closed form of
# direct lighting on a Lambertian floor: L = Kd/pi * I cos(theta) / d^2 (point), Kd/pi * L cos(theta) (distant).
def _floor_radiance(scene_kw, integrator, orc_backend, spp):
    from workloads import scenes
    scene, camera, film = scenes.delta_lights_scene(backend=orc_backend, resolution=(9, 9), fov=2.0, occluder=False, **scene_kw)
    # look straight down at the origin from (0, 0, 20): every pixel sees the floor within 0.4 units of the origin
    from fountain_b200.transform import Transform
    camera = api.PerspectiveCamera(Transform.look_at((0, 0, 20), (0, 0, 0), (0, 1, 0)).inverse(), (9, 9), fov=2.0)
    api.SamplerIntegrator(camera, integrator).render_parallel(scene, film, api.RandomSampler.new_with_seed(spp, 1))
    rgb, _ = film.into_spectrum_buffer()
    return rgb.reshape(9, 9, 3)


def test_point_light_closed_form(orc_backend):
    rgb = _floor_radiance(dict(lights=("point",)), api.DirectLightingIntegrator(3), orc_backend, 4)
    kd, inten, h = np.array([0.6, 0.5, 0.4]), np.array([30.0, 28.0, 26.0]), 4.0
    expected = kd / np.pi * inten / h ** 2          # at the origin: d = 4, cos(theta) = 1
    assert np.allclose(rgb[4, 4], expected, rtol=2e-3)
    assert np.allclose(rgb, expected[None, None, :], rtol=0.03)          # cos/d^2 fall-off over +-0.35 units


def test_distant_light_closed_form(orc_backend):
    rgb = _floor_radiance(dict(lights=("distant",)), api.PathIntegrator(1, 1.0), orc_backend, 4)
    kd, rad = np.array([0.6, 0.5, 0.4]), np.array([1.5, 1.6, 1.7])
    cos_t = 2.0 / np.sqrt(5.0)                      # direction towards the light = normalize(1, 0, 2)
    assert np.allclose(rgb, (kd / np.pi * rad * cos_t)[None, None, :], rtol=1e-4)


def test_two_delta_lights_one_is_picked_per_sample(orc_backend):
    """uniform_sample_one_light (integrator/mod.rs:289-305): one of the two lights, weighted by 2."""
    rgb = _floor_radiance(dict(lights=("point", "distant")), api.DirectLightingIntegrator(3), orc_backend, 2048)
    kd = np.array([0.6, 0.5, 0.4])
    expected = kd / np.pi * (np.array([30.0, 28.0, 26.0]) / 16.0 + np.array([1.5, 1.6, 1.7]) * 2.0 / np.sqrt(5.0))
    assert np.allclose(rgb[4, 4], expected, rtol=0.05)


def test_delta_light_shadow(orc_backend):
    """The occluder at z = 1 over x in [1, 3] shadows the floor from the point light at (0, 0, 4):
    the shadow covers x in [4/3, 4] at y = 0; floor points there are exactly black."""
    from workloads import scenes
    from fountain_b200.transform import Transform
    scene, _, film = scenes.delta_lights_scene(backend=orc_backend, resolution=(5, 5), lights=("point",))
    camera = api.PerspectiveCamera(Transform.look_at((4.5, 0, 20), (4.5, 0, 0), (0, 1, 0)).inverse(), (5, 5), fov=1.0)   # x ~ 4.5: lit
    api.SamplerIntegrator(camera, api.DirectLightingIntegrator(3)).render_parallel(scene, film, api.RandomSampler.new_with_seed(4, 0))
    lit = film.into_spectrum_buffer()[0]
    camera = api.PerspectiveCamera(Transform.look_at((3.5, 0, 20), (3.5, 0, 0), (0, 1, 0)).inverse(), (5, 5), fov=1.0)   # x ~ 3.5: in shadow
    api.SamplerIntegrator(camera, api.DirectLightingIntegrator(3)).render_parallel(scene, film, api.RandomSampler.new_with_seed(4, 0))
    dark = film.into_spectrum_buffer()[0]
    assert np.all(lit > 0.05) and np.all(dark == 0.0)


# ---- mirror (material/mirror.rs + SpecularReflection<FresnelNoOp>, reflection/mod.rs:165-197) ----------------
# No reference test covers it: pinned against the closed form -- a mirror under a uniform environment
# of radiance 1 returns exactly Kr (one specular bounce, pdf 1, f = Kr / |cos|, times |cos|).
def test_mirror_closed_form_path(orc_backend):
    from workloads import scenes
    scene, camera, film = scenes.mirror_scene(backend=orc_backend, resolution=(9, 9), with_floor=False)
    from fountain_b200.transform import Transform
    camera = api.PerspectiveCamera(Transform.look_at((0, -7, 1.5), (0, 2, 1.0), (0, 0, 1)).inverse(), (9, 9), fov=5.0)
    api.SamplerIntegrator(camera, api.PathIntegrator(5, 1.0)).render_parallel(scene, film, api.RandomSampler.new_with_seed(4, 0))
    rgb = film.into_spectrum_buffer()[0]
    assert np.allclose(rgb, np.array([0.8, 0.7, 0.6])[None, :], rtol=1e-5)
    # depth 0: the camera ray hits the mirror, nothing is added (no emission, no non-specular lobe)
    api.SamplerIntegrator(camera, api.PathIntegrator(0, 1.0)).render_parallel(scene, film, api.RandomSampler.new_with_seed(4, 0))
    assert np.all(film.into_spectrum_buffer()[0] == 0.0)


def test_mirror_closed_form_direct_lighting(orc_backend):
    """specular_reflect (integrator/mod.rs:40-103): with max_depth 2 the mirror shows Kr * environment,
    with max_depth 1 the recursion is cut (depth + 1 < max_depth fails) and the mirror is black."""
    from workloads import scenes
    from fountain_b200.transform import Transform
    scene, _, film = scenes.mirror_scene(backend=orc_backend, resolution=(9, 9), with_floor=False)
    camera = api.PerspectiveCamera(Transform.look_at((0, -7, 1.5), (0, 2, 1.0), (0, 0, 1)).inverse(), (9, 9), fov=5.0)
    api.SamplerIntegrator(camera, api.DirectLightingIntegrator(2)).render_parallel(scene, film, api.RandomSampler.new_with_seed(4, 0))
    assert np.allclose(film.into_spectrum_buffer()[0], np.array([0.8, 0.7, 0.6])[None, :], rtol=1e-5)
    api.SamplerIntegrator(camera, api.DirectLightingIntegrator(1)).render_parallel(scene, film, api.RandomSampler.new_with_seed(4, 0))
    assert np.all(film.into_spectrum_buffer()[0] == 0.0)


@pytest.mark.parametrize("mode", [A.FTN_SAMPLER_COUNTER, A.FTN_SAMPLER_REFERENCE_TILE_STREAM])
def test_mirror_scene_renders_in_both_sampler_modes(orc_backend, mode):
    from workloads import scenes
    scene, camera, film = scenes.mirror_scene(backend=orc_backend, resolution=(32, 32))
    api.SamplerIntegrator(camera, api.PathIntegrator(5, 1.0)).render_parallel(scene, film, api.RandomSampler.new_with_seed(16, 0, mode=mode))
    rgb = film.into_spectrum_buffer()[0]
    assert np.isfinite(rgb).all() and rgb.mean() > 0.1


# ---- textured Kd: Checkerboard2DTexture (AAMethod::None), UVTexture, UVMapping -------------------------------
# No reference test: pinned by closed form -- under a distant light from straight above a Lambertian floor
# shows Kd(texel) / pi * L exactly; the floor's uv are its world (x, y).
def _probe(orc_backend, xy, texture):
    from workloads import scenes
    scene, camera, film = scenes.textured_floor_scene(backend=orc_backend, resolution=(5, 5), texture=texture, look_at=xy, fov=0.5)
    api.SamplerIntegrator(camera, api.DirectLightingIntegrator(1)).render_parallel(scene, film, api.RandomSampler.new_with_seed(2, 0))
    return film.into_spectrum_buffer()[0]


def test_checkerboard_closed_form(orc_backend):
    t1, t2 = np.array([0.8, 0.2, 0.2]) / np.pi * 3.0, np.array([0.1, 0.1, 0.9]) / np.pi * 3.0
    # (floor(s) + floor(t)) % 2 == 0 with Rust's truncating %: (-1 + 0) % 2 = -1 -> tex2, (-1 + -1) % 2 = 0 -> tex1
    for xy, expected in (((0.5, 0.5), t1), ((1.5, 0.5), t2), ((1.5, 1.5), t1), ((-0.5, 0.5), t2), ((-0.5, -0.5), t1), ((-1.5, 0.5), t1), ((2.5, -3.5), t1)):
        assert np.allclose(_probe(orc_backend, xy, "checkerboard"), expected[None, :], rtol=1e-5), xy
    # UVMapping(2, .5, .25, -.5): s = 2x + .25, t = .5y - .5
    for xy in ((0.1, 0.5), (0.6, 0.5), (0.1, 3.5), (0.1, 2.5), (-0.3, 2.5), (-0.9, -2.5)):
        s, t = 2 * xy[0] + 0.25, 0.5 * xy[1] - 0.5
        want = t1 if int(np.fmod(np.floor(s) + np.floor(t), 2)) == 0 else t2     # fmod truncates like Rust's %
        assert np.allclose(_probe(orc_backend, xy, "checkerboard_scaled"), want[None, :], rtol=1e-5), xy


def test_uv_texture_closed_form(orc_backend):
    # the probe's 5x5 pixels cover +-0.13 units around the point: compare the mean (the texture is linear there)
    rgb = _probe(orc_backend, (1.3, 2.6), "uv")            # s = .5 * 1.3 + .1 = .75, t = .25 * 2.6 + .2 = .85
    assert np.allclose(rgb.mean(axis=0), np.array([0.75, 0.85, 0.0]) / np.pi * 3.0, rtol=5e-3, atol=1e-6)
    rgb = _probe(orc_backend, (-1.3, -2.6), "uv")          # s = -.55 -> .45, t = -.45 -> .55
    assert np.allclose(rgb.mean(axis=0), np.array([0.45, 0.55, 0.0]) / np.pi * 3.0, rtol=5e-3, atol=1e-6)
    assert np.all(rgb[:, 2] == 0.0)


# ---- Oren-Nayar (matte with sigma != 0: matte.rs:42-49, reflection/mod.rs:252-296) -------------------------
# No reference test: pinned by closed form under a distant light: L_o = r/pi (a + b max(0, cos dphi) sin(alpha) tan(beta)) L cos(theta_i).
def _oren_nayar_probe(backend, sigma, theta_i, theta_o, dphi):
    from fountain_b200.transform import Transform
    v = np.array([[-6, -6, 0], [6, -6, 0], [6, 6, 0], [-6, 6, 0]], np.float32)
    mesh = api.TriangleMesh(Transform.identity(), np.array([0, 1, 2, 0, 2, 3], np.uint32), v)
    ti, to = np.radians(theta_i), np.radians(theta_o)
    light = api.DistantLight.from_params(L=2.0, from_=(np.sin(ti), 0.0, np.cos(ti)), to=(0.0, 0.0, 0.0))
    scene = api.Scene([api.GeometricPrimitive(mesh, api.MatteMaterial((0.7, 0.5, 0.3), sigma=sigma))], [light], backend=backend)
    eye = 30.0 * np.array([np.sin(to) * np.cos(np.radians(dphi)), np.sin(to) * np.sin(np.radians(dphi)), np.cos(to)])
    camera = api.PerspectiveCamera(Transform.look_at(tuple(eye), (0, 0, 0), (0, 1, 0) if theta_o == 0 else (0, 0, 1)).inverse(), (5, 5), fov=0.3)
    film = api.Film((5, 5), backend=backend)
    api.SamplerIntegrator(camera, api.DirectLightingIntegrator(1)).render_parallel(scene, film, api.RandomSampler.new_with_seed(2, 0))
    return film.into_spectrum_buffer()[0]


def _oren_nayar_expected(sigma, theta_i, theta_o, dphi):
    s = np.radians(min(max(sigma, 0.0), 90.0)); s2 = s * s
    a, b = 1.0 - s2 / (2.0 * (s2 + 0.33)), 0.45 * s2 / (s2 + 0.09)
    ti, to = np.radians(theta_i), np.radians(theta_o)
    max_cos = max(0.0, np.cos(np.radians(dphi))) if (np.sin(ti) > 1e-4 and np.sin(to) > 1e-4) else 0.0
    if abs(np.cos(ti)) > abs(np.cos(to)):
        sin_alpha, tan_beta = np.sin(to), np.sin(ti) / abs(np.cos(ti))
    else:
        sin_alpha, tan_beta = np.sin(ti), np.sin(to) / abs(np.cos(to))
    return np.array([0.7, 0.5, 0.3]) / np.pi * (a + b * max_cos * sin_alpha * tan_beta) * 2.0 * np.cos(ti)


@pytest.mark.parametrize("sigma,theta_i,theta_o,dphi", [(20.0, 0.0, 0.0, 0.0), (20.0, 50.0, 30.0, 0.0), (35.0, 30.0, 55.0, 0.0),
                                                       (20.0, 50.0, 30.0, 180.0), (60.0, 40.0, 40.0, 60.0), (120.0, 25.0, 45.0, 0.0)])
def test_oren_nayar_closed_form(orc_backend, sigma, theta_i, theta_o, dphi):
    rgb = _oren_nayar_probe(orc_backend, sigma, theta_i, theta_o, dphi)
    assert np.allclose(rgb, _oren_nayar_expected(sigma, theta_i, theta_o, dphi)[None, :], rtol=3e-3)


def test_oren_nayar_sigma_zero_is_lambert(orc_backend):
    a = _oren_nayar_probe(orc_backend, 0.0, 40.0, 20.0, 30.0)
    assert np.allclose(a, (np.array([0.7, 0.5, 0.3]) / np.pi * 2.0 * np.cos(np.radians(40.0)))[None, :], rtol=1e-5)


# ---- image texture (texture/image.rs, mipmap.rs:245-279, interaction.rs:124-176, camera/mod.rs:145-205) --------
# No reference test renders one.  Pinned by closed forms: a pyramid whose level l is the constant c_l returns
# lerp(level - floor(level), c_floor, c_floor+1) whatever st is, and for a camera looking straight down a plane
# from height H the texture-space footprint is uscale * (pixel size on the plane) / sqrt(spp) (integrator/mod.rs:249).
def constant_level_mipmap(n=64, wrap="repeat"):
    levels, l = [], 0
    while True:
        m = max(1, n >> l)
        levels.append(np.full((m, m, 3), 0.1 + 0.1 * l, np.float32))
        if m == 1:
            break
        l += 1
    return api.MIPMap(levels=levels, wrap=wrap)


def expected_mip_level(n_levels, uscale, fov_deg, res, height, spp):
    pix = 2.0 * height * np.tan(np.radians(fov_deg) / 2.0) / res
    width = 2.0 * uscale * pix / np.sqrt(spp)
    return n_levels - 1 + np.log2(max(width, 1e-8))


@pytest.mark.parametrize("uscale", [0.02, 1.0, 3.0, 11.0, 500.0])
def test_image_texture_level_selection_closed_form(orc_backend, uscale):
    mp = constant_level_mipmap()
    tex = api.ImageTexture(mp, api.UVMapping(uscale, uscale, 0.3, 0.1))
    scene, camera, film = scenes.textured_floor_scene(backend=orc_backend, resolution=(5, 5), texture=tex, look_at=(0.7, -0.4), fov=0.5)
    api.SamplerIntegrator(camera, api.DirectLightingIntegrator(1)).render_parallel(scene, film, api.RandomSampler.new_with_seed(4, 0))
    rgb = film.into_spectrum_buffer()[0]
    level = expected_mip_level(len(mp.levels), uscale, 0.5, 5, 30.0, 4)
    n = len(mp.levels)
    if level < 0: c = 0.1
    elif level >= n - 1: c = 0.1 + 0.1 * (n - 1)
    else: c = 0.1 + 0.1 * level          # the lerp of c_l = .1 + .1 l is linear in the level
    assert np.allclose(rgb, c / np.pi * 3.0, rtol=2e-3), (level, rgb[:2])


def test_image_texture_bilinear_closed_form(orc_backend):
    """Level 0 of a ramp image is reproduced exactly by the bilinear `triangle` filter in the interior."""
    n = 16
    ramp = np.zeros((n, n, 3), np.float32)
    ramp[..., 0] = (np.arange(n)[None, :] + 0.5) / n          # = s at texel centres
    ramp[..., 1] = (np.arange(n)[:, None] + 0.5) / n          # = t
    ramp[..., 2] = 0.25
    tex = api.ImageTexture(api.MIPMap(ramp, "clamp"), api.UVMapping(1.0 / 12.0, 1.0 / 12.0, 0.5, 0.5))   # the 12x12 floor -> [0, 1]^2
    for xy in ((0.0, 0.0), (2.2, -3.1), (-4.0, 1.5)):
        scene, camera, film = scenes.textured_floor_scene(backend=orc_backend, resolution=(5, 5), texture=tex, look_at=xy, fov=0.5)
        api.SamplerIntegrator(camera, api.DirectLightingIntegrator(1)).render_parallel(scene, film, api.RandomSampler.new_with_seed(2, 0))
        rgb = film.into_spectrum_buffer()[0]
        want = np.array([xy[0] / 12 + 0.5, xy[1] / 12 + 0.5, 0.25]) / np.pi * 3.0
        assert np.allclose(rgb.mean(axis=0), want, rtol=3e-3), (xy, rgb.mean(axis=0), want)


def test_image_texture_bad_pyramid_is_rejected(orc_backend):
    mp = api.MIPMap(scenes.procedural_image(), "repeat")
    mp.levels = mp.levels[:-1]          # one level short of 1 + floor(log2(max(w, h)))
    with pytest.raises(api.FountainError):
        scenes.textured_floor_scene(backend=orc_backend, texture=api.ImageTexture(mp))


# ---- textured Kr (mirror.rs:23 evaluates its Kr texture like matte.rs:37 its Kd) -----------------------------------
def mirror_kr_closed_form(backend):
    t1, t2 = np.array([0.8, 0.7, 0.6]), np.array([0.3, 0.2, 0.1])
    for xz, want in (((0.5, 0.5), t1), ((1.5, 0.5), t2), ((-0.5, 0.5), t2), ((-0.5, -0.5), t1)):
        for integ in (api.PathIntegrator(3, 1.0), api.DirectLightingIntegrator(3)):
            scene, camera, film = scenes.textured_mirror_probe(backend=backend, xz=xz)
            api.SamplerIntegrator(camera, integ).render_parallel(scene, film, api.RandomSampler.new_with_seed(2, 0))
            assert np.allclose(film.into_spectrum_buffer()[0], want[None, :], rtol=1e-5), (xz, type(integ).__name__)
    # an image texture as Kr: level 0 of a ramp, bilinear (the footprint of the narrow camera is far below a texel)
    n = 16
    ramp = np.zeros((n, n, 3), np.float32)
    ramp[..., 0] = (np.arange(n)[None, :] + 0.5) / n; ramp[..., 1] = (np.arange(n)[:, None] + 0.5) / n; ramp[..., 2] = 0.5
    tex = api.ImageTexture(api.MIPMap(ramp, "clamp"), api.UVMapping(1.0 / 12.0, 1.0 / 12.0, 0.5, 0.5))
    scene, camera, film = scenes.textured_mirror_probe(backend=backend, xz=(2.0, -3.0), texture=tex)
    api.SamplerIntegrator(camera, api.PathIntegrator(3, 1.0)).render_parallel(scene, film, api.RandomSampler.new_with_seed(2, 0))
    assert np.allclose(film.into_spectrum_buffer()[0].mean(axis=0), [2.0 / 12 + 0.5, -3.0 / 12 + 0.5, 0.5], rtol=3e-3)


def test_mirror_textured_kr_closed_form(orc_backend):
    mirror_kr_closed_form(orc_backend)


# ---- ray differentials behind a mirror under the direct-lighting integrator (integrator/mod.rs:59-83) -------------------
# No reference test renders one.  Closed form: a flat mirror unfolds the path, so the ceiling seen in a mirror on the floor
# is filtered with the footprint of a camera h_cam + h_ceiling away (the level formula of the floor test above); the radiance
# is Kr * c(level) / pi * L * cos(60 deg).  Without the mirrored differentials the lookup would be bilinear at level 0 (c = .1).
def mirror_differentials_closed_form(backend, uscales=(1.0, 3.0, 11.0)):
    mp = constant_level_mipmap()
    n = len(mp.levels)
    for uscale in uscales:
        tex = api.ImageTexture(mp, api.UVMapping(uscale, uscale, 0.3, 0.1))
        scene, camera, film = scenes.mirror_ceiling_probe(backend=backend, texture=tex)
        api.SamplerIntegrator(camera, api.DirectLightingIntegrator(3)).render_parallel(scene, film, api.RandomSampler.new_with_seed(4, 0))
        rgb = film.into_spectrum_buffer()[0]
        level = min(max(expected_mip_level(n, uscale, 0.5, 5, 30.0, 4), 0.0), n - 1.0)
        assert level > 0.5                                     # the case tells the mirrored footprint from a point lookup
        assert np.allclose(rgb, (0.1 + 0.1 * level) / np.pi * 3.0 * 0.5, rtol=3e-3), (uscale, level, rgb[:2])


def test_mirror_differentials_closed_form(orc_backend):
    mirror_differentials_closed_form(orc_backend)


def test_mirror_differentials_stop_with_the_recursion(orc_backend):
    """max_depth 1: specular_reflect is not called (direct_lighting.rs:93), the mirror is black."""
    mp = constant_level_mipmap()
    scene, camera, film = scenes.mirror_ceiling_probe(backend=orc_backend, texture=api.ImageTexture(mp, api.UVMapping(3.0, 3.0, 0.0, 0.0)))
    api.SamplerIntegrator(camera, api.DirectLightingIntegrator(1)).render_parallel(scene, film, api.RandomSampler.new_with_seed(2, 0))
    assert np.all(film.into_spectrum_buffer()[0] == 0.0)
