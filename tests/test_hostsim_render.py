"""CPU tier: the wavefront path logic (ftn_path.cuh, host-compiled) against the oracle --
tests/furnace.rs thresholds, counter-sampler image A/B, sample sharding, film footprint."""
import numpy as np
import pytest

from fountain_b200 import _abi as A
from fountain_b200 import api
from workloads import scenes
from fountain_b200.transform import Transform
from tests import parity


@pytest.fixture(scope="module")
def sim_backend():
    from tests.hostsim import sim
    return sim.backend()


# ---- tests/furnace.rs on the device path logic ------------------------------------------------
def test_furnace_path(sim_backend):
    rgb, _, _ = parity.render(sim_backend, scenes.furnace_scene, api.PathIntegrator(10, 1.0), 128)
    assert np.all(np.abs(rgb - 2.0) <= 0.1)                      # furnace.rs:20


def test_furnace_path_no_rr(sim_backend, orc_backend):
    rgb, px, st = parity.render(sim_backend, scenes.furnace_scene, api.PathIntegrator(10, 0.0), 128)
    assert np.all(np.abs(rgb - 2.0) <= 0.001)                    # furnace.rs:36
    assert st["rays_closest"] == 16 * 16 * 128 * 21 and st["rays_any"] == 16 * 16 * 128 * 10
    ref, rpx, rst = parity.render(orc_backend, scenes.furnace_scene, api.PathIntegrator(10, 0.0), 128)
    assert np.allclose(rgb, ref, rtol=2e-5, atol=0)
    assert np.array_equal(px[..., 3], rpx[..., 3])


def test_furnace_directlighting(sim_backend):
    rgb, _, _ = parity.render(sim_backend, scenes.furnace_scene, api.DirectLightingIntegrator(3), 128)
    assert np.all(np.abs(rgb - 1.5) <= 0.00001)                  # furnace.rs:55


# ---- image A/B at equal counter streams ----------------------------------------------------------
@pytest.mark.parametrize("material", ["matte", "metal", "plastic"])
def test_cube_image_matches_oracle(sim_backend, orc_backend, material):
    mat = {"matte": lambda: api.MatteMaterial((0.5, 0.4, 0.6)),
           "metal": lambda: api.MetalMaterial((0.2, 0.92, 1.1), (3.9, 2.45, 2.14), roughness=0.1),
           "plastic": lambda: api.PlasticMaterial(0.3, 0.4, 0.15)}[material]
    kw = dict(resolution=(40, 40))
    a, apx, ast = parity.render(sim_backend, lambda backend, **k: scenes.rounded_cube_scene(backend=backend, material=mat(), **k), api.PathIntegrator(5, 1.0), 4, seed=3, **kw)
    b, bpx, bst = parity.render(orc_backend, lambda backend, **k: scenes.rounded_cube_scene(backend=backend, material=mat(), **k), api.PathIntegrator(5, 1.0), 4, seed=3, **kw)
    mean_rel, frac_off = parity.image_diff(a, b)
    assert mean_rel < 2e-3 and frac_off < 0.02, (mean_rel, frac_off)
    assert np.array_equal(apx[..., 3], bpx[..., 3])
    # identical control flow almost everywhere => nearly identical ray counts
    assert abs(ast["rays_closest"] - bst["rays_closest"]) <= 0.002 * bst["rays_closest"]
    assert abs(ast["rays_any"] - bst["rays_any"]) <= 0.002 * bst["rays_any"]


def test_envmap_thin_lens_image_matches_oracle(sim_backend, orc_backend):
    def build(backend):
        env = api.InfiniteAreaLight.new_envmap(scenes.sky_sun_envmap(64, 32, peak=50.0), Transform.rotate(20, (0, 0, 1)))
        scene, camera, film = scenes.rounded_cube_scene(backend=backend, resolution=(36, 24), light=env,
                                                        material=api.MetalMaterial((0.2, 0.92, 1.1), (3.9, 2.45, 2.14), roughness=0.3))
        camera = api.PerspectiveCamera(camera.camera_to_world, (36, 24), fov=40.0, lens_radius=0.5, focal_dist=32.0)
        return scene, camera, film
    a, _, _ = parity.render(sim_backend, lambda backend: build(backend), api.PathIntegrator(4, 1.0), 4, seed=9)
    b, _, _ = parity.render(orc_backend, lambda backend: build(backend), api.PathIntegrator(4, 1.0), 4, seed=9)
    mean_rel, frac_off = parity.image_diff(a, b)
    assert mean_rel < 5e-3 and frac_off < 0.03, (mean_rel, frac_off)


def test_null_material_is_skipped(sim_backend, orc_backend):
    """A mesh without a material has a null BSDF: paths pass through (path.rs:76-80)."""
    def build_null(backend):
        mesh = api.TriangleMesh.from_ply(scenes.ROUNDED_CUBE_PLY)
        scene = api.Scene([api.GeometricPrimitive(mesh, None)], [api.InfiniteAreaLight.new_uniform(1.0)], backend=backend)
        _, camera, film = scenes.rounded_cube_scene(backend=backend, resolution=(24, 24))
        return scene, camera, film
    a, _, _ = parity.render(sim_backend, build_null, api.PathIntegrator(5, 1.0), 2)
    b, _, _ = parity.render(orc_backend, build_null, api.PathIntegrator(5, 1.0), 2)
    assert np.allclose(a, 1.0, atol=1e-5) and np.allclose(b, 1.0, atol=1e-5)    # sees the environment through the cube


def test_sample_sharding_partitions_the_render(sim_backend):
    scene, camera, film = scenes.rounded_cube_scene(backend=sim_backend, resolution=(24, 24))
    integ = api.SamplerIntegrator(camera, api.PathIntegrator(5, 1.0))
    sampler = api.RandomSampler.new_with_seed(8, 5)
    integ.render_parallel(scene, film, sampler)
    full = film.pixels.copy()
    acc = np.zeros_like(full)
    for rank in range(4):
        integ.render_parallel(scene, film, sampler, sample_begin=rank, sample_stride=4)
        assert np.all(film.pixels[..., 3] == 2.0)
        acc += film.pixels
    assert np.allclose(acc, full, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("radius,crop", [((0.5, 0.5), ((0.0, 0.0), (1.0, 1.0))), ((1.5, 1.0), ((0.0, 0.0), (1.0, 1.0))),
                                         ((0.5, 0.5), ((0.25, 0.1), (0.8, 0.75))), ((0.3, 0.3), ((0.0, 0.0), (1.0, 1.0)))])
def test_film_footprint_matches_oracle(sim_backend, orc_backend, radius, crop):
    """Filter radius / crop window variants of Film::add_sample_to_tile + get_film_tile (incl. the
    tile clipping and the `- radius` quirk of film.rs:100): weights are small integers -> exact."""
    def build(backend):
        scene, camera, _ = scenes.rounded_cube_scene(backend=backend, resolution=(40, 36))
        film = api.Film((40, 36), crop_window=crop, filter=api.BoxFilter(radius), backend=backend)
        return scene, camera, film
    a, apx, _ = parity.render(sim_backend, build, api.PathIntegrator(1, 1.0), 3, seed=2)
    b, bpx, _ = parity.render(orc_backend, build, api.PathIntegrator(1, 1.0), 3, seed=2)
    assert apx.shape == bpx.shape
    assert np.array_equal(apx[..., 3], bpx[..., 3])                   # filter-weight sums: exact
    assert np.allclose(apx[..., :3], bpx[..., :3], rtol=2e-4, atol=1e-5)


def test_reference_stream_is_rejected(sim_backend):
    scene, camera, film = scenes.furnace_scene(backend=sim_backend)
    with pytest.raises(api.FountainError) as e:
        api.SamplerIntegrator(camera, api.PathIntegrator(2, 1.0)).render_parallel(
            scene, film, api.RandomSampler.new_with_seed(1, 0, mode=A.FTN_SAMPLER_REFERENCE_TILE_STREAM))
    assert e.value.code == A.FTN_ERR_UNSUPPORTED


# ---- delta lights (point.rs, distant.rs): device path logic vs oracle, same counter stream -----------
@pytest.mark.parametrize("integrator", ["path", "direct"])
def test_delta_lights_image_matches_oracle(sim_backend, orc_backend, integrator):
    integ = api.PathIntegrator(4, 1.0) if integrator == "path" else api.DirectLightingIntegrator(3)
    a, apx, ast = parity.render(sim_backend, scenes.delta_lights_scene, integ, 4, seed=5, resolution=(40, 40))
    b, bpx, bst = parity.render(orc_backend, scenes.delta_lights_scene, integ, 4, seed=5, resolution=(40, 40))
    mean_rel, frac_off = parity.image_diff(a, b)
    assert mean_rel < 2e-3 and frac_off < 0.02, (mean_rel, frac_off)
    assert np.array_equal(apx[..., 3], bpx[..., 3])
    assert ast["rays_any"] == bst["rays_any"] and ast["rays_closest"] == bst["rays_closest"]
    assert b.max() > 0.1 and (b == 0.0).any()            # lit floor, black shadow / background


# ---- mirror: SpecularReflection lobe, specular-bounce flag, direct-lighting specular chain ------------
@pytest.mark.parametrize("integrator", ["path", "direct2", "direct4"])
def test_mirror_image_matches_oracle(sim_backend, orc_backend, integrator):
    integ = {"path": api.PathIntegrator(5, 1.0), "direct2": api.DirectLightingIntegrator(2), "direct4": api.DirectLightingIntegrator(4)}[integrator]
    a, apx, ast = parity.render(sim_backend, scenes.mirror_scene, integ, 4, seed=7, resolution=(40, 40))
    b, bpx, bst = parity.render(orc_backend, scenes.mirror_scene, integ, 4, seed=7, resolution=(40, 40))
    mean_rel, frac_off = parity.image_diff(a, b)
    assert mean_rel < 2e-3 and frac_off < 0.02, (mean_rel, frac_off)
    assert np.array_equal(apx[..., 3], bpx[..., 3])
    assert abs(ast["rays_closest"] - bst["rays_closest"]) <= 0.002 * bst["rays_closest"]
    assert abs(ast["rays_any"] - bst["rays_any"]) <= 0.002 * bst["rays_any"]


# ---- textured Kd (checkerboard / uv through UVMapping) -----------------------------------------------------
@pytest.mark.parametrize("texture,material", [("checkerboard", "matte"), ("checkerboard_scaled", "plastic"), ("uv", "matte")])
def test_textured_floor_matches_oracle(sim_backend, orc_backend, texture, material):
    kw = dict(resolution=(40, 40), texture=texture, material=material)
    a, apx, ast = parity.render(sim_backend, scenes.textured_floor_scene, api.PathIntegrator(3, 1.0), 4, seed=11, **kw)
    b, bpx, bst = parity.render(orc_backend, scenes.textured_floor_scene, api.PathIntegrator(3, 1.0), 4, seed=11, **kw)
    mean_rel, frac_off = parity.image_diff(a, b)
    assert mean_rel < 2e-3 and frac_off < 0.02, (mean_rel, frac_off)
    assert np.array_equal(apx[..., 3], bpx[..., 3])
    assert len(np.unique(np.round(b.reshape(-1, 3), 3), axis=0)) > 4       # the texture is visible


# ---- Oren-Nayar (its own material class on the device) ------------------------------------------------------
def test_oren_nayar_image_matches_oracle(sim_backend, orc_backend):
    build = lambda backend, **k: scenes.rounded_cube_scene(backend=backend, material=api.MatteMaterial((0.6, 0.5, 0.4), sigma=35.0), **k)
    a, apx, ast = parity.render(sim_backend, build, api.PathIntegrator(5, 1.0), 4, seed=3, resolution=(40, 40))
    b, bpx, bst = parity.render(orc_backend, build, api.PathIntegrator(5, 1.0), 4, seed=3, resolution=(40, 40))
    mean_rel, frac_off = parity.image_diff(a, b)
    assert mean_rel < 2e-3 and frac_off < 0.02, (mean_rel, frac_off)
    lam, _, _ = parity.render(orc_backend, lambda backend, **k: scenes.rounded_cube_scene(backend=backend, material=api.MatteMaterial((0.6, 0.5, 0.4)), **k),
                              api.PathIntegrator(5, 1.0), 4, seed=3, resolution=(40, 40))
    assert parity.image_diff(b, lam)[0] > 1e-2            # and it is not the Lambert image


# ---- image-textured Kd (ImageTexture + MIPMap::lookup_trilinear with the camera ray's differentials) --------
@pytest.mark.parametrize("wrap,material,integ,lens", [("repeat", "matte", "path", 0.0), ("clamp", "plastic", "path", 0.15),
                                                      ("black", "oren_nayar", "direct", 0.0)])
def test_image_texture_matches_oracle(sim_backend, orc_backend, wrap, material, integ, lens):
    kw = dict(resolution=(40, 40), wrap=wrap, material=material, lens_radius=lens)
    integrator = api.PathIntegrator(3, 1.0) if integ == "path" else api.DirectLightingIntegrator(2)
    a, apx, ast = parity.render(sim_backend, scenes.image_texture_scene, integrator, 4, seed=13, **kw)
    b, bpx, bst = parity.render(orc_backend, scenes.image_texture_scene, integrator, 4, seed=13, **kw)
    mean_rel, frac_off = parity.image_diff(a, b)
    assert mean_rel < 2e-3 and frac_off < 0.02, (mean_rel, frac_off)
    assert np.array_equal(apx[..., 3], bpx[..., 3])
    assert len(np.unique(np.round(b.reshape(-1, 3), 2), axis=0)) > 50       # the texture is visible


def test_image_texture_level_selection_closed_form_hostsim(sim_backend):
    from tests.test_oracle_render import constant_level_mipmap, expected_mip_level
    mp = constant_level_mipmap()
    for uscale in (1.0, 3.0, 11.0):
        tex = api.ImageTexture(mp, api.UVMapping(uscale, uscale, 0.3, 0.1))
        scene, camera, film = scenes.textured_floor_scene(backend=sim_backend, resolution=(5, 5), texture=tex, look_at=(0.7, -0.4), fov=0.5)
        api.SamplerIntegrator(camera, api.DirectLightingIntegrator(1)).render_parallel(scene, film, api.RandomSampler.new_with_seed(4, 0))
        level = min(max(expected_mip_level(len(mp.levels), uscale, 0.5, 5, 30.0, 4), 0.0), len(mp.levels) - 1.0)
        assert np.allclose(film.into_spectrum_buffer()[0], (0.1 + 0.1 * level) / np.pi * 3.0, rtol=2e-3), uscale


def test_mirror_textured_kr_closed_form(sim_backend):
    from tests.test_oracle_render import mirror_kr_closed_form
    mirror_kr_closed_form(sim_backend)


# ---- emissive triangle meshes: DiffuseAreaLight<Triangle> (light/diffuse.rs, triangle.rs:395-420, shapes/mod.rs:41-66) ----
@pytest.mark.parametrize("integrator", ["path", "direct"])
def test_quad_light_image_matches_oracle(sim_backend, orc_backend, integrator):
    integ = api.PathIntegrator(4, 1.0) if integrator == "path" else api.DirectLightingIntegrator(3)
    order = parity.quad_light_order(orc_backend)
    assert order is not None
    kw = dict(resolution=(40, 40), light_order=order)
    a, apx, ast = parity.render(sim_backend, scenes.quad_light_scene, integ, 4, seed=3, **kw)
    b, bpx, bst = parity.render(orc_backend, scenes.quad_light_scene, integ, 4, seed=3, **kw)
    mean_rel, frac_off = parity.image_diff(a, b)
    assert mean_rel < 2e-3 and frac_off < 0.02, (mean_rel, frac_off)
    assert np.array_equal(apx[..., 3], bpx[..., 3])
    assert ast["rays_any"] == bst["rays_any"] and ast["rays_closest"] == bst["rays_closest"]
    assert b.max() > 1.0 and (b == 0.0).any()            # the emitter is seen directly; the background is black


@pytest.mark.parametrize("integrator", ["path", "direct"])
def test_quad_light_closed_form(sim_backend, integrator):
    integ = api.PathIntegrator(1, 1.0) if integrator == "path" else api.DirectLightingIntegrator(1)
    scene, camera, film, expected = scenes.quad_light_probe(backend=sim_backend, side=1.0, height=2.0, emit=5.0, kd=0.5)
    api.SamplerIntegrator(camera, integ).render_parallel(scene, film, api.RandomSampler.new_with_seed(4096, 11))
    rgb, _ = film.into_spectrum_buffer()
    assert abs(rgb.mean() - expected) < 0.01 * expected, (rgb.mean(), expected)


def test_two_quad_lights_statistical_parity(sim_backend, orc_backend):
    """Two emissive meshes (4 area lights): the reference enumerates them in its BVH's order, the ABI in primitive order,
    so the same random number picks different lights and the comparison is statistical -- the image must sit inside the
    oracle's own seed-to-seed noise."""
    integ = api.PathIntegrator(3, 1.0)
    kw = dict(resolution=(24, 24), two_lights=True)
    a, _, ast = parity.render(sim_backend, scenes.quad_light_scene, integ, 256, seed=1, **kw)
    b, _, bst = parity.render(orc_backend, scenes.quad_light_scene, integ, 256, seed=1, **kw)
    b2, _, _ = parity.render(orc_backend, scenes.quad_light_scene, integ, 256, seed=2, **kw)
    assert abs(a.mean() - b.mean()) < 0.02 * b.mean(), (a.mean(), b.mean())
    assert parity.rel_mse(a, b) <= 1.5 * parity.rel_mse(b2, b), (parity.rel_mse(a, b), parity.rel_mse(b2, b))


# ---- rough glass: MicrofacetReflection<FresnelDielectric(1, eta)> + MicrofacetTransmission (glass.rs:52-96, reflection/mod.rs:365-436) ----
@pytest.mark.parametrize("integrator,roughness,remap", [("path", 0.25, True), ("path", 0.0, True), ("direct", 0.1, False)])
def test_rough_glass_image_matches_oracle(sim_backend, orc_backend, integrator, roughness, remap):
    integ = api.PathIntegrator(6, 1.0) if integrator == "path" else api.DirectLightingIntegrator(3)
    kw = dict(resolution=(40, 40), roughness=roughness, remap=remap)
    a, apx, ast = parity.render(sim_backend, scenes.rough_glass_scene, integ, 4, seed=7, **kw)
    b, bpx, bst = parity.render(orc_backend, scenes.rough_glass_scene, integ, 4, seed=7, **kw)
    mean_rel, frac_off = parity.image_diff(a, b)
    assert mean_rel < 3e-3 and frac_off < 0.03, (mean_rel, frac_off)
    assert np.array_equal(apx[..., 3], bpx[..., 3])
    assert abs(ast["rays_closest"] - bst["rays_closest"]) <= 0.002 * bst["rays_closest"]
    assert b.max() > 0.1


def test_smooth_glass_is_rejected(sim_backend):
    """Alphas of exactly 0 = FresnelSpecular, `todo!()` in the reference (glass.rs:66): refused at scene creation."""
    with pytest.raises(api.FountainError) as e:
        scenes.rough_glass_scene(backend=sim_backend, roughness=0.0, remap=False)
    assert e.value.code == A.FTN_ERR_UNSUPPORTED


# ---- the texture table: Ks, eta, k, roughnesses, sigma, Kt, the glass index (loaders/constructors.rs:192-238) -------------
@pytest.mark.parametrize("integrator", ["path", "direct"])
def test_textured_parameters_match_oracle(sim_backend, orc_backend, integrator):
    integ = api.PathIntegrator(5, 1.0) if integrator == "path" else api.DirectLightingIntegrator(3)
    a, apx, ast = parity.render(sim_backend, scenes.textured_params_scene, integ, 4, seed=9, resolution=(56, 40))
    b, bpx, bst = parity.render(orc_backend, scenes.textured_params_scene, integ, 4, seed=9, resolution=(56, 40))
    mean_rel, frac_off = parity.image_diff(a, b)
    assert mean_rel < 3e-3 and frac_off < 0.03, (mean_rel, frac_off)
    assert np.array_equal(apx[..., 3], bpx[..., 3])
    assert abs(ast["rays_closest"] - bst["rays_closest"]) <= 0.002 * bst["rays_closest"]


def test_texture_table_kd_equals_inline_slot_and_constants(sim_backend):
    """Kd / Kr through the table == through the materials' inline slot, bit for bit; and the table is really read: the
    constant-texture variant of the scene gives a different image."""
    integ = api.PathIntegrator(4, 1.0)
    a, _, _ = parity.render(sim_backend, scenes.textured_params_scene, integ, 2, seed=1, resolution=(40, 28), variant="table")
    b, _, _ = parity.render(sim_backend, scenes.textured_params_scene, integ, 2, seed=1, resolution=(40, 28), variant="inline")
    c, _, _ = parity.render(sim_backend, scenes.textured_params_scene, integ, 2, seed=1, resolution=(40, 28), variant="constant")
    assert np.array_equal(a, b)
    assert not np.allclose(a, c, rtol=1e-3)


# ---- ray differentials behind a mirror under the direct-lighting integrator (integrator/mod.rs:59-83) -------------------
def test_mirror_differentials_closed_form(sim_backend):
    from tests.test_oracle_render import mirror_differentials_closed_form
    mirror_differentials_closed_form(sim_backend)


@pytest.mark.parametrize("depth", [2, 4])
def test_mirrored_image_texture_matches_oracle(sim_backend, orc_backend, depth):
    """A splayed-normal mirror quad (triangle dndu / dndv) and a mirror sphere (Weingarten dndu / dndv) over the image-textured
    floor: the differentials specular_reflect derives select the same mip levels on both sides."""
    integrator = api.DirectLightingIntegrator(depth)
    a, apx, _ = parity.render(sim_backend, scenes.mirrored_image_texture_scene, integrator, 4, seed=21, resolution=(40, 40))
    b, bpx, _ = parity.render(orc_backend, scenes.mirrored_image_texture_scene, integrator, 4, seed=21, resolution=(40, 40))
    mean_rel, frac_off = parity.image_diff(a, b)
    assert mean_rel < 2e-3 and frac_off < 0.02, (mean_rel, frac_off)
    assert np.array_equal(apx[..., 3], bpx[..., 3])
