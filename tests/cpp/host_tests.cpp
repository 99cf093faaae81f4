// C++ host-side tests above the C ABI, written to read like the reference's own integration
// tests: tests/furnace.rs (three white-furnace renders) and tests/tri_watertight.rs (100k
// random directions from inside rounded_cube.ply must hit, any-hit and closest-hit).
//
//   host_tests <library.so> <symbol prefix> <path to rounded_cube.ply> [test name ...]
//
// The `-m gpu` tier runs it against fountain_b200/csrc/libfountain_gpu.so ("ftn_"); the CPU
// tier runs the SAME host code against the checker library to cover the host logic
// (scene flattening, transforms, camera matrices, PLY reading, film conversion).
#include <cstdio>
#include <cstring>
#include <functional>
#include <map>
#include <random>
#include <string>

#include "fountain_host.hpp"

using namespace fountain;

static LibraryPtr g_lib;
static std::string g_ply;
static int g_failures = 0;

#define CHECK(cond, ...)                                                         \
    do {                                                                         \
        if (!(cond)) {                                                           \
            std::fprintf(stderr, "  CHECK failed %s:%d: %s -- ", __FILE__, __LINE__, #cond); \
            std::fprintf(stderr, __VA_ARGS__);                                   \
            std::fprintf(stderr, "\n");                                          \
            ++g_failures;                                                        \
            return;                                                              \
        }                                                                        \
    } while (0)

// ---- tests/furnace.rs ----------------------------------------------------------------------
// testscenes/furnace_empty.pbrt: LookAt 0 -2 0  0 0 0  0 0 1; Camera perspective fov 60;
// Film 16x16; Sampler random 128 spp; ReverseOrientation sphere r=100, matte Kd .5, diffuse
// area light L 1.
template <class R>
static std::pair<std::vector<Spectrum>, std::pair<uint32_t, uint32_t>> do_render(R radiance) {
    auto sphere = std::make_shared<Sphere>(Transform::identity(), /*reverse_orientation=*/true, 100.0f);
    std::vector<GeometricPrimitive> prims;
    prims.emplace_back(sphere, std::make_shared<MatteMaterial>(Spectrum(0.5f)), std::make_shared<DiffuseAreaLight>(Spectrum(1.0f)));
    Scene scene(g_lib, prims);

    Transform camera_to_world = Transform::look_at({0, -2, 0}, {0, 0, 0}, {0, 0, 1}).inverse();   // pbrt.rs:430
    PerspectiveCamera camera(camera_to_world, 16, 16, 60.0f);
    RandomSampler sampler = RandomSampler::new_with_seed(128, 0);
    Film film(g_lib, 16, 16);

    SamplerIntegrator<R> integrator(camera, radiance);
    integrator.render_parallel(scene, film, sampler);
    return film.into_spectrum_buffer();
}

static void furnace_test_path() {
    auto [img, wh] = do_render(PathIntegrator(10, 1.0f));
    CHECK(wh.first == 16 && wh.second == 16 && img.size() == 256, "film size");
    const float expected = 1.0f / (1.0f - 0.5f);
    for (const Spectrum& s : img)
        for (float comp : s.into_array())        // Russian roulette causes some variance
            CHECK(std::fabs(comp - expected) <= 0.1f, "%.9g", comp);
}

static void furnace_test_path_no_rr() {
    auto [img, wh] = do_render(PathIntegrator(10, 0.0f));
    const float expected = 1.0f / (1.0f - 0.5f);
    for (const Spectrum& s : img)
        for (float comp : s.into_array())        // no roulette: the same value for every sample
            CHECK(std::fabs(comp - expected) <= 0.001f, "%.9g", comp);
}

static void furnace_test_directlighting() {
    auto [img, wh] = do_render(DirectLightingIntegrator(LightStrategy::UniformSampleOne, 3));
    const float expected = 1.0f + 0.5f;
    for (const Spectrum& s : img)
        for (float comp : s.into_array())
            CHECK(std::fabs(comp - expected) <= 0.00001f, "%.9g", comp);
}

static void uniform_sample_all_is_unimplemented() {
    bool threw = false;
    try { DirectLightingIntegrator d(LightStrategy::UniformSampleAll, 3); } catch (const Error& e) { threw = e.code == FTN_ERR_UNSUPPORTED; }
    CHECK(threw, "UniformSampleAll must be rejected");
}

// ---- tests/tri_watertight.rs ------------------------------------------------------------------
static void test_rounded_cube() {
    auto mesh = TriangleMesh::from_ply(g_ply, Transform::identity(), false);
    CHECK(mesh->n_triangles() == 4332, "%zu triangles", mesh->n_triangles());
    std::vector<GeometricPrimitive> prims;
    prims.emplace_back(mesh, nullptr);            // material: None, light: None
    Scene scene(g_lib, prims);

    // UnitSphereSurface: normalised Gaussian triples
    std::mt19937_64 rng(0x5EEDu);
    std::normal_distribution<double> gauss;
    std::vector<Ray> rays;
    const size_t n = 100000;
    rays.reserve(n);
    while (rays.size() < n) {
        double x = gauss(rng), y = gauss(rng), z = gauss(rng), l = std::sqrt(x * x + y * y + z * z);
        if (l < 1e-12) continue;
        rays.emplace_back(Point3f(0, 0, 0), Vec3f((float)(x / l), (float)(y / l), (float)(z / l)));
    }
    auto any = scene.intersect_test(rays);
    auto hits = scene.intersect(rays);
    for (size_t i = 0; i < n; ++i) {
        CHECK(any[i], "ray %zu: intersect_test missed", i);
        CHECK(hits[i].prim != FTN_NO_HIT, "ray %zu: Did not intersect", i);
        CHECK(hits[i].prim < mesh->n_triangles() && hits[i].t > 0.0f && std::isfinite(hits[i].t), "ray %zu: bad hit", i);
    }
}

// ---- host-logic checks with no counterpart file in the reference -----------------------------------
// camera/mod.rs:362-380 test_fov pins the projection: raster corners map to +-tan(fov/2) on the z=1 plane
static void camera_raster_to_camera_fov() {
    PerspectiveCamera cam(Transform::identity(), 200, 200, 90.0f);
    const auto& m = cam.raster_to_camera.m;
    auto apply = [&](double x, double y) {
        double c[4];
        for (int r = 0; r < 4; ++r) c[r] = m[4 * r] * x + m[4 * r + 1] * y + m[4 * r + 3];
        return std::array<double, 3>{c[0] / c[3], c[1] / c[3], c[2] / c[3]};
    };
    auto a = apply(0, 100), b = apply(200, 100);      // left and right edge, mid height
    double cosang = (a[0] * b[0] + a[1] * b[1] + a[2] * b[2]) /
                    (std::sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]) * std::sqrt(b[0] * b[0] + b[1] * b[1] + b[2] * b[2]));
    double fov = std::acos(cosang) * 180.0 / M_PI;
    CHECK(std::fabs(fov - 90.0) < 0.01, "fov %.6f", fov);
}

static void world_bound_and_morton_order() {
    auto mesh = TriangleMesh::from_ply(g_ply, Transform::translate(1, 2, 3) * Transform::scale(2, 2, 2), false);
    std::vector<GeometricPrimitive> prims;
    prims.emplace_back(mesh, std::make_shared<MatteMaterial>());
    Scene scene(g_lib, prims, {InfiniteAreaLight::new_uniform(Spectrum(1.0f))});
    auto [lo, hi] = scene.world_bound();
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (size_t i = 0; i < mesh->vertices.size(); ++i) {
        mn[i % 3] = std::min(mn[i % 3], mesh->vertices[i]);
        mx[i % 3] = std::max(mx[i % 3], mesh->vertices[i]);
    }
    CHECK(lo.x == mn[0] && lo.y == mn[1] && lo.z == mn[2] && hi.x == mx[0] && hi.y == mx[1] && hi.z == mx[2], "world bound");
    std::vector<uint32_t> codes, order;
    scene.morton_codes_and_order(codes, order);
    CHECK(codes.size() == 4332 && order.size() == 4332, "sizes");
    std::vector<uint8_t> seen(order.size(), 0);
    for (size_t i = 0; i < order.size(); ++i) {
        CHECK(order[i] < order.size() && !seen[order[i]], "order is not a permutation at %zu", i);
        seen[order[i]] = 1;
        if (i) {
            uint32_t a = codes[order[i - 1]], b = codes[order[i]];
            CHECK(a < b || (a == b && order[i - 1] < order[i]), "not sorted (stable) at %zu", i);
        }
        CHECK(codes[i] < (1u << 30), "30-bit code");
    }
}

// checkerboard Kd (texture/checkerboard.rs) + distant light (light/distant.rs) + mirror (material/mirror.rs):
// closed forms through the C++ host side -- Kd(texel) / pi * L under a light from straight above, Kr * L_env
static void textured_floor_and_mirror_closed_forms() {
    const float pi = 3.14159265358979f;
    auto probe = [&](double x, double y) {
        std::vector<float> v = {-6, -6, 0, 6, -6, 0, 6, 6, 0, -6, 6, 0}, uv = {-6, -6, 6, -6, 6, 6, -6, 6};
        auto mesh = std::make_shared<TriangleMesh>(Transform::identity(), std::vector<uint32_t>{0, 1, 2, 0, 2, 3}, v, std::vector<float>{}, uv);
        auto mat = std::make_shared<MatteMaterial>(SpectrumTexture::checkerboard(Spectrum(0.8f, 0.2f, 0.2f), Spectrum(0.1f, 0.1f, 0.9f)));
        std::vector<GeometricPrimitive> prims; prims.emplace_back(mesh, mat);
        Scene scene(g_lib, prims, {Light(DistantLight::from_to(Point3f(0, 0, 1), Point3f(0, 0, 0), Spectrum(3.0f)))});
        PerspectiveCamera camera(Transform::look_at({x, y, 30.0}, {x, y, 0.0}, {0, 1, 0}).inverse(), 5, 5, 0.5f);
        Film film(g_lib, 5, 5);
        SamplerIntegrator<DirectLightingIntegrator> integrator(camera, DirectLightingIntegrator(LightStrategy::UniformSampleOne, 1));
        integrator.render_parallel(scene, film, RandomSampler::new_with_seed(2, 0));
        return film.into_spectrum_buffer().first;
    };
    for (const Spectrum& s : probe(0.5, 0.5)) CHECK(std::fabs(s.r - 0.8f / pi * 3.0f) < 1e-5f && std::fabs(s.b - 0.2f / pi * 3.0f) < 1e-5f, "tex1 %.7g %.7g", s.r, s.b);
    for (const Spectrum& s : probe(-0.5, 0.5)) CHECK(std::fabs(s.r - 0.1f / pi * 3.0f) < 1e-5f && std::fabs(s.b - 0.9f / pi * 3.0f) < 1e-5f, "tex2 %.7g %.7g", s.r, s.b);
    // mirror quad facing the camera under a uniform environment of radiance 1
    std::vector<float> v = {-3, 2, -1, 3, 2, -1, 3, 2, 3, -3, 2, 3};
    auto mesh = std::make_shared<TriangleMesh>(Transform::identity(), std::vector<uint32_t>{0, 1, 2, 0, 2, 3}, v);
    std::vector<GeometricPrimitive> prims; prims.emplace_back(mesh, std::make_shared<MirrorMaterial>(Spectrum(0.8f, 0.7f, 0.6f)));
    Scene scene(g_lib, prims, {Light(InfiniteAreaLight::new_uniform(Spectrum(1.0f)))});
    PerspectiveCamera camera(Transform::look_at({0, -7, 1.5}, {0, 2, 1.0}, {0, 0, 1}).inverse(), 9, 9, 5.0f);
    Film film(g_lib, 9, 9);
    SamplerIntegrator<PathIntegrator> integrator(camera, PathIntegrator(5, 1.0f));
    integrator.render_parallel(scene, film, RandomSampler::new_with_seed(4, 0));
    for (const Spectrum& s : film.into_spectrum_buffer().first)
        CHECK(std::fabs(s.r - 0.8f) < 1e-5f && std::fabs(s.g - 0.7f) < 1e-5f && std::fabs(s.b - 0.6f) < 1e-5f, "mirror %.7g %.7g %.7g", s.r, s.g, s.b);
}

// The texture table (FtnSceneDesc::textures): Kd of a matte floor through the table gives the closed forms of the inline
// slot; constant textures for plastic's Ks / roughness and rough glass's Kt / index give the image of the plain constants.
static void texture_table_parameters() {
    const float pi = 3.14159265358979f;
    auto probe = [&](std::shared_ptr<Material> mat, double x, double y) {
        std::vector<float> v = {-6, -6, 0, 6, -6, 0, 6, 6, 0, -6, 6, 0}, uv = {-6, -6, 6, -6, 6, 6, -6, 6};
        auto mesh = std::make_shared<TriangleMesh>(Transform::identity(), std::vector<uint32_t>{0, 1, 2, 0, 2, 3}, v, std::vector<float>{}, uv);
        std::vector<GeometricPrimitive> prims; prims.emplace_back(mesh, mat);
        Scene scene(g_lib, prims, {Light(DistantLight::from_to(Point3f(0.3, 0, 1), Point3f(0, 0, 0), Spectrum(3.0f))), Light(InfiniteAreaLight::new_uniform(Spectrum(0.3f)))});
        PerspectiveCamera camera(Transform::look_at({x, y - 6.0, 20.0}, {x, y, 0.0}, {0, 1, 0}).inverse(), 6, 6, 1.0f);
        Film film(g_lib, 6, 6);
        SamplerIntegrator<PathIntegrator> integrator(camera, PathIntegrator(3, 1.0f));
        integrator.render_parallel(scene, film, RandomSampler::new_with_seed(8, 3));
        return film.into_spectrum_buffer().first;
    };
    auto same = [&](const std::vector<Spectrum>& a, const std::vector<Spectrum>& b, const char* what) {
        bool eq = a.size() == b.size();
        for (size_t i = 0; eq && i < a.size(); ++i) eq = a[i].r == b[i].r && a[i].g == b[i].g && a[i].b == b[i].b;
        CHECK(eq, "%s: the table's constant textures differ from the constants", what);
    };
    auto plastic_c = std::make_shared<PlasticMaterial>(SpectrumTexture(0.3f), Spectrum(0.5f, 0.4f, 0.3f), 0.2f);
    auto plastic_t = std::make_shared<PlasticMaterial>(SpectrumTexture(0.3f), Spectrum(0.0f), 0.9f);
    plastic_t->with_texture(FTN_PARAM_KS, SpectrumTexture(Spectrum(0.5f, 0.4f, 0.3f))).with_texture(FTN_PARAM_UROUGHNESS, SpectrumTexture(0.2f));
    same(probe(plastic_c, 0.5, 0.5), probe(plastic_t, 0.5, 0.5), "plastic");
    auto glass_c = std::make_shared<GlassMaterial>(Spectrum(1.0f), Spectrum(0.8f, 0.9f, 1.0f), 1.4f, 0.3f, 0.3f);
    auto glass_t = std::make_shared<GlassMaterial>(Spectrum(1.0f), Spectrum(0.0f), 2.0f, 0.3f, 0.3f);
    glass_t->with_texture(FTN_PARAM_KT, SpectrumTexture(Spectrum(0.8f, 0.9f, 1.0f))).with_texture(FTN_PARAM_INDEX, SpectrumTexture(1.4f));
    same(probe(glass_c, 0.5, 0.5), probe(glass_t, 0.5, 0.5), "glass");
    // Kd checkerboard through the table: under the direct-lighting integrator the two cells give Kd / pi * (L cos + pi L_env)
    auto matte_t = std::make_shared<MatteMaterial>(SpectrumTexture(0.0f));
    matte_t->with_texture(FTN_PARAM_KD, SpectrumTexture::checkerboard(Spectrum(0.8f, 0.2f, 0.2f), Spectrum(0.1f, 0.1f, 0.9f)));
    auto matte_i = std::make_shared<MatteMaterial>(SpectrumTexture::checkerboard(Spectrum(0.8f, 0.2f, 0.2f), Spectrum(0.1f, 0.1f, 0.9f)));
    same(probe(matte_i, 0.5, 0.5), probe(matte_t, 0.5, 0.5), "matte Kd, tex1 cell");
    same(probe(matte_i, -0.5, 0.5), probe(matte_t, -0.5, 0.5), "matte Kd, tex2 cell");
    (void)pi;
}

// texture/image.rs through MIPMap::from_image; no reference test renders one -- closed forms as in tests/test_oracle_render.py:
// (1) a ramp image is reproduced by the bilinear filter at level 0, (2) with a huge uscale the footprint exceeds the
// image and the 1x1 top level (the mean of a two-valued checker) is returned
static void image_texture_closed_forms() {
    const float pi = 3.14159265358979f;
    const int n = 16;
    std::vector<float> ramp((size_t)3 * n * n), checker((size_t)3 * n * n);
    for (int t = 0; t < n; ++t) for (int s = 0; s < n; ++s) {
        float* px = &ramp[3 * ((size_t)t * n + s)];
        px[0] = (s + 0.5f) / n; px[1] = (t + 0.5f) / n; px[2] = 0.25f;
        float* cx = &checker[3 * ((size_t)t * n + s)];
        cx[0] = cx[1] = cx[2] = ((s + t) % 2) ? 0.8f : 0.2f;
    }
    auto probe = [&](SpectrumTexture tex, double x, double y) {
        std::vector<float> v = {-6, -6, 0, 6, -6, 0, 6, 6, 0, -6, 6, 0}, uv = {-6, -6, 6, -6, 6, 6, -6, 6};
        auto mesh = std::make_shared<TriangleMesh>(Transform::identity(), std::vector<uint32_t>{0, 1, 2, 0, 2, 3}, v, std::vector<float>{}, uv);
        std::vector<GeometricPrimitive> prims; prims.emplace_back(mesh, std::make_shared<MatteMaterial>(tex));
        Scene scene(g_lib, prims, {Light(DistantLight::from_to(Point3f(0, 0, 1), Point3f(0, 0, 0), Spectrum(3.0f)))});
        PerspectiveCamera camera(Transform::look_at({x, y, 30.0}, {x, y, 0.0}, {0, 1, 0}).inverse(), 5, 5, 0.5f);
        Film film(g_lib, 5, 5);
        SamplerIntegrator<DirectLightingIntegrator> integrator(camera, DirectLightingIntegrator(LightStrategy::UniformSampleOne, 1));
        integrator.render_parallel(scene, film, RandomSampler::new_with_seed(2, 0));
        Spectrum mean(0.0f);
        const auto px = film.into_spectrum_buffer().first;
        for (const Spectrum& s : px) { mean.r += s.r / px.size(); mean.g += s.g / px.size(); mean.b += s.b / px.size(); }
        return mean;
    };
    auto mp = MIPMap::from_image(ramp, n, n, FTN_WRAP_CLAMP);
    CHECK(mp->n_levels == 5 && mp->packed.size() == (size_t)3 * (256 + 64 + 16 + 4 + 1), "pyramid of a 16x16 image: %d levels", mp->n_levels);
    const Spectrum a = probe(SpectrumTexture::image(mp, UVMapping{1.0f / 12.0f, 1.0f / 12.0f, 0.5f, 0.5f}), 2.2, -3.1);
    CHECK(std::fabs(a.r - (2.2f / 12 + 0.5f) / pi * 3.0f) < 2e-3f && std::fabs(a.g - (-3.1f / 12 + 0.5f) / pi * 3.0f) < 2e-3f && std::fabs(a.b - 0.25f / pi * 3.0f) < 1e-5f,
          "ramp %.7g %.7g %.7g", a.r, a.g, a.b);
    const Spectrum b = probe(SpectrumTexture::image(MIPMap::from_image(checker, n, n, FTN_WRAP_REPEAT), UVMapping{500.0f, 500.0f, 0.0f, 0.0f}), 0.7, -0.4);
    CHECK(std::fabs(b.r - 0.5f / pi * 3.0f) < 1e-4f, "top level of the checker pyramid %.7g", b.r);
    bool threw = false;
    try {
        auto bad = MIPMap::from_image(ramp, n, n); bad->n_levels = 4;
        probe(SpectrumTexture::image(bad), 0.0, 0.0);
    } catch (const Error& e) { threw = e.code == FTN_ERR_INVALID_ARGUMENT; }
    CHECK(threw, "a pyramid with the wrong level count must be FTN_ERR_INVALID_ARGUMENT");
}

static void invalid_arguments_are_errors() {
    bool threw = false;
    try {
        std::vector<uint32_t> idx = {0, 1, 7};               // index out of range
        auto mesh = std::make_shared<TriangleMesh>(Transform::identity(), idx, std::vector<float>{0, 0, 0, 1, 0, 0, 0, 1, 0});
        std::vector<GeometricPrimitive> prims;
        prims.emplace_back(mesh, nullptr);
        Scene scene(g_lib, prims);
    } catch (const Error& e) { threw = e.code == FTN_ERR_INVALID_ARGUMENT; }
    CHECK(threw, "out-of-range vertex index must be FTN_ERR_INVALID_ARGUMENT");
    threw = false;
    try {
        TriangleMesh bad(Transform::identity(), {0, 1}, std::vector<float>{0, 0, 0});
    } catch (const Error&) { threw = true; }
    CHECK(threw, "index count not a multiple of 3 (triangle.rs:38 assert)");
}

int main(int argc, char** argv) {
    if (argc < 4) {
        std::fprintf(stderr, "usage: %s <library.so> <prefix> <rounded_cube.ply> [tests...]\n", argv[0]);
        return 2;
    }
    g_ply = argv[3];
    try {
        g_lib = Library::open(argv[1], argv[2]);
    } catch (const Error& e) {
        std::fprintf(stderr, "%s\n", e.what());
        return 3;
    }
    const std::vector<std::pair<std::string, std::function<void()>>> tests = {
        {"furnace_test_path", furnace_test_path},
        {"furnace_test_path_no_rr", furnace_test_path_no_rr},
        {"furnace_test_directlighting", furnace_test_directlighting},
        {"uniform_sample_all_is_unimplemented", uniform_sample_all_is_unimplemented},
        {"test_rounded_cube", test_rounded_cube},
        {"camera_raster_to_camera_fov", camera_raster_to_camera_fov},
        {"world_bound_and_morton_order", world_bound_and_morton_order},
        {"textured_floor_and_mirror_closed_forms", textured_floor_and_mirror_closed_forms},
        {"image_texture_closed_forms", image_texture_closed_forms},
        {"texture_table_parameters", texture_table_parameters},
        {"invalid_arguments_are_errors", invalid_arguments_are_errors},
    };
    int ran = 0;
    for (const auto& [name, fn] : tests) {
        bool wanted = argc == 4;
        for (int i = 4; i < argc; ++i) wanted |= name == argv[i];
        if (!wanted) continue;
        int before = g_failures;
        try {
            fn();
        } catch (const std::exception& e) {
            std::fprintf(stderr, "  exception: %s\n", e.what());
            ++g_failures;
        }
        std::printf("test %s ... %s\n", name.c_str(), g_failures == before ? "ok" : "FAILED");
        ++ran;
    }
    std::printf("%d run, %d failed\n", ran, g_failures);
    return g_failures ? 1 : (ran ? 0 : 2);
}
