// TEST INFRASTRUCTURE ONLY.  C ABI of the CPU oracle (libfountain_oracle.so): the same
// structs and call shapes as include/fountain_gpu.h with the prefix `orc_`, so parity tests
// drive both backends with the same inputs.  It is a checker / CPU baseline; the product
// never loads it.
//
// Parity status: the reference is Rust and cannot be built or run here (no rustc/cargo).
// The oracle is pinned against every known-answer test the reference's own test-suite holds
// for this path (tests/test_oracle_kat.py lists them with file:line).  The RNG stream
// (rand/rand_xoshiro, un-vendored) and cgmath ulp-level behaviour are restated from their
// published sources and are "parity unpinned" -- no reference test pins them.  The round-1 widenings
// (point / distant lights, mirror, checkerboard / uv textures) have no reference test either: their
// restatements are pinned against closed forms only (tests/test_oracle_render.py).
#include "../include/fountain_gpu.h"
#include "ref_render.h"
#include <chrono>
#include <string>
#include <numeric>

using namespace ref;

namespace {
thread_local std::string g_err;
int fail(int code, const std::string& msg) { g_err = msg; return code; }
int g_threads = 0;

struct OrcScene {
    Scene scene;
    bool built = false;
    double build_seconds = 0.0;
    uint32_t n_tris = 0;
};

Transform make_transform(const float* t, const float* inv) {
    Transform T; T.t = mat_from_flat(t); T.invt = mat_from_flat(inv); return T;
}
}  // namespace

#define ORC_API extern "C" __attribute__((visibility("default")))

ORC_API uint32_t orc_abi_version(void) { return FTN_ABI_VERSION; }
ORC_API const char* orc_last_error(void) { return g_err.c_str(); }
ORC_API int orc_set_threads(int n) { g_threads = n; return 0; }
ORC_API int orc_hardware_threads(void) { return (int)std::max(1u, std::thread::hardware_concurrency()); }

// MIPMap from the ABI's concatenated pyramid (level l: max(1, w >> l) x max(1, h >> l), mipmap.rs:107-121)
static std::shared_ptr<MIPMap> make_mipmap(const float* data, int w, int h, int levels, int wrap) {
    if (!data || w < 1 || h < 1 || wrap < 0 || wrap > 2) return nullptr;
    int expect = 1; for (int m = std::max(w, h); m > 1; m >>= 1) ++expect;   // 1 + log2_usize(max(w, h))
    if (levels != expect || levels > FTN_MAX_MIP_LEVELS) return nullptr;
    auto mp = std::make_shared<MIPMap>();
    mp->wrap = wrap;
    size_t off = 0;
    for (int l = 0; l < levels; ++l) {
        int lw = std::max(1, w >> l), lh = std::max(1, h >> l);
        std::vector<Spectrum> lv((size_t)lw * lh);
        for (size_t k = 0; k < lv.size(); ++k) lv[k] = Spectrum(data[off + 3 * k], data[off + 3 * k + 1], data[off + 3 * k + 2]);
        off += 3 * lv.size();
        mp->w.push_back(lw); mp->h.push_back(lh); mp->pyramid.push_back(std::move(lv));
    }
    return mp;
}

ORC_API int orc_scene_create(const FtnSceneDesc* d, OrcScene** out) {
    if (!d || !out) return fail(FTN_ERR_INVALID_ARGUMENT, "null argument");
    if (d->abi_version != FTN_ABI_VERSION) return fail(FTN_ERR_INVALID_ARGUMENT, "abi version mismatch");
    OrcScene* os = new OrcScene();
    Scene& s = os->scene;
    if (d->n_textures && !d->textures) { delete os; return fail(FTN_ERR_INVALID_ARGUMENT, "n_textures != 0 but textures is null"); }
    s.textures.resize(d->n_textures);
    for (uint32_t i = 0; i < d->n_textures; ++i) {
        const FtnTexture& t = d->textures[i];
        TextureDef& td = s.textures[i];
        if (t.type < FTN_TEXTURE_CONSTANT || t.type > FTN_TEXTURE_IMAGE) { delete os; return fail(FTN_ERR_INVALID_ARGUMENT, "texture table: unknown texture type"); }
        td.type = t.type; td.value = Spectrum(t.value[0], t.value[1], t.value[2]);
        td.tex1 = Spectrum(t.tex1[0], t.tex1[1], t.tex1[2]); td.tex2 = Spectrum(t.tex2[0], t.tex2[1], t.tex2[2]);
        for (int c = 0; c < 2; ++c) { td.uv_scale[c] = t.uv_scale[c]; td.uv_delta[c] = t.uv_delta[c]; }
        if (t.type == FTN_TEXTURE_IMAGE) {
            td.image = make_mipmap(t.image, t.image_width, t.image_height, t.image_levels, t.image_wrap);
            if (!td.image) { delete os; return fail(FTN_ERR_INVALID_ARGUMENT, "texture table: bad pyramid description"); }
        }
    }
    for (uint32_t i = 0; i < d->n_materials; ++i) {
        const FtnMaterial& m = d->materials[i];
        Material mm{};
        mm.type = m.type;
        mm.table = &s.textures;
        for (int p = 0; p < FTN_PARAM_COUNT; ++p) {
            if (m.param_texture[p] > d->n_textures) { delete os; return fail(FTN_ERR_INVALID_ARGUMENT, "material: param_texture id out of range"); }
            mm.ptex[p] = m.param_texture[p];
        }
        if (m.type == FTN_MATERIAL_PLASTIC) mm.ptex[FTN_PARAM_VROUGHNESS] = mm.ptex[FTN_PARAM_UROUGHNESS];
        mm.kd = Spectrum(m.kd[0], m.kd[1], m.kd[2]); mm.ks = Spectrum(m.ks[0], m.ks[1], m.ks[2]);
        mm.eta = Spectrum(m.eta[0], m.eta[1], m.eta[2]); mm.k = Spectrum(m.k[0], m.k[1], m.k[2]);
        mm.u_rough = m.u_roughness; mm.v_rough = m.v_roughness; mm.remap = m.remap_roughness != 0;
        mm.kr = Spectrum(m.kr[0], m.kr[1], m.kr[2]); mm.sigma = m.sigma;
        mm.kt = Spectrum(m.kt[0], m.kt[1], m.kt[2]);
        if (m.type == FTN_MATERIAL_GLASS) {   // glass.rs:64-67: the specular branch is todo!() under the path integrator
            Float ur = m.u_roughness, vr = m.v_roughness;
            if (m.remap_roughness) { ur = roughness_to_alpha(ur); vr = roughness_to_alpha(vr); }
            const bool textured = m.param_texture[FTN_PARAM_UROUGHNESS] || m.param_texture[FTN_PARAM_VROUGHNESS];
            if (!textured && ur == 0.0f && vr == 0.0f) { delete os; return fail(FTN_ERR_UNSUPPORTED, "smooth glass is todo!() in the reference (glass.rs:66)"); }
        }
        mm.kd_texture = m.kd_texture; mm.tex1 = Spectrum(m.tex1[0], m.tex1[1], m.tex1[2]); mm.tex2 = Spectrum(m.tex2[0], m.tex2[1], m.tex2[2]);
        for (int c = 0; c < 2; ++c) { mm.uv_scale[c] = m.uv_scale[c]; mm.uv_delta[c] = m.uv_delta[c]; }
        if (m.kd_texture == FTN_TEXTURE_IMAGE && (m.type == FTN_MATERIAL_MATTE || m.type == FTN_MATERIAL_PLASTIC || m.type == FTN_MATERIAL_MIRROR)) {
            mm.image = make_mipmap(m.image, m.image_width, m.image_height, m.image_levels, m.image_wrap);
            if (!mm.image) { delete os; return fail(FTN_ERR_INVALID_ARGUMENT, "image texture: bad pyramid description"); }
        }
        s.materials.push_back(mm);
    }
    // scene-wide arrays, one TriangleMesh view per FtnMeshDesc
    s.vertices.resize(d->n_vertices);
    for (uint32_t v = 0; v < d->n_vertices; ++v) s.vertices[v] = Point3(d->positions[3 * v], d->positions[3 * v + 1], d->positions[3 * v + 2]);
    if (d->normals) { s.normals.resize(d->n_vertices); for (uint32_t v = 0; v < d->n_vertices; ++v) s.normals[v] = Vec3(d->normals[3 * v], d->normals[3 * v + 1], d->normals[3 * v + 2]); }
    if (d->uvs) s.uvs.assign(d->uvs, d->uvs + 2 * (size_t)d->n_vertices);
    s.indices.assign(d->indices, d->indices + 3 * (size_t)d->n_triangles);
    for (size_t k = 0; k < s.indices.size(); ++k)
        if (s.indices[k] >= d->n_vertices) { delete os; return fail(FTN_ERR_INVALID_ARGUMENT, "vertex index out of range"); }
    s.meshes.resize(d->n_meshes);
    uint32_t covered = 0;
    for (uint32_t mi = 0; mi < d->n_meshes; ++mi) {
        const FtnMeshDesc& md = d->meshes[mi];
        if (md.first_tri != covered) { delete os; return fail(FTN_ERR_INVALID_ARGUMENT, "meshes must tile the index buffer in order"); }
        if (md.material_id >= (int32_t)d->n_materials) { delete os; return fail(FTN_ERR_INVALID_ARGUMENT, "material id out of range"); }
        covered += md.n_tris;
        if (covered > d->n_triangles) { delete os; return fail(FTN_ERR_INVALID_ARGUMENT, "mesh range exceeds the index buffer"); }
        TriangleMesh& tm = s.meshes[mi];
        tm.flip_normals = (md.flags & FTN_MESH_FLIP_NORMALS) != 0;
        tm.vertices = s.vertices.data();
        tm.normals = d->normals ? s.normals.data() : nullptr;
        tm.uvs = d->uvs ? s.uvs.data() : nullptr;
        tm.vertex_indices = s.indices.data() + 3 * (size_t)md.first_tri;
    }
    if (covered != d->n_triangles) { delete os; return fail(FTN_ERR_INVALID_ARGUMENT, "meshes do not cover all triangles"); }
    s.spheres.resize(d->n_spheres);
    s.pending_emit.assign(d->n_triangles + d->n_spheres, Spectrum(0.0f));
    for (uint32_t mi = 0; mi < d->n_meshes; ++mi) {
        const FtnMeshDesc& md = d->meshes[mi];
        for (uint32_t t = 0; t < md.n_tris; ++t) {
            Primitive p{}; p.kind = 0; p.tri.mesh = &s.meshes[mi]; p.tri.tri_id = t; p.sphere = nullptr;
            // `AreaLightSource` in front of the shape: one DiffuseAreaLight per triangle (loaders/pbrt.rs:275-316)
            p.material = md.material_id; p.light = md.emissive ? -2 : -1;
            if (md.emissive) s.pending_emit[s.prims.size()] = Spectrum(md.emit[0], md.emit[1], md.emit[2]);
            s.prims.push_back(p);
        }
    }
    for (uint32_t i = 0; i < d->n_spheres; ++i) {
        const FtnSphere& fs = d->spheres[i];
        Sphere& sp = s.spheres[i];
        sp.object_to_world = make_transform(fs.object_to_world, fs.world_to_object);
        sp.world_to_object = inverse(sp.object_to_world);
        sp.reverse_orientation = fs.reverse_orientation != 0;
        sp.init(fs.radius, fs.z_min, fs.z_max, fs.phi_max_deg);
        Primitive p{}; p.kind = 1; p.sphere = &sp; p.tri.mesh = nullptr; p.tri.tri_id = 0;
        p.material = fs.material_id; p.light = fs.emissive ? -2 : -1;
        s.pending_emit[d->n_triangles + i] = Spectrum(fs.emit[0], fs.emit[1], fs.emit[2]);
        s.prims.push_back(p);
    }
    for (uint32_t i = 0; i < d->n_lights; ++i) {
        const FtnLight& fl = d->lights[i];
        if (fl.type == FTN_LIGHT_POINT || fl.type == FTN_LIGHT_DISTANT) {
            s.lights.emplace_back();
            Light& l = s.lights.back();
            l.type = fl.type == FTN_LIGHT_POINT ? 2 : 3; l.prims = &s.prims; l.prim = -1;
            l.world_point = Point3(fl.point[0], fl.point[1], fl.point[2]);
            l.dir_to_light = Vec3(fl.direction[0], fl.direction[1], fl.direction[2]);
            l.intensity = Spectrum(fl.intensity[0], fl.intensity[1], fl.intensity[2]);
            l.world_center = Point3(0, 0, 0); l.world_radius = 0.0f;
            continue;
        }
        if (fl.type != FTN_LIGHT_INFINITE || fl.width < 1 || fl.height < 1 || !fl.texels) { delete os; return fail(FTN_ERR_INVALID_ARGUMENT, "bad light"); }
        s.lights.emplace_back();
        Light& l = s.lights.back();
        l.type = 0; l.prims = &s.prims; l.prim = -1;
        l.map.w = fl.width; l.map.h = fl.height;
        l.map.texels.assign(fl.texels, fl.texels + 3 * (size_t)fl.width * fl.height);
        l.light_to_world = make_transform(fl.light_to_world, fl.world_to_light);
        l.world_to_light = inverse(l.light_to_world);
        l.world_center = Point3(0, 0, 0); l.world_radius = 0.0f;
        l.compute_distribution();
    }
    os->n_tris = d->n_triangles;
    *out = os;
    return FTN_OK;
}

ORC_API int orc_scene_destroy(OrcScene* s) { delete s; return FTN_OK; }

// Scene::new + BVH::build (scene/mod.rs:32, bvh.rs:27).
ORC_API int orc_bvh_build(OrcScene* os) {
    if (!os) return fail(FTN_ERR_INVALID_ARGUMENT, "null scene");
    if (os->built) return FTN_OK;
    auto t0 = std::chrono::steady_clock::now();
    os->scene.finish();
    os->build_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    os->built = true;
    return FTN_OK;
}

// Test hook: the primitive (insertion index) behind every light of Scene::lights, -1 for lights that are not area
// lights.  The reference lists area lights in ITS BVH's primitive order (scene/mod.rs:40-44); the parity tests use this
// to lay the emissive triangles out so that the GPU's primitive-order light list enumerates them the same way.
ORC_API int orc_debug_light_prims(const OrcScene* os, int32_t* out, uint32_t cap, uint32_t* n_lights) {
    if (!os || !os->built) return fail(FTN_ERR_INVALID_ARGUMENT, "scene not built");
    const std::vector<Light>& ls = os->scene.lights;
    if (n_lights) *n_lights = (uint32_t)ls.size();
    for (uint32_t i = 0; i < cap && i < ls.size(); ++i) out[i] = ls[i].type == 1 ? (int32_t)ls[i].prim : -1;
    return FTN_OK;
}

// The GPU LBVH's sort key, stated on the CPU: morton3 (morton.rs:3-36) of each TRIANGLE's
// bounds centroid (bounds.rs:161-163) normalised with Bounds3f::offset (bounds.rs:200-206)
// in the bounds of all centroids.  offset can be exactly 1.0, which morton.rs's debug_assert
// excludes, so the fixed-point value is clamped to 1023.  order = stable sort by code.
ORC_API int orc_bvh_debug_morton(const OrcScene* os, uint32_t* codes, uint32_t* order) {
    if (!os) return fail(FTN_ERR_INVALID_ARGUMENT, "null scene");
    const Scene& s = os->scene;
    uint32_t n = os->n_tris;
    std::vector<Point3> cen(n);
    Bounds3 cb = Bounds3::empty();
    for (uint32_t i = 0; i < n; ++i) { cen[i] = s.prims[i].world_bound().centroid(); cb = cb.join_point(cen[i]); }
    std::vector<uint32_t> c(n);
    for (uint32_t i = 0; i < n; ++i) {
        Vec3 o = cb.offset(cen[i]);
        uint32_t fx = std::min(to_fixed_point(o.x), 1023u), fy = std::min(to_fixed_point(o.y), 1023u), fz = std::min(to_fixed_point(o.z), 1023u);
        c[i] = (expand_bits(fx) << 2) | (expand_bits(fy) << 1) | expand_bits(fz);
    }
    if (codes) std::copy(c.begin(), c.end(), codes);
    if (order) {
        std::vector<uint32_t> idx(n);
        std::iota(idx.begin(), idx.end(), 0u);
        std::stable_sort(idx.begin(), idx.end(), [&](uint32_t a, uint32_t b) { return c[a] < c[b]; });
        std::copy(idx.begin(), idx.end(), order);
    }
    return FTN_OK;
}

ORC_API int orc_scene_world_bound(const OrcScene* os, float out[6]) {
    if (!os || !os->built) return fail(FTN_ERR_INVALID_ARGUMENT, "scene not built");
    const Bounds3& b = os->scene.bvh.bounds;
    out[0] = b.min.x; out[1] = b.min.y; out[2] = b.min.z; out[3] = b.max.x; out[4] = b.max.y; out[5] = b.max.z;
    return FTN_OK;
}

ORC_API int orc_scene_stats(const OrcScene* os, FtnStats* st) {
    if (!os || !st) return fail(FTN_ERR_INVALID_ARGUMENT, "null argument");
    std::memset(st, 0, sizeof(*st));
    st->bvh_build_seconds = os->build_seconds;
    st->bvh_nodes = (uint32_t)os->scene.bvh.nodes.size();
    st->bvh_node_bytes = 32;   // bvh.rs:269 "Should be 32 bytes"
    st->bvh_tri_bytes = 48;    // 12 B indices + 36 B vertices gathered per test
    return FTN_OK;
}

namespace {
template <class F> void parallel_for(size_t n, F f) {
    int nt = g_threads > 0 ? g_threads : (int)std::max(1u, std::thread::hardware_concurrency());
    if (n < 4096) nt = 1;
    std::atomic<size_t> next{0};
    const size_t chunk = 1024;
    auto worker = [&]() {
        for (;;) { size_t b = next.fetch_add(chunk); if (b >= n) break; size_t e = std::min(n, b + chunk); for (size_t i = b; i < e; ++i) f(i); }
    };
    std::vector<std::thread> th;
    for (int i = 1; i < nt; ++i) th.emplace_back(worker);
    worker();
    for (auto& t : th) t.join();
}
Ray to_ray(const FtnRay& r) { Ray q; q.origin = Point3(r.o[0], r.o[1], r.o[2]); q.dir = Vec3(r.d[0], r.d[1], r.d[2]); q.t_max = r.t_max; q.time = r.time; return q; }
}  // namespace

// Scene::intersect over a batch (scene/mod.rs:51 -> bvh.rs:160).
ORC_API int orc_intersect(const OrcScene* os, size_t n, const FtnRay* rays, FtnHit* hits) {
    if (!os || !os->built) return fail(FTN_ERR_INVALID_ARGUMENT, "scene not built");
    parallel_for(n, [&](size_t i) {
        Ray r = to_ray(rays[i]); SurfaceInteraction si;
        if (os->scene.intersect(&r, &si)) { hits[i].prim = (uint32_t)si.prim; hits[i].t = r.t_max; hits[i].b1 = si.b[1]; hits[i].b2 = si.b[2]; }
        else { hits[i].prim = FTN_NO_HIT; hits[i].t = rays[i].t_max; hits[i].b1 = 0.0f; hits[i].b2 = 0.0f; }
    });
    return FTN_OK;
}
// As orc_intersect, plus the reference structure's node visits and primitive tests
// (counters[0] += nodes, counters[1] += prims): the "reference-structure" bytes/ray figure.
ORC_API int orc_intersect_count(const OrcScene* os, size_t n, const FtnRay* rays, FtnHit* hits, uint64_t* counters) {
    if (!os || !os->built) return fail(FTN_ERR_INVALID_ARGUMENT, "scene not built");
    TraversalCounters tc;
    for (size_t i = 0; i < n; ++i) {
        Ray r = to_ray(rays[i]); SurfaceInteraction si;
        if (os->scene.intersect(&r, &si, &tc)) { hits[i].prim = (uint32_t)si.prim; hits[i].t = r.t_max; hits[i].b1 = si.b[1]; hits[i].b2 = si.b[2]; }
        else { hits[i].prim = FTN_NO_HIT; hits[i].t = rays[i].t_max; hits[i].b1 = 0.0f; hits[i].b2 = 0.0f; }
    }
    counters[0] += tc.nodes; counters[1] += tc.prims;
    return FTN_OK;
}
// Scene::intersect_test over a batch (scene/mod.rs:55 -> bvh.rs:217).
ORC_API int orc_intersect_test(const OrcScene* os, size_t n, const FtnRay* rays, uint8_t* out) {
    if (!os || !os->built) return fail(FTN_ERR_INVALID_ARGUMENT, "scene not built");
    parallel_for(n, [&](size_t i) { out[i] = os->scene.intersect_test(to_ray(rays[i])) ? 1 : 0; });
    return FTN_OK;
}
// Brute force over the primitive list, as bvh.rs:446-458's test helpers.
ORC_API int orc_intersect_brute(const OrcScene* os, size_t n, const FtnRay* rays, FtnHit* hits) {
    if (!os) return fail(FTN_ERR_INVALID_ARGUMENT, "null scene");
    const Scene& s = os->scene;
    parallel_for(n, [&](size_t i) {
        Ray r = to_ray(rays[i]);
        hits[i].prim = FTN_NO_HIT; hits[i].t = rays[i].t_max; hits[i].b1 = hits[i].b2 = 0.0f;
        for (size_t p = 0; p < s.prims.size(); ++p) {
            Float t; SurfaceInteraction si;
            if (s.prims[p].shape_intersect(r, &t, &si)) { r.t_max = t; hits[i].prim = (uint32_t)p; hits[i].t = t; hits[i].b1 = si.b[1]; hits[i].b2 = si.b[2]; }
        }
    });
    return FTN_OK;
}

ORC_API int orc_film_pixel_count(const FtnFilm* f, int32_t* w, int32_t* h) {
    if (!f) return fail(FTN_ERR_INVALID_ARGUMENT, "null film");
    Film film; film.init(f->x_resolution, f->y_resolution, f->crop_window, f->filter_radius);
    if (w) *w = film.width();
    if (h) *h = film.height();
    return FTN_OK;
}

// SamplerIntegrator::render_parallel (integrator/mod.rs:218) + film.pixels (film.rs:24).
ORC_API int orc_render(const OrcScene* os, const FtnCamera* cam, const FtnFilm* f, const FtnSampler* smp,
                       const FtnIntegrator* integ, FtnPixel* out, FtnStats* stats) {
    if (!os || !cam || !f || !smp || !integ || !out) return fail(FTN_ERR_INVALID_ARGUMENT, "null argument");
    if (!os->built) return fail(FTN_ERR_INVALID_ARGUMENT, "scene not built");
    if (smp->sample_stride < 1 || smp->sample_begin < 0) return fail(FTN_ERR_INVALID_ARGUMENT, "bad sample shard");
    if (smp->mode == FTN_SAMPLER_REFERENCE_TILE_STREAM && (smp->sample_begin != 0 || smp->sample_stride != 1))
        return fail(FTN_ERR_INVALID_ARGUMENT, "the reference tile stream cannot be sharded");
    Camera c;
    c.camera_to_world = mat_from_flat(cam->camera_to_world); c.raster_to_camera = mat_from_flat(cam->raster_to_camera);
    c.lens_radius = cam->lens_radius; c.focal_distance = cam->focal_distance;
    c.shutter_open = cam->shutter_open; c.shutter_close = cam->shutter_close;
    Film film; film.init(f->x_resolution, f->y_resolution, f->crop_window, f->filter_radius);
    RenderParams rp{};
    rp.integrator = integ->type; rp.max_depth = integ->max_depth; rp.rr_threshold = integ->rr_threshold;
    rp.spp = smp->samples_per_pixel; rp.seed = smp->seed; rp.sampler_mode = smp->mode;
    rp.sample_begin = smp->sample_begin; rp.sample_stride = smp->sample_stride;
    rp.threads = g_threads; rp.count_traversal = stats && (stats->flags & FTN_STATS_COUNT_TRAVERSAL);
    Counters ctr;
    auto t0 = std::chrono::steady_clock::now();
    RenderResult rr = render(os->scene, c, film, rp, &ctr);
    double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    std::memcpy(out, film.pixels.data(), film.pixels.size() * sizeof(float));
    if (stats) {
        std::memset(stats, 0, sizeof(*stats));
        stats->camera_samples = ctr.camera_samples; stats->rays_closest = ctr.rays_closest; stats->rays_any = ctr.rays_any;
        stats->node_visits = ctr.nodes; stats->tri_tests = ctr.prim_tests;
        stats->device_seconds = secs; stats->bvh_build_seconds = os->build_seconds;
        stats->bvh_nodes = (uint32_t)os->scene.bvh.nodes.size(); stats->bvh_node_bytes = 32; stats->bvh_tri_bytes = 48;
    }
    if (rr.nan_radiance) return fail(FTN_ERR_NAN_RADIANCE, "NaN radiance value (check_radiance, integrator/mod.rs:285)");
    if (rr.unsupported) return fail(FTN_ERR_UNSUPPORTED, "reference hits unimplemented!() on this input");
    return FTN_OK;
}

ORC_API int orc_film_to_rgb(size_t n, const FtnPixel* px, float* rgb) {
    for (size_t i = 0; i < n; ++i) pixel_to_rgb(px[i].xyz, rgb + 3 * i);
    return FTN_OK;
}

// ---- known-answer hooks: one per unit test the reference holds for this path --------------------
ORC_API uint32_t orc_kat_morton3(float x, float y, float z) { return morton3(x, y, z); }
ORC_API uint32_t orc_kat_expand_bits(uint32_t v) { return expand_bits(v); }
ORC_API uint32_t orc_kat_to_fixed_point(float v) { return to_fixed_point(v); }
ORC_API int orc_kat_sign_differs(float a, float b, float c) { return sign_differs(a, b, c) ? 1 : 0; }
ORC_API float orc_kat_gamma(int n) { return gamma(n); }
ORC_API float orc_kat_next_float_up(float v) { return next_float_up(v); }
ORC_API float orc_kat_next_float_down(float v) { return next_float_down(v); }
ORC_API int orc_kat_bounds_intersect(const float bmin[3], const float bmax[3], const FtnRay* ray, float out_t[2]) {
    Bounds3 b; b.min = Point3(bmin[0], bmin[1], bmin[2]); b.max = Point3(bmax[0], bmax[1], bmax[2]);
    return b.intersect_test(to_ray(*ray), &out_t[0], &out_t[1]) ? 1 : 0;
}
ORC_API float orc_kat_fresnel_dielectric(float cos_i, float eta_i, float eta_t) { return fresnel_dielectric(cos_i, eta_i, eta_t); }
ORC_API void orc_kat_fresnel_conductor(float cos_i, const float eta[3], const float k[3], float out[3]) {
    Spectrum r = fresnel_conductor(std::fabs(cos_i), Spectrum(1.0f), Spectrum(eta[0], eta[1], eta[2]), Spectrum(k[0], k[1], k[2]));
    out[0] = r.c[0]; out[1] = r.c[1]; out[2] = r.c[2];
}
ORC_API void orc_kat_distribution1d_sample(const float* func, int n, float u, float* x, float* pdf, int* idx) {
    Distribution1D d; d.init(func, (size_t)n); size_t i; d.sample_continuous(u, x, pdf, &i); *idx = (int)i;
}
ORC_API void orc_kat_concentric_sample_disk(float u0, float u1, float out[2]) { concentric_sample_disk(u0, u1, &out[0], &out[1]); }
ORC_API void orc_kat_offset_ray_origin(const float p[3], const float e[3], const float n[3], const float d[3], float out[3]) {
    Point3 r = offset_ray_origin(Point3(p[0], p[1], p[2]), Vec3(e[0], e[1], e[2]), Vec3(n[0], n[1], n[2]), Vec3(d[0], d[1], d[2]));
    out[0] = r.x; out[1] = r.y; out[2] = r.z;
}
ORC_API int orc_kat_triangle_intersect(const float p0[3], const float p1[3], const float p2[3], const FtnRay* ray, float out_tb[4]) {
    TriHitCore h;
    if (!triangle_intersect_core(Point3(p0[0], p0[1], p0[2]), Point3(p1[0], p1[1], p1[2]), Point3(p2[0], p2[1], p2[2]), to_ray(*ray), &h)) return 0;
    out_tb[0] = h.t; out_tb[1] = h.b0; out_tb[2] = h.b1; out_tb[3] = h.b2;
    return 1;
}
// Sphere::intersect against one sphere: returns hit, t, world p, p_err, n (13 floats).
ORC_API int orc_kat_sphere_intersect(const FtnSphere* fs, const FtnRay* ray, float out[13]) {
    Sphere sp; sp.object_to_world = make_transform(fs->object_to_world, fs->world_to_object);
    sp.world_to_object = inverse(sp.object_to_world); sp.reverse_orientation = fs->reverse_orientation != 0;
    sp.init(fs->radius, fs->z_min, fs->z_max, fs->phi_max_deg);
    Float t; SurfaceInteraction si;
    if (!sp.intersect(to_ray(*ray), &t, &si)) return 0;
    out[0] = t; out[1] = si.hit.p.x; out[2] = si.hit.p.y; out[3] = si.hit.p.z;
    out[4] = si.hit.p_err.x; out[5] = si.hit.p_err.y; out[6] = si.hit.p_err.z;
    out[7] = si.hit.n.x; out[8] = si.hit.n.y; out[9] = si.hit.n.z;
    out[10] = si.wo.x; out[11] = si.wo.y; out[12] = si.wo.z;
    return 1;
}
// Camera::generate_ray for one sample (camera/mod.rs:117-143).
ORC_API void orc_kat_camera_ray(const FtnCamera* cam, float fx, float fy, float lx, float ly, float tu, FtnRay* out) {
    Camera c; c.camera_to_world = mat_from_flat(cam->camera_to_world); c.raster_to_camera = mat_from_flat(cam->raster_to_camera);
    c.lens_radius = cam->lens_radius; c.focal_distance = cam->focal_distance; c.shutter_open = cam->shutter_open; c.shutter_close = cam->shutter_close;
    Ray r = c.generate_ray(fx, fy, lx, ly, tu);
    out->o[0] = r.origin.x; out->o[1] = r.origin.y; out->o[2] = r.origin.z;
    out->d[0] = r.dir.x; out->d[1] = r.dir.y; out->d[2] = r.dir.z; out->t_max = r.t_max; out->time = r.time;
}
// Material -> Bsdf evaluation at a canonical frame (ns = ng = +z, dpdu = +x): f, pdf for (wo, wi),
// and one sample_f(wo, u).  out: f[3], pdf, ok, sf[3], swi[3], spdf  (12 floats).
ORC_API void orc_kat_bsdf(const FtnMaterial* m, const float wo[3], const float wi[3], const float u[2], float out[12]) {
    Material mm{}; mm.type = m->type;
    mm.kd = Spectrum(m->kd[0], m->kd[1], m->kd[2]); mm.ks = Spectrum(m->ks[0], m->ks[1], m->ks[2]);
    mm.eta = Spectrum(m->eta[0], m->eta[1], m->eta[2]); mm.k = Spectrum(m->k[0], m->k[1], m->k[2]);
    mm.u_rough = m->u_roughness; mm.v_rough = m->v_roughness; mm.remap = m->remap_roughness != 0;
    mm.kr = Spectrum(m->kr[0], m->kr[1], m->kr[2]); mm.kd_texture = 0; mm.sigma = m->sigma;
    mm.kt = Spectrum(m->kt[0], m->kt[1], m->kt[2]);
    SurfaceInteraction si{};
    si.hit.n = Vec3(0, 0, 1); si.shading_n = Vec3(0, 0, 1); si.shading_dpdu = Vec3(1, 0, 0); si.dpdu = Vec3(1, 0, 0);
    Bsdf b; compute_scattering_functions(mm, si, &b);
    Vec3 o(wo[0], wo[1], wo[2]), i(wi[0], wi[1], wi[2]);
    Spectrum f = b.f(o, i, BXDF_ALL);
    out[0] = f.c[0]; out[1] = f.c[1]; out[2] = f.c[2]; out[3] = b.pdf(o, i, BXDF_ALL);
    ScatterSample s;
    bool ok = b.sample_f(o, u[0], u[1], BXDF_ALL, &s);
    out[4] = ok ? 1.0f : 0.0f;
    if (ok) { out[5] = s.f.c[0]; out[6] = s.f.c[1]; out[7] = s.f.c[2]; out[8] = s.wi.x; out[9] = s.wi.y; out[10] = s.wi.z; out[11] = s.pdf; }
    else for (int k = 5; k < 12; ++k) out[k] = 0.0f;
}
// InfiniteAreaLight sample / pdf / Le for light 0 of a scene. out: wi[3], pdf, L[3], pdf_of(wi), Le(wi)[3]
ORC_API int orc_kat_env(const OrcScene* os, const float u[2], float out[11]) {
    if (!os || os->scene.lights.empty() || os->scene.lights[0].type != 0) return fail(FTN_ERR_INVALID_ARGUMENT, "no infinite light");
    const Light& l = os->scene.lights[0];
    SurfaceHit ref{}; ref.p = Point3(0, 0, 0); ref.n = Vec3(0, 0, 1);
    LiSample ls; bool unsup = false;
    if (!l.sample_incident_radiance(ref, u[0], u[1], &ls, &unsup)) return FTN_ERR_UNSUPPORTED;
    out[0] = ls.wi.x; out[1] = ls.wi.y; out[2] = ls.wi.z; out[3] = ls.pdf;
    out[4] = ls.radiance.c[0]; out[5] = ls.radiance.c[1]; out[6] = ls.radiance.c[2];
    out[7] = l.pdf_incident_radiance(ref, ls.wi);
    Ray r; r.origin = Point3(0, 0, 0); r.dir = ls.wi; r.t_max = INF; r.time = 0;
    Spectrum le = l.environment_emitted_radiance(r);
    out[8] = le.c[0]; out[9] = le.c[1]; out[10] = le.c[2];
    return FTN_OK;
}
// MIPMap::lookup_trilinear_width on a pyramid in the ABI's layout (mipmap.rs test_mipmap_lookup :364-382)
ORC_API int orc_kat_mipmap_lookup(const float* pyramid, int w, int h, int levels, int wrap, float s, float t, float width, float out[3]) {
    auto mp = make_mipmap(pyramid, w, h, levels, wrap);
    if (!mp) return FTN_ERR_INVALID_ARGUMENT;
    Spectrum v = mp->lookup_trilinear_width(s, t, width);
    out[0] = v.c[0]; out[1] = v.c[1]; out[2] = v.c[2];
    return FTN_OK;
}
ORC_API float orc_kat_counter_uniform(uint64_t seed, uint64_t sample_index, uint32_t dim) { return counter_uniform(seed, sample_index, dim); }
// First n floats of the reference tile stream for a seed (xoshiro256+ via SplitMix64).
ORC_API void orc_kat_reference_stream(uint64_t seed, int n, float* out) {
    Sampler s{}; s.mode = 1; s.seed_reference(seed);
    for (int i = 0; i < n; ++i) out[i] = s.get_1d();
}
