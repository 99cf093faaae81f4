// TEST INFRASTRUCTURE ONLY (see ref_math.h).  Lights, Scene, the path / direct-lighting
// integrators, thin-lens camera, samplers, box-filter film and the 16x16-tile render loop.
#pragma once
#include "ref_shading.h"
#include <thread>
#include <atomic>
#include <mutex>

namespace ref {

// ---- samplers -------------------------------------------------------------------------------
// u32 -> f32 as rand 0.6.5 `Standard` for f32 (absent; Cargo.lock:1121): (u >> 8) * 2^-24.
inline Float u32_to_unit_float(uint32_t u) { return (Float)(u >> 8) * (1.0f / 16777216.0f); }

// Counter-based stream shared with the GPU (FTN_SAMPLER_COUNTER): SplitMix64 finaliser of
// a 64-bit key (seed, pixel-sample index, dimension).  Defined by this project, not by the
// reference; see DESIGN.md "Sampler".
inline uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
inline Float counter_uniform(uint64_t seed, uint64_t sample_index, uint32_t dim) {
    uint64_t key = mix64(seed * 0x9E3779B97F4A7C15ull + 0xD1B54A32D192ED03ull) ^ (sample_index * 0x9E3779B97F4A7C15ull);
    uint64_t z = mix64(key + (uint64_t)dim * 0xC2B2AE3D27D4EB4Full);
    return u32_to_unit_float((uint32_t)(z >> 32));
}
enum { DIM_CAMERA = 5, DIM_PER_BOUNCE = 8 };

struct Sampler {
    int mode;            // FtnSamplerMode
    int spp;
    // reference stream: xoshiro256+ seeded by SplitMix64 (rand_xoshiro 0.2.0, absent; Cargo.lock:1227)
    uint64_t s[4];
    // counter stream
    uint64_t seed, sample_index; uint32_t dim;
    int current_sample;

    void seed_reference(uint64_t sd) {   // Xoshiro256Plus::seed_from_u64 -> SplitMix64
        uint64_t x = sd;
        for (int i = 0; i < 4; ++i) {
            x += 0x9E3779B97F4A7C15ull;
            uint64_t z = x;
            z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
            z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
            s[i] = z ^ (z >> 31);
        }
    }
    uint64_t next_u64() {   // xoshiro256+
        uint64_t result = s[0] + s[3];
        uint64_t t = s[1] << 17;
        s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3];
        s[2] ^= t;
        s[3] = (s[3] << 45) | (s[3] >> 19);
        return result;
    }
    Float get_1d() {   // sampler/random.rs:37-39
        if (mode == 1) return u32_to_unit_float((uint32_t)(next_u64() >> 32));
        return counter_uniform(seed, sample_index, dim++);
    }
    void get_2d(Float* a, Float* b) { *a = get_1d(); *b = get_1d(); }   // random.rs:41-43 (x then y)
    // counter mode: jump to a fixed dimension so that the stream layout does not depend on
    // data-dependent consumption (film 0-1, lens 2-3, time 4, then 8 per bounce).
    void set_dim(uint32_t d) { dim = d; }
};

// ---- lights --------------------------------------------------------------------------------
struct LiSample { Spectrum radiance; Vec3 wi; Float pdf; SurfaceHit p0, p1; };

struct Scene;

struct Light {
    int type;   // 0 infinite, 1 diffuse area, 2 point (light/point.rs), 3 distant (light/distant.rs)
    Point3 world_point; Vec3 dir_to_light; Spectrum intensity;   // point: position + I; distant: direction + L
    // infinite (light/infinite.rs)
    EnvMap map; Distribution2D distribution;
    Transform light_to_world, world_to_light;
    Point3 world_center; Float world_radius;
    // diffuse area (light/diffuse.rs)
    Spectrum emit; int prim; Float area;
    const std::vector<Primitive>* prims;

    // infinite.rs:63-77 -- note `(height, width) = resolution()` where resolution is (w, h)
    void compute_distribution() {
        size_t height = (size_t)map.w, width = (size_t)map.h;
        Float filter = 1.0f / (Float)std::max(width, height);
        std::vector<Float> img(width * height, 0.0f);
        for (size_t j = 0; j < height; ++j) {
            Float v = (Float)j / (Float)height;
            Float sin_theta = std::sin(PI * ((Float)j + 0.5f) / (Float)height);
            for (size_t i = 0; i < width; ++i) {
                Float u = (Float)i / (Float)width;
                Float lum = map.lookup_width(u, v, filter).luminance();
                img[i + j * width] = lum * sin_theta;
            }
        }
        distribution.init(img.data(), width, height);
    }
    bool is_delta() const { return type == 2 || type == 3; }   // LightFlags::is_delta_light, light/mod.rs:66-73

    // DiffuseAreaLight::emitted_radiance, diffuse.rs:45-51
    Spectrum area_emitted(const SurfaceHit& hit, Vec3 w) const { return (dot(hit.n, w) > 0.0f) ? emit : Spectrum(0.0f); }

    // Shape::pdf_from_ref, shapes/mod.rs:55-66
    Float shape_pdf_from_ref(const SurfaceHit& reference, Vec3 wi) const {
        Ray ray = reference.spawn_ray(wi);
        Float t; SurfaceInteraction isect;
        if ((*prims)[prim].shape_intersect(ray, &t, &isect)) {
            Vec3 d = reference.p - isect.hit.p;
            return magnitude2(d) / (abs_dot(isect.hit.n, -wi) * (*prims)[prim].area());
        }
        return 0.0f;
    }

    bool sample_incident_radiance(const SurfaceHit& reference, Float u0, Float u1, LiSample* out, bool* unsupported) const {
        if (type == 0) {   // infinite.rs:99-140
            Float uvx, uvy, map_pdf;
            distribution.sample_continuous(u0, u1, &uvx, &uvy, &map_pdf);
            if (map_pdf == 0.0f) { *unsupported = true; return false; }   // unimplemented!() in the reference
            Float theta = uvy * PI, phi = uvx * 2.0f * PI;
            Vec3 wi = transform_vector(light_to_world.t, Vec3(std::sin(theta) * std::cos(phi), std::sin(theta) * std::sin(phi), std::cos(theta)));
            Float pdf = (std::sin(theta) == 0.0f) ? 0.0f : map_pdf / (2.0f * PI * PI * std::sin(theta));
            out->p0 = reference;
            out->p1.p = reference.p + wi * (2.0f * world_radius);
            out->p1.p_err = Vec3(0, 0, 0); out->p1.time = reference.time; out->p1.n = Vec3(0, 0, 0);
            out->radiance = map.lookup_width(uvx, uvy, 0.0f);
            out->wi = wi; out->pdf = pdf;
            return true;
        }
        if (type == 2) {   // point.rs:44-64
            out->wi = normalize(world_point - reference.p);
            out->pdf = 1.0f;
            out->p0 = reference;
            out->p1.p = world_point; out->p1.p_err = Vec3(0, 0, 0); out->p1.time = reference.time; out->p1.n = Vec3(0, 0, 0);
            out->radiance = intensity / magnitude2(world_point - reference.p);
            return true;
        }
        if (type == 3) {   // distant.rs:52-71
            out->p0 = reference;
            out->p1.p = reference.p + dir_to_light * (2.0f * world_radius);
            out->p1.p_err = Vec3(0, 0, 0); out->p1.time = reference.time; out->p1.n = Vec3(0, 0, 0);
            out->radiance = intensity; out->wi = dir_to_light; out->pdf = 1.0f;
            return true;
        }
        // diffuse.rs:74-89
        SurfaceHit p_shape = (*prims)[prim].sample(u0, u1);
        Vec3 wi = normalize(p_shape.p - reference.p);
        out->pdf = shape_pdf_from_ref(reference, wi);
        out->p0 = reference; out->p1 = p_shape;
        out->radiance = area_emitted(p_shape, -wi);
        out->wi = wi;
        return true;
    }
    Float pdf_incident_radiance(const SurfaceHit& reference, Vec3 wi_w) const {
        if (type == 0) {   // infinite.rs:142-154
            Vec3 wi = transform_vector(world_to_light.t, wi_w);
            Float theta = spherical_theta(wi), phi = spherical_phi(wi);
            if (std::sin(theta) == 0.0f) return 0.0f;
            return distribution.pdf(phi * (1.0f / (2.0f * PI)), theta * FRAC_1_PI) / (2.0f * PI * PI * std::sin(theta));
        }
        if (is_delta()) return 0.0f;                  // point.rs:65, distant.rs:73
        return shape_pdf_from_ref(reference, wi_w);   // diffuse.rs:91-93
    }
    Spectrum environment_emitted_radiance(const Ray& ray) const {   // infinite.rs:156-164; light/mod.rs:32 default
        if (type != 0) return Spectrum(0.0f);
        Vec3 w = normalize(transform_vector(world_to_light.t, ray.dir));
        Float s = spherical_phi(w) * (1.0f / (2.0f * PI)), t = spherical_theta(w) * FRAC_1_PI;
        return map.lookup_width(s, t, 0.0f);
    }
};

// ---- Scene, scene/mod.rs ---------------------------------------------------------------------
struct Scene {
    std::vector<Point3> vertices; std::vector<Vec3> normals; std::vector<Float> uvs; std::vector<uint32_t> indices;
    std::vector<TriangleMesh> meshes;
    std::vector<Sphere> spheres;
    std::vector<Material> materials;
    std::vector<TextureDef> textures;   // the table Material::ptex indexes
    std::vector<Primitive> prims;     // insertion order: triangles (mesh order, tri_id), then spheres
    std::vector<Light> lights;
    BVH bvh;
    bool intersect(Ray* ray, SurfaceInteraction* si, TraversalCounters* c = nullptr) const { return bvh.intersect(ray, si, c); }
    bool intersect_test(const Ray& ray, TraversalCounters* c = nullptr) const { return bvh.intersect_test(ray, c); }
    Spectrum environment_emitted_radiance(const Ray& ray) const {   // :58-64
        Spectrum s(0.0f);
        for (const Light& l : lights) s = s + l.environment_emitted_radiance(ray);
        return s;
    }
    Spectrum emitted_radiance(const SurfaceInteraction& si, Vec3 w) const {   // interaction.rs:175-180
        int li = prims[si.prim].light;
        if (li < 0) return Spectrum(0.0f);
        return lights[li].area_emitted(si.hit, w);
    }
    // Scene::new, :32-49: preprocess explicit lights, then append area lights in BVH-permuted order.
    void finish() {
        bvh.build(&prims);
        for (Light& l : lights) if (l.type == 0 || l.type == 3) bvh.bounds.bounding_sphere(&l.world_center, &l.world_radius);   // infinite.rs / distant.rs:46-50
        for (int slot = 0; slot < (int)bvh.prim_order.size(); ++slot) {
            Primitive& p = prims[bvh.prim_order[slot]];
            if (p.light == -2) {   // marked emissive by the builder
                Light l{}; l.type = 1; l.prims = &prims; l.prim = bvh.prim_order[slot];
                l.emit = pending_emit[bvh.prim_order[slot]];
                l.area = p.area();
                lights.push_back(l);
                p.light = (int)lights.size() - 1;
            }
        }
    }
    std::vector<Spectrum> pending_emit;
};

struct Counters {
    std::atomic<uint64_t> rays_closest{0}, rays_any{0}, camera_samples{0}, nodes{0}, prim_tests{0};
};
struct LocalCounters { uint64_t rays_closest = 0, rays_any = 0, camera_samples = 0; TraversalCounters trav; };

// ---- integrator/mod.rs:289-395 -----------------------------------------------------------------
struct RenderCtx { const Scene* scene; Sampler* sampler; LocalCounters* ctr; bool* unsupported; bool count_traversal; };

inline Spectrum estimate_direct(const Bsdf& bsdf, const SurfaceInteraction& isect, Float us0, Float us1,
                                const Light& light, int light_index, Float ul0, Float ul1, RenderCtx& cx) {
    const Scene& scene = *cx.scene;
    TraversalCounters* tc = cx.count_traversal ? &cx.ctr->trav : nullptr;
    int flags = BXDF_ALL & ~BXDF_SPECULAR;
    Spectrum radiance(0.0f);
    LiSample ls;
    if (light.sample_incident_radiance(isect.hit, ul0, ul1, &ls, cx.unsupported)) {
        if (ls.pdf > 0.0f && !ls.radiance.is_black()) {
            Spectrum f = bsdf.f(isect.wo, ls.wi, flags) * abs_dot(ls.wi, isect.shading_n);
            Float scattering_pdf = bsdf.pdf(isect.wo, ls.wi, flags);
            if (!f.is_black()) {
                cx.ctr->rays_any++;
                bool occluded = scene.intersect_test(ls.p0.spawn_ray_to_hit(ls.p1), tc);   // light/mod.rs:82-84
                if (!occluded) {
                    if (light.is_delta()) radiance = radiance + f * ls.radiance / ls.pdf;   // :331-332
                    else {
                        Float weight = power_heuristic(1, ls.pdf, 1, scattering_pdf);
                        radiance = radiance + f * ls.radiance * weight / ls.pdf;
                    }
                }
            }
        }
    }
    if (light.is_delta()) return radiance;   // :343: a delta light cannot be hit by BSDF sampling
    ScatterSample sc;
    if (bsdf.sample_f(isect.wo, us0, us1, flags, &sc)) {
        Spectrum f = sc.f * abs_dot(sc.wi, isect.shading_n);
        bool sampled_specular = (sc.sampled_type & BXDF_SPECULAR) != 0;
        if (f.is_black()) return radiance;
        Float weight;
        if (sampled_specular) weight = 1.0f;
        else {
            Float light_pdf = light.pdf_incident_radiance(isect.hit, sc.wi);
            if (light_pdf == 0.0f) return radiance;
            weight = power_heuristic(1, sc.pdf, 1, light_pdf);
        }
        Ray ray = isect.hit.spawn_ray(sc.wi);
        SurfaceInteraction si2;
        cx.ctr->rays_closest++;
        Spectrum incident;
        if (scene.intersect(&ray, &si2, tc)) {
            // only the SAME light's emission counts (integrator/mod.rs:370-381)
            if (scene.prims[si2.prim].light >= 0 && scene.prims[si2.prim].light == light_index)
                incident = scene.emitted_radiance(si2, -sc.wi);
            else incident = Spectrum(0.0f);
        } else incident = light.environment_emitted_radiance(ray);
        if (!incident.is_black()) radiance = radiance + f * incident * weight / sc.pdf;
    }
    return radiance;
}

inline Spectrum uniform_sample_one_light(const SurfaceInteraction& isect, const Bsdf& bsdf, RenderCtx& cx) {   // :289-305
    size_t n_lights = cx.scene->lights.size();
    if (n_lights == 0) return Spectrum(0.0f);
    Float pick = cx.sampler->get_1d() * (Float)n_lights;
    Float capped = fmin_(pick, (Float)(n_lights - 1));
    size_t light_num = (capped > 0.0f) ? (size_t)capped : 0;
    Float ul0, ul1, us0, us1;
    cx.sampler->get_2d(&ul0, &ul1);
    cx.sampler->get_2d(&us0, &us1);
    return (Float)n_lights * estimate_direct(bsdf, isect, us0, us1, cx.scene->lights[light_num], (int)light_num, ul0, ul1, cx);
}

// geometry/mod.rs:107-133 Differential (the `Option` is `has`)
struct Differential {
    bool has = false;
    Point3 rx_origin, ry_origin; Vec3 rx_dir, ry_dir;
};

// math.rs:56-72
inline bool solve_linear_system_2x2(const Float A[2][2] /* [col][row] */, Float b0, Float b1, Float* x0, Float* x1) {
    Float det = A[0][0] * A[1][1] - A[1][0] * A[0][1];   // cgmath Matrix2::determinant
    if (std::fabs(det) < 1.0e-10f) return false;
    *x0 = (A[1][1] * b0 - A[1][0] * b1) / det;
    *x1 = (A[0][0] * b1 - A[0][1] * b0) / det;
    return !(std::isnan(*x0) || std::isnan(*x1));
}

// SurfaceInteraction::compute_tex_differentials, interaction.rs:124-176; None -> all zero (:117)
inline void compute_tex_differentials(SurfaceInteraction* si, const Differential& diff) {
    si->dudx = si->dvdx = si->dudy = si->dvdy = 0.0f;
    if (!diff.has) return;
    Vec3 n = si->hit.n;
    Float d = dot(n, si->hit.p);
    Float tx = -(dot(n, diff.rx_origin) - d) / dot(n, diff.rx_dir);
    Point3 px = diff.rx_origin + tx * diff.rx_dir;
    Float ty = -(dot(n, diff.ry_origin) - d) / dot(n, diff.ry_dir);
    Point3 py = diff.ry_origin + ty * diff.ry_dir;
    Vec3 dpdx = px - si->hit.p, dpdy = py - si->hit.p;
    si->dpdx = Vec3(0.0f, 0.0f, 0.0f); si->dpdy = Vec3(0.0f, 0.0f, 0.0f);   // unwrap_or_default (:117) when a solve below fails
    int d0, d1;
    if (std::fabs(n.x) > std::fabs(n.y) && std::fabs(n.x) > std::fabs(n.z)) { d0 = 1; d1 = 2; }
    else if (std::fabs(n.y) > std::fabs(n.z)) { d0 = 0; d1 = 2; }
    else { d0 = 0; d1 = 1; }
    Float A[2][2] = {{si->dpdu[d0], si->dpdu[d1]}, {si->dpdv[d0], si->dpdv[d1]}};   // from_cols(dpdu, dpdv)
    Float dudx, dvdx, dudy, dvdy;
    if (!solve_linear_system_2x2(A, dpdx[d0], dpdx[d1], &dudx, &dvdx)) return;
    if (!solve_linear_system_2x2(A, dpdy[d0], dpdy[d1], &dudy, &dvdy)) return;
    si->dudx = dudx; si->dvdx = dvdx; si->dudy = dudy; si->dvdy = dvdy;
    si->dpdx = dpdx; si->dpdy = dpdy;
}

// SurfaceInteraction::compute_scattering_functions, interaction.rs:111-121
inline bool compute_bsdf(const Scene& scene, SurfaceInteraction& si, const Differential& diff, Bsdf* bsdf, bool* unsupported = nullptr) {
    compute_tex_differentials(&si, diff);
    int m = scene.prims[si.prim].material;
    if (m < 0) return false;
    if (!compute_scattering_functions(scene.materials[m], si, bsdf) && unsupported) *unsupported = true;   // todo!() in the reference (glass.rs:66)
    return true;
}

// integrator/path.rs:25-95
// The spawned rays keep the CAMERA ray's differential unchanged (path.rs:73,79 pass `ray.diff` on), so every hit of
// a path computes its texture differentials from the camera's two offset rays.
inline Spectrum path_incident_radiance(Ray ray, const Differential& diff, int max_depth, Float rr_threshold, RenderCtx& cx) {
    const Scene& scene = *cx.scene;
    TraversalCounters* tc = cx.count_traversal ? &cx.ctr->trav : nullptr;
    Spectrum L(0.0f), beta(1.0f);
    int bounces = 0;
    bool specular_bounce = false;
    for (;;) {
        SurfaceInteraction si;
        cx.ctr->rays_closest++;
        bool hit = scene.intersect(&ray, &si, tc);
        if (bounces == 0 || specular_bounce) {
            if (hit) L = L + beta * scene.emitted_radiance(si, -ray.dir);
            else L = L + beta * scene.environment_emitted_radiance(ray);
        }
        if (!hit || bounces >= max_depth) break;
        Bsdf bsdf;
        if (compute_bsdf(scene, si, diff, &bsdf, cx.unsupported)) {
            if (cx.sampler->mode == 0) cx.sampler->set_dim(DIM_CAMERA + DIM_PER_BOUNCE * (uint32_t)bounces);
            if (bsdf.num_components(BXDF_ALL & ~BXDF_SPECULAR) > 0) {
                Spectrum direct = beta * uniform_sample_one_light(si, bsdf, cx);
                L = L + direct;
            }
            if (cx.sampler->mode == 0) cx.sampler->set_dim(DIM_CAMERA + DIM_PER_BOUNCE * (uint32_t)bounces + 5);
            Vec3 wo = -ray.dir;
            Float u0, u1; cx.sampler->get_2d(&u0, &u1);
            ScatterSample s;
            if (bsdf.sample_f(wo, u0, u1, BXDF_ALL, &s) && !s.f.is_black()) {
                beta = beta * (s.f * abs_dot(s.wi, si.shading_n) / s.pdf);
                specular_bounce = (s.sampled_type & BXDF_SPECULAR) != 0;
                ray = si.hit.spawn_ray(s.wi);
            } else break;
        } else {
            ray = si.hit.spawn_ray(ray.dir);   // null BSDF: skip without counting a bounce (:76-80)
            continue;
        }
        if (beta.max_component_value() < rr_threshold && bounces > 3) {
            Float q = fmax_(0.05f, 1.0f - beta.max_component_value());
            if (cx.sampler->get_1d() < q) break;
            beta = beta / (1.0f - q);
        }
        bounces += 1;
    }
    return L;
}

// integrator/direct_lighting.rs:50-106 with LightStrategy::UniformSampleOne, and the specular
// recursion of integrator/mod.rs:40-178: specular_reflect / specular_transmit sample the BSDF with
// (REFLECTION | SPECULAR) / (TRANSMISSION | SPECULAR) and recurse with depth + 1.  sampler.get_2d()
// is evaluated as an argument even when no such lobe exists (integrator/mod.rs:53,113).
// specular_reflect derives the differentials of the mirrored ray from the incoming ones, the hit's texture differentials and
// the shading geometry's dndu / dndv (integrator/mod.rs:59-83), so an image texture seen through a mirror is filtered with
// the reflected footprint.
inline Spectrum direct_incident_radiance(Ray ray, const Differential& diff, int max_depth, int depth, RenderCtx& cx, bool* unsupported_null) {
    const Scene& scene = *cx.scene;
    TraversalCounters* tc = cx.count_traversal ? &cx.ctr->trav : nullptr;
    SurfaceInteraction si;
    cx.ctr->rays_closest++;
    if (!scene.intersect(&ray, &si, tc)) return scene.environment_emitted_radiance(ray);
    Bsdf bsdf;
    Spectrum radiance(0.0f);
    if (!compute_bsdf(scene, si, diff, &bsdf, cx.unsupported)) { *unsupported_null = true; return radiance; }   // unimplemented!() :98
    radiance = radiance + scene.emitted_radiance(si, si.wo);
    if (cx.sampler->mode == 0) cx.sampler->set_dim(DIM_CAMERA + DIM_PER_BOUNCE * (uint32_t)depth);
    radiance = radiance + uniform_sample_one_light(si, bsdf, cx);
    if (depth + 1 < max_depth) {
        if (cx.sampler->mode == 0) cx.sampler->set_dim(DIM_CAMERA + DIM_PER_BOUNCE * (uint32_t)depth + 5);
        Float a, b;
        cx.sampler->get_2d(&a, &b);   // specular_reflect, integrator/mod.rs:40-103
        ScatterSample s;
        if (bsdf.sample_f(si.wo, a, b, BXDF_REFLECTION | BXDF_SPECULAR, &s) && abs_dot(s.wi, si.shading_n) != 0.0f) {
            Differential nd;   // `ray.diff.map(...)`, integrator/mod.rs:59-83
            if (diff.has) {
                nd.has = true;
                nd.rx_origin = si.hit.p + si.dpdx; nd.ry_origin = si.hit.p + si.dpdy;
                Vec3 ns = si.shading_n, wo = si.wo;
                Vec3 dndx = si.shading_dndu * si.dudx + si.shading_dndv * si.dvdx;
                Vec3 dndy = si.shading_dndu * si.dudy + si.shading_dndv * si.dvdy;
                Vec3 dwo_dx = -diff.rx_dir - wo, dwo_dy = -diff.ry_dir - wo;
                Float dDN_dx = dot(dwo_dx, ns) + dot(wo, dndx), dDN_dy = dot(dwo_dy, ns) + dot(wo, dndy);
                nd.rx_dir = (s.wi - dwo_dx) + 2.0f * dot(wo, ns) * dndx + dDN_dx * ns;
                nd.ry_dir = (s.wi - dwo_dy) + 2.0f * dot(wo, ns) * dndy + dDN_dy * ns;
            }
            Spectrum li = direct_incident_radiance(si.hit.spawn_ray(s.wi), nd, max_depth, depth + 1, cx, unsupported_null);
            radiance = radiance + s.f * li * std::fabs(dot(s.wi, si.shading_n)) / s.pdf;
        }
        if (cx.sampler->mode == 0) cx.sampler->set_dim(DIM_CAMERA + DIM_PER_BOUNCE * (uint32_t)depth + 7);
        cx.sampler->get_2d(&a, &b);   // specular_transmit's argument: no transmissive specular lobe is in scope
    }
    return radiance;
}

// ---- camera/mod.rs ------------------------------------------------------------------------------
struct Camera {
    Mat4 camera_to_world, raster_to_camera;
    Float lens_radius, focal_distance, shutter_open, shutter_close;
    // generate_ray :117-143 (= the main ray of generate_ray_differential :145-205)
    Ray generate_ray(Float fx, Float fy, Float lx, Float ly, Float time_u) const {
        Point3 p_camera = transform_point(raster_to_camera, Point3(fx, fy, 0.0f));
        Float time = (1.0f - time_u) * shutter_open + time_u * shutter_close;   // Float::lerp math.rs:21-23
        Ray ray; ray.origin = Point3(0, 0, 0); ray.dir = normalize(p_camera - Point3(0, 0, 0)); ray.time = time; ray.t_max = INF;
        if (lens_radius > 0.0f) {
            Float dx, dy; concentric_sample_disk(lx, ly, &dx, &dy);
            Float plx = lens_radius * dx, ply = lens_radius * dy;
            Float ft = focal_distance / ray.dir.z;
            Point3 p_focus = ray.at(ft);
            ray.origin = Point3(plx, ply, 0.0f);
            ray.dir = normalize(p_focus - ray.origin);
        }
        return ray_transform(camera_to_world, ray);
    }
    // The two offset rays of generate_ray_differential :145-205, transformed (transform.rs:324-338: plain point /
    // vector transforms) and scaled about the main ray (geometry/mod.rs:125-132, integrator/mod.rs:249-251).
    // With a lens BOTH offsets use dx_camera (:178 repeats :172), as in the reference.
    Differential generate_differential(const Ray& main_ray, Float fx, Float fy, Float lx, Float ly, Float scale) const {
        Point3 p_camera = transform_point(raster_to_camera, Point3(fx, fy, 0.0f));
        Point3 o0 = transform_point(raster_to_camera, Point3(0, 0, 0));
        Vec3 dx_camera = transform_point(raster_to_camera, Point3(1, 0, 0)) - o0;   // :101-102
        Vec3 dy_camera = transform_point(raster_to_camera, Point3(0, 1, 0)) - o0;
        Differential df; df.has = true;
        if (lens_radius > 0.0f) {
            Float dx, dy; concentric_sample_disk(lx, ly, &dx, &dy);
            Float plx = lens_radius * dx, ply = lens_radius * dy;
            Vec3 ddx = normalize(p_camera + dx_camera);
            Float ft = focal_distance / ddx.z;
            Point3 p_focus = Point3(0, 0, 0) + (ft * ddx);
            df.rx_origin = Point3(plx, ply, 0.0f); df.rx_dir = normalize(p_focus - df.rx_origin);
            Vec3 ddy = normalize(p_camera + dx_camera);
            ft = focal_distance / ddy.z;
            p_focus = Point3(0, 0, 0) + (ft * ddy);
            df.ry_origin = Point3(plx, ply, 0.0f); df.ry_dir = normalize(p_focus - df.ry_origin);
        } else {
            df.rx_origin = Point3(0, 0, 0); df.ry_origin = Point3(0, 0, 0);
            df.rx_dir = normalize(p_camera + dx_camera); df.ry_dir = normalize(p_camera + dy_camera);
        }
        df.rx_origin = transform_point(camera_to_world, df.rx_origin); df.ry_origin = transform_point(camera_to_world, df.ry_origin);
        df.rx_dir = transform_vector(camera_to_world, df.rx_dir); df.ry_dir = transform_vector(camera_to_world, df.ry_dir);
        df.rx_origin = main_ray.origin + (df.rx_origin - main_ray.origin) * scale;
        df.ry_origin = main_ray.origin + (df.ry_origin - main_ray.origin) * scale;
        df.rx_dir = main_ray.dir + (df.rx_dir - main_ray.dir) * scale;
        df.ry_dir = main_ray.dir + (df.ry_dir - main_ray.dir) * scale;
        return df;
    }
};

// ---- film.rs ----------------------------------------------------------------------------------------
struct Film {
    int xres, yres;
    int crop_min[2], crop_max[2];     // cropped_pixel_bounds :49-58
    Float radius[2], inv_radius[2];
    Float table[16][16];
    std::vector<Float> pixels;        // 4 per pixel: xyz + weight (film.rs:12-16)
    std::mutex mtx;
    void init(int xr, int yr, const Float crop[4], const Float rad[2]) {
        xres = xr; yres = yr;
        crop_min[0] = (int)std::ceil((Float)xr * crop[0]); crop_min[1] = (int)std::ceil((Float)yr * crop[2]);
        crop_max[0] = (int)std::ceil((Float)xr * crop[1]); crop_max[1] = (int)std::ceil((Float)yr * crop[3]);
        radius[0] = rad[0]; radius[1] = rad[1]; inv_radius[0] = 1.0f / rad[0]; inv_radius[1] = 1.0f / rad[1];
        for (int y = 0; y < 16; ++y) for (int x = 0; x < 16; ++x) table[y][x] = 1.0f;   // BoxFilter::evaluate
        pixels.assign((size_t)4 * width() * height(), 0.0f);
    }
    int width() const { return crop_max[0] - crop_min[0]; }
    int height() const { return crop_max[1] - crop_min[1]; }
    void sample_bounds(int mn[2], int mx[2]) const {   // :86-93
        mn[0] = (int)std::floor((Float)crop_min[0] + 0.5f - radius[0]);
        mn[1] = (int)std::floor((Float)crop_min[1] + 0.5f - radius[1]);
        mx[0] = (int)std::ceil((Float)crop_max[0] - 0.5f + radius[0]);
        mx[1] = (int)std::ceil((Float)crop_max[1] - 0.5f + radius[1]);
    }
};
struct FilmTile {
    int mn[2], mx[2];
    std::vector<Float> px;   // rgb contrib sum + weight sum
    int w() const { return mx[0] - mn[0]; }
};
inline void get_film_tile(const Film& film, const int smn[2], const int smx[2], FilmTile* tile) {   // :95-113
    int p0x = (int)std::ceil((Float)smn[0] - 0.5f - film.radius[0]);
    int p0y = (int)std::ceil((Float)smn[1] - 0.5f - film.radius[1]);
    int p1x = (int)std::ceil((Float)smx[0] - 0.5f + film.radius[0] + 1.0f);
    int p1y = (int)std::ceil((Float)smx[1] - 0.5f - film.radius[1] + 1.0f);   // `- radius` as written (:100)
    tile->mn[0] = std::max(p0x, film.crop_min[0]); tile->mn[1] = std::max(p0y, film.crop_min[1]);
    tile->mx[0] = std::min(p1x, film.crop_max[0]); tile->mx[1] = std::min(p1y, film.crop_max[1]);
    long area = (long)(tile->mx[0] - tile->mn[0]) * (long)(tile->mx[1] - tile->mn[1]);
    tile->px.assign((size_t)4 * std::max(area, 0L), 0.0f);
}
inline void add_sample_to_tile(const Film& film, FilmTile* tile, Float fx, Float fy, Spectrum radiance, Float sample_weight) {   // :136-172
    Float dx = fx - 0.5f, dy = fy - 0.5f;
    int p0x = (int)std::ceil(dx - film.radius[0]), p0y = (int)std::ceil(dy - film.radius[1]);
    int p1x = (int)std::floor(dx + film.radius[0]) + 1, p1y = (int)std::floor(dy + film.radius[1]) + 1;
    p0x = std::max(p0x, tile->mn[0]); p0y = std::max(p0y, tile->mn[1]);
    p1x = std::min(p1x, tile->mx[0]); p1y = std::min(p1y, tile->mx[1]);
    for (int y = p0y; y < p1y; ++y) {
        Float filt_y = std::fabs(((Float)y - dy) * film.inv_radius[1] * 16.0f);
        int yi = std::min((int)std::floor(filt_y), 15);
        for (int x = p0x; x < p1x; ++x) {
            Float filt_x = std::fabs(((Float)x - dx) * film.inv_radius[0] * 16.0f);
            int xi = std::min((int)std::floor(filt_x), 15);
            Float wgt = film.table[yi][xi];
            Float* px = &tile->px[4 * ((size_t)(y - tile->mn[1]) * tile->w() + (x - tile->mn[0]))];
            Spectrum c = radiance * sample_weight * wgt;
            px[0] += c.c[0]; px[1] += c.c[1]; px[2] += c.c[2];
            px[3] += wgt;
        }
    }
}
inline void merge_film_tile(Film& film, const FilmTile& tile) {   // :121-132
    std::lock_guard<std::mutex> g(film.mtx);
    for (int y = tile.mn[1]; y < tile.mx[1]; ++y) for (int x = tile.mn[0]; x < tile.mx[0]; ++x) {
        const Float* tp = &tile.px[4 * ((size_t)(y - tile.mn[1]) * tile.w() + (x - tile.mn[0]))];
        Float xyz[3]; rgb_to_xyz(tp, xyz);
        Float* fp = &film.pixels[4 * ((size_t)(y - film.crop_min[1]) * film.width() + (x - film.crop_min[0]))];
        fp[0] += xyz[0]; fp[1] += xyz[1]; fp[2] += xyz[2];
        fp[3] += tp[3];
    }
}
inline void pixel_to_rgb(const Float px[4], Float rgb[3]) {   // into_spectrum_buffer :195-210
    xyz_to_rgb(px, rgb);
    if (px[3] != 0.0f) {
        Float inv = 1.0f / px[3];
        for (int i = 0; i < 3; ++i) rgb[i] = fmax_(0.0f, rgb[i] * inv);
    }
}

// ---- SamplerIntegrator::render_parallel, integrator/mod.rs:182-283 -------------------------------
struct RenderParams {
    int integrator;      // 0 path, 1 direct lighting
    int max_depth; Float rr_threshold;
    int spp; uint64_t seed; int sampler_mode;
    int sample_begin, sample_stride;
    int threads;
    bool count_traversal;
};
struct RenderResult { bool nan_radiance = false, unsupported = false; };

inline void render_tile(const Scene& scene, const Camera& cam, Film& film, const RenderParams& rp,
                        const int tmn[2], const int tmx[2], const int sbmx_x, LocalCounters* ctr, RenderResult* res) {
    Sampler sampler{}; sampler.mode = rp.sampler_mode; sampler.spp = rp.spp; sampler.seed = rp.seed;
    uint64_t tile_id = (uint64_t)((int64_t)tmn[1] * (int64_t)sbmx_x + (int64_t)tmn[0]);   // :182-185
    if (rp.sampler_mode == 1) sampler.seed_reference(tile_id);                              // clone_with_seed, random.rs:61-67
    FilmTile tile; get_film_tile(film, tmn, tmx, &tile);
    bool unsupported = false, null_unsupported = false;
    RenderCtx cx{&scene, &sampler, ctr, &unsupported, rp.count_traversal};
    for (int y = tmn[1]; y < tmx[1]; ++y) for (int x = tmn[0]; x < tmx[0]; ++x) {   // iter_points: x fastest
        for (int s = 0; s < rp.spp; ++s) {
            if (rp.sampler_mode == 0) {
                if (s < rp.sample_begin || ((s - rp.sample_begin) % rp.sample_stride) != 0) continue;
                // global pixel-sample index: ((y * xres) + x) * spp + s in RASTER coordinates
                sampler.sample_index = ((uint64_t)((int64_t)y * film.xres + x)) * (uint64_t)rp.spp + (uint64_t)s;
                sampler.dim = 0;
            }
            Float jx, jy, lx, ly, tu;
            sampler.get_2d(&jx, &jy);           // get_camera_sample, sampler/mod.rs:43-51
            Float fx = (Float)x + jx, fy = (Float)y + jy;
            sampler.get_2d(&lx, &ly);
            tu = sampler.get_1d();
            Ray ray = cam.generate_ray(fx, fy, lx, ly, tu);
            Float ray_weight = 1.0f;
            Spectrum L(0.0f);
            ctr->camera_samples++;
            if (ray_weight > 0.0f) {
                // scale_differentials(1 / sqrt(samples_per_pixel)), integrator/mod.rs:249-251
                Differential diff = cam.generate_differential(ray, fx, fy, lx, ly, 1.0f / std::sqrt((Float)rp.spp));
                if (rp.integrator == 0) L = path_incident_radiance(ray, diff, rp.max_depth, rp.rr_threshold, cx);
                else L = direct_incident_radiance(ray, diff, rp.max_depth, 0, cx, &null_unsupported);
                if (L.has_nans()) res->nan_radiance = true;   // check_radiance panics, :285-287
            }
            add_sample_to_tile(film, &tile, fx, fy, L, ray_weight);
        }
    }
    if (unsupported || null_unsupported) res->unsupported = true;
    merge_film_tile(film, tile);
}

inline RenderResult render(const Scene& scene, const Camera& cam, Film& film, const RenderParams& rp, Counters* counters) {
    int smn[2], smx[2]; film.sample_bounds(smn, smx);
    struct T { int mn[2], mx[2]; };
    std::vector<T> tiles;
    for (int y = smn[1]; y < smx[1]; y += 16) for (int x = smn[0]; x < smx[0]; x += 16) {   // bounds.rs:85-97
        T t; t.mn[0] = x; t.mn[1] = y; t.mx[0] = std::min(x + 16, smx[0]); t.mx[1] = std::min(y + 16, smx[1]);
        tiles.push_back(t);
    }
    int nthreads = rp.threads > 0 ? rp.threads : (int)std::max(1u, std::thread::hardware_concurrency());
    nthreads = std::min<int>(nthreads, (int)std::max<size_t>(1, tiles.size()));
    std::atomic<size_t> next{0};
    std::vector<RenderResult> results(nthreads);
    std::vector<LocalCounters> lcs(nthreads);
    auto worker = [&](int ti) {
        for (;;) {
            size_t i = next.fetch_add(1);
            if (i >= tiles.size()) break;
            render_tile(scene, cam, film, rp, tiles[i].mn, tiles[i].mx, smx[0], &lcs[ti], &results[ti]);
        }
    };
    std::vector<std::thread> th;
    for (int i = 1; i < nthreads; ++i) th.emplace_back(worker, i);
    worker(0);
    for (auto& t : th) t.join();
    RenderResult out;
    for (int i = 0; i < nthreads; ++i) {
        out.nan_radiance |= results[i].nan_radiance; out.unsupported |= results[i].unsupported;
        if (counters) {
            counters->rays_closest += lcs[i].rays_closest; counters->rays_any += lcs[i].rays_any;
            counters->camera_samples += lcs[i].camera_samples;
            counters->nodes += lcs[i].trav.nodes; counters->prim_tests += lcs[i].trav.prims;
        }
    }
    return out;
}

}  // namespace ref
