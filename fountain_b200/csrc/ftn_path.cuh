// Per-path logic of the wavefront path tracer as FTN_HD functions; the kernels in render.cu are
// thin wrappers (load state, call, store state, push queues).  Restates
// PathIntegrator::incident_radiance (integrator/path.rs:25-95), uniform_sample_one_light /
// estimate_direct (integrator/mod.rs:289-395), DirectLightingIntegrator (direct_lighting.rs:50-106)
// and Film::add_sample_to_tile / merge_film_tile (film.rs:121-172).
#pragma once
#include "ftn_shade.cuh"
#include "ftn_trace.cuh"

namespace ftn {

enum { ERR_NAN = 1u, ERR_UNSUPPORTED = 2u };

FTN_HD void flag_error(uint32_t* err, uint32_t bit) {
#if defined(__CUDA_ARCH__)
    atomicOr(err, bit);
#else
    *err |= bit;
#endif
}

struct FilmGeom {
    int xres, yres;
    int crop_min[2], crop_max[2];      // cropped_pixel_bounds, film.rs:49-58
    int sb_min[2], sb_max[2];          // sample_bounds, film.rs:86-93
    float radius[2], inv_radius[2];
};

struct PassParams {
    FilmGeom film;
    FtnCamera cam;
    uint64_t seed_key;
    int spp;                // samples_per_pixel of the whole render
    int s_first;            // first global sample index of this pass
    int s_stride;           // stride between consecutive samples of this pass
    int s_count;            // samples per pixel in this pass
    uint32_t n_paths;       // sample pixels * s_count
    int integrator;         // FtnIntegratorType
    int max_depth;
    float rr_threshold;
};

#define FTN_STATE_BOUNCES 0xFFFFu
#define FTN_STATE_SPECULAR 0x10000u
#define FTN_STATE_HAS_DIFF 0x20000u   /* direct-lighting integrator: the path carries the ray differential specular_reflect derived */

// path id -> counter-sampler key.  path = sample_pixel * s_count + local sample; the global
// pixel-sample index is ((y * xres) + x) * spp + s in RASTER coordinates (same in the oracle).
FTN_HD uint64_t path_sample_key(const PassParams& pp, uint32_t path, int* x_out, int* y_out) {
    const int sbw = pp.film.sb_max[0] - pp.film.sb_min[0];
    const uint32_t pix = path / (uint32_t)pp.s_count, sl = path % (uint32_t)pp.s_count;
    const int x = pp.film.sb_min[0] + (int)(pix % (uint32_t)sbw), y = pp.film.sb_min[1] + (int)(pix / (uint32_t)sbw);
    const int s = pp.s_first + (int)sl * pp.s_stride;
    const uint64_t sample_index = ((uint64_t)((int64_t)y * pp.film.xres + x)) * (uint64_t)pp.spp + (uint64_t)s;
    if (x_out) *x_out = x;
    if (y_out) *y_out = y;
    return sampler_sample_key(pp.seed_key, sample_index);
}

// true when the footprint of a film sample (film.rs:138-141) is anything but exactly the pixel
// (x, y) the sample was generated for
FTN_HD bool film_sample_spills(const FilmGeom& f, int x, int y, float fx, float fy) {
    const float dx = rn_sub(fx, 0.5f), dy = rn_sub(fy, 0.5f);
    const int p0x = (int)ceilf(rn_sub(dx, f.radius[0])), p1x = (int)floorf(rn_add(dx, f.radius[0])) + 1;
    const int p0y = (int)ceilf(rn_sub(dy, f.radius[1])), p1y = (int)floorf(rn_add(dy, f.radius[1])) + 1;
    return !(p0x == x && p1x == x + 1 && p0y == y && p1y == y + 1);
}

// Sampler::get_camera_sample (sampler/mod.rs:43-51) + Camera::generate_ray
FTN_HD RayF raygen_path(const PassParams& pp, uint32_t path, float* fx, float* fy, bool* spills) {
    int x, y;
    const uint64_t key = path_sample_key(pp, path, &x, &y);
    const float jx = sampler_uniform(key, 0), jy = sampler_uniform(key, 1);
    const float lx = sampler_uniform(key, 2), ly = sampler_uniform(key, 3), tu = sampler_uniform(key, 4);
    *fx = rn_add((float)x, jx); *fy = rn_add((float)y, jy);
    *spills = film_sample_spills(pp.film, x, y, *fx, *fy);
    return camera_ray(pp.cam, *fx, *fy, lx, ly, tu);
}

// Re-derive the surface at the hit the extend stage found: the same deterministic test on the
// same (ray, primitive) pair reproduces t and the barycentrics bit for bit, so the queues carry a
// 4-byte slot instead of a fat hit record (the reference builds a full SurfaceInteraction per
// accepted candidate, triangle.rs:270-392).
FTN_HD bool surface_at_hit(const SceneView& sc, uint32_t slot, const RayF& ray, Surface* s) {
    if (slot & FTN_SPHERE_SLOT_FLAG) {
        const SphereData& sd = sc.spheres[slot & ~FTN_SPHERE_SLOT_FLAG];
        SphereHit sh;
        if (!sphere_intersect(sd, ray, &sh)) return false;
        sphere_surface(sd, sh, s);
        return true;
    }
    F4 a, b, c; load_tri(sc.bvh, slot, &a, &b, &c);
    TriHit th;
    const RayShear sh = make_ray_shear(ray.d);
    if (!triangle_intersect(V3(a.x, a.y, a.z), V3(b.x, b.y, b.z), V3(c.x, c.y, c.z), ray.o, sh, ray.t_max, &th)) return false;
    triangle_surface(sc, slot, th, ray.d, s);
    return true;
}

FTN_HD int hit_material(const SceneView& sc, uint32_t slot) {
    if (slot & FTN_SPHERE_SLOT_FLAG) return sc.spheres[slot & ~FTN_SPHERE_SLOT_FLAG].material;
    return sc.meshes[f2u(ld4(sc.bvh.tris + (size_t)FTN_TRI_F4 * (size_t)slot + 1).w)].material;
}

FTN_HD V3 area_emitted(const LightData& l, V3 n, V3 w) {   // DiffuseAreaLight::emitted_radiance, diffuse.rs:45-51
    return (x_dot(n, w) > 0.0f) ? V3(l.emit[0], l.emit[1], l.emit[2]) : v3s(0.0f);
}

// Scene::environment_emitted_radiance, scene/mod.rs:58-64
FTN_HD V3 scene_env_radiance(const SceneView& sc, V3 dir) {
    V3 le = v3s(0.0f);
    for (uint32_t l = 0; l < sc.n_lights; ++l) if (sc.lights[l].type == 0) le = le + env_emitted(sc.lights[l].env, dir);
    return le;
}

// ---- uniform_sample_one_light + estimate_direct (integrator/mod.rs:289-395) ------------------------------
// The two rays estimate_direct spawns.  A SINK receives each of them as soon as it is known: the default one (DirectOut, used by
// the host harness of the tests) just keeps them; the kernels' sink writes them straight into the path's shadow / MIS slots, so
// their 18 floats are not live across the continuation sampling that follows (the shaders run at a 128-register cap).
struct DirectOut {
    bool has_shadow; V3 sh_o, sh_d, sh_L;
    bool has_mis; V3 mis_o, mis_d, mis_w; int mis_light;
    FTN_HD void reset() { has_shadow = false; has_mis = false; }
    FTN_HD void emitted(V3) {}                                            // the caller reads ShadeOut::L
    FTN_HD void shadow(V3 o, V3 d, V3 L) { has_shadow = true; sh_o = o; sh_d = d; sh_L = L; }
    FTN_HD void mis(V3 o, V3 d, V3 w, int light) { has_mis = true; mis_o = o; mis_d = d; mis_w = w; mis_light = light; }
};

template <int MAT, class Sink>
FTN_HD void sample_direct(const SceneView& sc, const Surface& s, const Bsdf& bsdf, V3 scale,
                          uint64_t key, uint32_t dim0, Sink* out, uint32_t* err) {
    out->reset();
    const uint32_t n_lights = sc.n_lights;
    if (n_lights == 0u) return;
    const float pick = sampler_uniform(key, dim0) * (float)n_lights;
    const float capped = fminf(pick, (float)(n_lights - 1u));
    const uint32_t li = (capped > 0.0f) ? (uint32_t)capped : 0u;
    const float ul0 = sampler_uniform(key, dim0 + 1), ul1 = sampler_uniform(key, dim0 + 2);
    const float us0 = sampler_uniform(key, dim0 + 3), us1 = sampler_uniform(key, dim0 + 4);
    const LightData& light = sc.lights[li];
    const float nl = (float)n_lights;
    const int flags = BXDF_ALL & ~BXDF_SPECULAR;
    // --- light sample ---
    V3 wi = v3s(0.0f), Li = v3s(0.0f); float pdf = 0.0f; V3 p1 = v3s(0.0f), p1_err = v3s(0.0f), p1_n = v3s(0.0f);
    bool ok = true;
    if (light.type == 0) {
        if (!env_sample(light.env, ul0, ul1, &wi, &pdf, &Li)) { flag_error(err, ERR_UNSUPPORTED); ok = false; }
        else {
            const float two_r = rn_mul(2.0f, light.env.world_radius);
            p1 = x_add(s.p, x_scale(wi, two_r));   // infinite.rs:121-129: far endpoint, n = 0, p_err = 0
        }
    } else if (light.type == 2) {   // point.rs:44-64
        const V3 lp = V3(light.vec[0], light.vec[1], light.vec[2]);
        const V3 d = x_sub(lp, s.p);
        wi = x_normalize(d);
        pdf = 1.0f;
        Li = V3(light.emit[0], light.emit[1], light.emit[2]) / x_dot(d, d);
        p1 = lp;
    } else if (light.type == 3) {   // distant.rs:52-71
        wi = V3(light.vec[0], light.vec[1], light.vec[2]);
        pdf = 1.0f;
        Li = V3(light.emit[0], light.emit[1], light.emit[2]);
        p1 = x_add(s.p, x_scale(wi, rn_mul(2.0f, light.env.world_radius)));
    } else if (light.type == FTN_LIGHT_TYPE_TRIANGLE) {   // diffuse.rs:74-89 over a triangle of an emissive mesh
        const TriLightGeom g = tri_light_geom(sc, light);
        const ShapeSample ps = triangle_sample(sc, g, ul0, ul1);
        wi = x_normalize(x_sub(ps.p, s.p));
        pdf = triangle_pdf_from_ref(sc, g, s, wi);
        Li = area_emitted(light, ps.n, x_neg(wi));
        p1 = ps.p; p1_err = ps.p_err; p1_n = ps.n;
    } else {   // diffuse.rs:74-89
        const SphereData& sd = sc.spheres[light.sphere];
        const ShapeSample ps = sphere_sample(sd, ul0, ul1);
        wi = x_normalize(x_sub(ps.p, s.p));
        pdf = sphere_pdf_from_ref(sd, s, wi);
        Li = area_emitted(light, ps.n, x_neg(wi));
        p1 = ps.p; p1_err = ps.p_err; p1_n = ps.n;
    }
    if (ok && pdf > 0.0f && !is_black(Li)) {
        const V3 f = bsdf_f<MAT>(bsdf, s.wo, wi, flags) * abs_dot(wi, s.ns);
        const float spdf = bsdf_pdf<MAT>(bsdf, s.wo, wi, flags);
        if (!is_black(f)) {
            // VisibilityTester -> SurfaceHit::spawn_ray_to_hit, interaction.rs:48-58
            const V3 origin = offset_ray_origin(s.p, s.p_err, s.n, x_sub(p1, s.p));
            const V3 target = offset_ray_origin(p1, p1_err, p1_n, x_sub(origin, p1));
            const float w = (light.type == 2 || light.type == 3) ? 1.0f : power_heuristic1(pdf, spdf);   // delta lights: f * Li / pdf (:331-332)
            out->shadow(origin, x_sub(target, origin), scale * (nl * (f * Li * w / pdf)));
        }
    }
    // --- BSDF sample ---
    if (light.type == 2 || light.type == 3) return;   // a delta light cannot be reached by sampling the BSDF (integrator/mod.rs:343)
    ScatterSample bs;
    if (bsdf_sample_f<MAT>(bsdf, s.wo, us0, us1, flags, &bs)) {
        const V3 f = bs.f * abs_dot(bs.wi, s.ns);
        if (is_black(f)) return;
        float lpdf;
        if (light.type == 0) lpdf = env_pdf(light.env, bs.wi);
        else if (light.type == FTN_LIGHT_TYPE_TRIANGLE) lpdf = triangle_pdf_from_ref(sc, tri_light_geom(sc, light), s, bs.wi);
        else lpdf = sphere_pdf_from_ref(sc.spheres[light.sphere], s, bs.wi);
        if (lpdf == 0.0f) return;
        const float w = power_heuristic1(bs.pdf, lpdf);
        out->mis(spawn_origin(s, bs.wi), bs.wi, scale * (nl * (f * w / bs.pdf)), (int)li);
    }
}

// ---- one shaded bounce ---------------------------------------------------------------------------------------
struct ShadeOut {
    V3 L, beta;
    uint32_t state;
    bool alive; V3 next_o, next_d;
    DirectOut direct;
    bool has_diff; RayDiff diff;   // the differential of the mirrored ray (specular_reflect, integrator/mod.rs:59-83)
};

// The camera ray's differential of this path (camera/mod.rs:145-205 scaled by 1/sqrt(spp), integrator/mod.rs:249-251),
// re-derived from the path's camera sample instead of being carried.
FTN_HD_COLD RayDiff path_camera_differential(const PassParams& pp, uint32_t path) {
    int x, y;
    const uint64_t key = path_sample_key(pp, path, &x, &y);
    const float fx = rn_add((float)x, sampler_uniform(key, 0)), fy = rn_add((float)y, sampler_uniform(key, 1));
    const float lx = sampler_uniform(key, 2), ly = sampler_uniform(key, 3), tu = sampler_uniform(key, 4);
    const RayF main_ray = camera_ray(pp.cam, fx, fy, lx, ly, tu);
    return camera_differential(pp.cam, main_ray, fx, fy, lx, ly, 1.0f / sqrtf((float)pp.spp));
}
// Texture differentials of the hit (interaction.rs:117) from the ray's differential `df`.  The path integrator hands the
// CAMERA ray's differential on to every spawned ray unchanged (path.rs:73,79), so each hit of a path intersects the camera's
// two offset rays with its own tangent plane; the direct-lighting integrator uses the camera's at depth 0 and the mirrored
// one behind a specular reflection.
FTN_HD_COLD TexDiffs hit_tex_differentials(const SceneView& sc, uint32_t slot, const RayF& ray, V3 p, V3 n, const RayDiff& df,
                                           V3* dpdx = nullptr, V3* dpdy = nullptr) {
    V3 dpdu, dpdv;
    if (slot & FTN_SPHERE_SLOT_FLAG) {
        SphereHit sh;
        if (!sphere_intersect(sc.spheres[slot & ~FTN_SPHERE_SLOT_FLAG], ray, &sh)) {
            TexDiffs z; z.dudx = z.dvdx = z.dudy = z.dvdy = 0.0f;
            if (dpdx) { *dpdx = v3s(0.0f); *dpdy = v3s(0.0f); }
            return z;
        }
        dpdu = sh.dpdu; dpdv = sh.dpdv;
    } else {
        triangle_dpduv(sc, slot, &dpdu, &dpdv);
    }
    return tex_differentials(p, n, dpdu, dpdv, df, dpdx, dpdy);
}
// specular_reflect's `ray.diff.map(...)`, integrator/mod.rs:59-83
FTN_HD_COLD RayDiff reflect_differential(const SceneView& sc, uint32_t slot, const Surface& s, const RayDiff& df, const TexDiffs& td,
                                         V3 dpdx, V3 dpdy, V3 wi) {
    V3 dndu, dndv;
    if (slot & FTN_SPHERE_SLOT_FLAG) sphere_dnduv(sc.spheres[slot & ~FTN_SPHERE_SLOT_FLAG], s.p, &dndu, &dndv);
    else triangle_dnduv(sc, slot, &dndu, &dndv);
    const V3 dndx = dndu * td.dudx + dndv * td.dvdx, dndy = dndu * td.dudy + dndv * td.dvdy;
    const V3 dwo_dx = -df.rx_d - s.wo, dwo_dy = -df.ry_d - s.wo;
    const float ddn_dx = dot(dwo_dx, s.ns) + dot(s.wo, dndx), ddn_dy = dot(dwo_dy, s.ns) + dot(s.wo, dndy);
    const float two_won = 2.0f * dot(s.wo, s.ns);
    RayDiff o;
    o.rx_o = s.p + dpdx; o.ry_o = s.p + dpdy;
    o.rx_d = (wi - dwo_dx) + dndx * two_won + s.ns * ddn_dx;
    o.ry_d = (wi - dwo_dy) + dndy * two_won + s.ns * ddn_dy;
    return o;
}

// `ray` is the ray that produced `slot`; state/beta/L are the path's values on entry.
// MAT = the material class of the queue this path sits in (FtnMaterialType), or -1 for the
// null-BSDF queue.
// `carried`: the differential a specular reflection left on the path (direct-lighting integrator, FTN_STATE_HAS_DIFF), or null.
// `direct`: the sink of the two rays of estimate_direct (see DirectOut); the convenience overload below keeps them in out->direct.
template <int MAT, bool IMG, class Sink>
FTN_HD void shade_surface_to(const SceneView& sc, const PassParams& pp, uint32_t path, const RayF& ray, uint32_t slot,
                             uint32_t state, V3 beta, V3 L, ShadeOut* out, Sink* direct, uint32_t* err, const RayDiff* carried = nullptr) {
    out->L = L; out->beta = beta; out->state = state; out->alive = false;
    direct->reset();
    out->has_diff = false;
    Surface s;
    if (!surface_at_hit(sc, slot, ray, &s)) return;
    const int bounces = (int)(state & FTN_STATE_BOUNCES);
    const bool direct_only = pp.integrator == FTN_INTEGRATOR_DIRECT_LIGHTING;
    // emitted light at the intersection: path.rs:45-51 / direct_lighting.rs:71
    if (s.light >= 0 && (direct_only || bounces == 0 || (state & FTN_STATE_SPECULAR))) {
        const V3 e = beta * area_emitted(sc.lights[s.light], s.n, direct_only ? s.wo : x_neg(ray.d));
        L = L + e;
        direct->emitted(e);   // a sink that accumulates radiance itself takes it now (out->L is then dead weight it does not read)
    }
    out->L = L;
    if (!direct_only && bounces >= pp.max_depth) return;   // path.rs:54
    if (MAT < 0) {
        // null BSDF: respawn in the same direction without counting a bounce (path.rs:76-80);
        // unimplemented!() under the direct-lighting integrator (direct_lighting.rs:98)
        if (direct_only) { flag_error(err, ERR_UNSUPPORTED); return; }
        out->alive = true; out->next_o = spawn_origin(s, ray.d); out->next_d = ray.d;
        return;
    }
    constexpr int M = MAT < 0 ? 0 : MAT;
    Bsdf bsdf;
    bsdf_init(&bsdf, s.ns, s.n, s.sdpdu);
    TexDiffs td; td.dudx = td.dvdx = td.dudy = td.dvdy = 0.0f;
    // the ray's differential at this hit: the camera's for every hit of the path integrator (path.rs:73,79) and at depth 0 of
    // the direct-lighting integrator, the mirrored one behind its specular reflections (integrator/mod.rs:59-83)
    RayDiff df; df.rx_o = df.rx_d = df.ry_o = df.ry_d = v3s(0.0f);
    V3 dpdx = v3s(0.0f), dpdy = v3s(0.0f);
    bool have_df = false;
    if (IMG) {
        const bool uses_image = sc.materials[s.material].uses_image != 0;
        if (!direct_only) {
            if (uses_image) td = hit_tex_differentials(sc, slot, ray, s.p, s.n, path_camera_differential(pp, path));
        } else {
            const bool chain_goes_on = bounces + 1 < pp.max_depth && LobeKind<M, 0>::value == 2;   // only the mirror has a specular lobe
            if (bounces == 0) { if (uses_image || chain_goes_on) { df = path_camera_differential(pp, path); have_df = true; } }
            else if (carried) { df = *carried; have_df = true; }
            if (have_df && (uses_image || chain_goes_on)) td = hit_tex_differentials(sc, slot, ray, s.p, s.n, df, &dpdx, &dpdy);
        }
    }
    bool unsupported = false;
    material_bsdf<M, IMG>(sc, sc.materials[s.material], s.u, s.v, td, &bsdf, &unsupported);
    if (unsupported) { flag_error(err, ERR_UNSUPPORTED); return; }
    const uint64_t key = path_sample_key(pp, path, nullptr, nullptr);
    // under the direct-lighting integrator `bounces` is the recursion depth of specular_reflect
    const uint32_t dim0 = DIM_CAMERA + (uint32_t)DIM_PER_BOUNCE * (uint32_t)bounces;
    if (direct_only || bsdf_num_components<M>(bsdf, BXDF_ALL & ~BXDF_SPECULAR) > 0)   // path.rs:60 guards; direct_lighting.rs:79 does not
        sample_direct<M>(sc, s, bsdf, beta, key, dim0, direct, err);
    if (direct_only) {
        // specular_reflect (integrator/mod.rs:40-103): follow the mirror direction with depth + 1;
        // the recursion is a chain, so it continues this path with beta *= f |wi.n| / pdf
        if (bounces + 1 >= pp.max_depth) return;
        ScatterSample rs;
        if (!bsdf_sample_f<M>(bsdf, s.wo, sampler_uniform(key, dim0 + 5), sampler_uniform(key, dim0 + 6), BXDF_REFLECTION | BXDF_SPECULAR, &rs)) return;
        if (abs_dot(rs.wi, s.ns) == 0.0f) return;
        out->alive = true; out->next_o = spawn_origin(s, rs.wi); out->next_d = rs.wi;
        out->beta = beta * (rs.f * fabsf(dot(rs.wi, s.ns)) / rs.pdf);
        out->state = (uint32_t)(bounces + 1);
        if (IMG && have_df) {
            out->has_diff = true; out->diff = reflect_differential(sc, slot, s, df, td, dpdx, dpdy, rs.wi);
            out->state |= FTN_STATE_HAS_DIFF;
        }
        return;
    }
    // continuation: Bsdf::sample_f(wo, get_2d(), ALL), path.rs:68-76
    const float u0 = sampler_uniform(key, dim0 + 5), u1 = sampler_uniform(key, dim0 + 6);
    ScatterSample cs;
    if (!bsdf_sample_f<M>(bsdf, x_neg(ray.d), u0, u1, BXDF_ALL, &cs) || is_black(cs.f)) return;
    beta = beta * (cs.f * abs_dot(cs.wi, s.ns) / cs.pdf);
    const uint32_t spec = (cs.type & BXDF_SPECULAR) ? FTN_STATE_SPECULAR : 0u;
    const float mb = max_component(beta);
    if (mb < pp.rr_threshold && bounces > 3) {   // path.rs:84-91
        const float q = fmaxf(0.05f, 1.0f - mb);
        if (sampler_uniform(key, dim0 + 7) < q) return;
        beta = beta / (1.0f - q);
    }
    out->alive = true; out->next_o = spawn_origin(s, cs.wi); out->next_d = cs.wi;
    out->beta = beta; out->state = spec | (uint32_t)(bounces + 1);
}

template <int MAT, bool IMG = true>
FTN_HD void shade_surface(const SceneView& sc, const PassParams& pp, uint32_t path, const RayF& ray, uint32_t slot,
                          uint32_t state, V3 beta, V3 L, ShadeOut* out, uint32_t* err, const RayDiff* carried = nullptr) {
    shade_surface_to<MAT, IMG>(sc, pp, path, ray, slot, state, beta, L, out, &out->direct, err, carried);
}

// Radiance arriving along an MIS (BSDF-sampled) ray, integrator/mod.rs:364-389.
FTN_HD V3 mis_incident(const SceneView& sc, const LightData& light, const RayF& ray, uint32_t slot) {
    if (slot == FTN_NO_HIT_SLOT) return (light.type == 0) ? env_emitted(light.env, ray.d) : v3s(0.0f);
    if (slot & FTN_SPHERE_SLOT_FLAG) {
        const uint32_t si = slot & ~FTN_SPHERE_SLOT_FLAG;
        if (light.type == 1 && (uint32_t)light.sphere == si) {   // the SAME light only (:370-381)
            SphereHit sh;
            if (sphere_intersect(sc.spheres[si], ray, &sh)) return area_emitted(light, sh.n, x_neg(ray.d));
        }
    } else if (light.type == FTN_LIGHT_TYPE_TRIANGLE) {
        const uint32_t prim = f2u(ld4(sc.bvh.tris + (size_t)FTN_TRI_F4 * (size_t)slot).w);
        if (prim == (uint32_t)light.sphere) {                      // the triangle that carries this very light
            Surface hs;
            if (surface_at_hit(sc, slot, ray, &hs)) return area_emitted(light, hs.n, x_neg(ray.d));
        }
    }
    return v3s(0.0f);
}

// ---- film ---------------------------------------------------------------------------------------------------------
FTN_HD int iceil(float v) { return (int)ceilf(v); }
FTN_HD int ifloor(float v) { return (int)floorf(v); }
FTN_HD int imin(int a, int b) { return a < b ? a : b; }
FTN_HD int imax(int a, int b) { return a > b ? a : b; }

// Everything that Film::add_sample_to_tile (film.rs:136-172) would add to film pixel `i` from the
// samples of this pass, gathered in a fixed order (sample rows, sample columns, samples).  The
// footprint is p0 = ceil(pd - r), p1 = floor(pd + r) + 1 clipped to the pixel bounds of the
// sample's 16x16 tile (get_film_tile, film.rs:95-113, including its `- radius` in p1y).
// `spill[sample pixel]` != 0 marks sample pixels with at least one sample whose footprint is not
// exactly its own pixel (for the box filter of radius 0.5 that only happens when x + jitter rounds
// to x or x + 1); neighbours without the mark are skipped without reading their samples.
FTN_HD void film_gather_pixel(const PassParams& pp, const float2* p_film, const float4* Lbuf, const uint8_t* spill, int i, int reach,
                              float4* acc_io, uint32_t* err) {
    const FilmGeom& f = pp.film;
    const int fw = f.crop_max[0] - f.crop_min[0];
    const int px = f.crop_min[0] + i % fw, py = f.crop_min[1] + i / fw;
    const int sbw = f.sb_max[0] - f.sb_min[0];
    float4 acc = *acc_io;
    for (int sy = imax(py - reach, f.sb_min[1]); sy <= imin(py + reach, f.sb_max[1] - 1); ++sy) {
        const int ty0 = f.sb_min[1] + ((sy - f.sb_min[1]) / 16) * 16, ty1 = imin(ty0 + 16, f.sb_max[1]);
        const int tp0y = imax(iceil(rn_sub(rn_sub((float)ty0, 0.5f), f.radius[1])), f.crop_min[1]);
        const int tp1y = imin(iceil(rn_add(rn_sub(rn_sub((float)ty1, 0.5f), f.radius[1]), 1.0f)), f.crop_max[1]);
        if (py < tp0y || py >= tp1y) continue;
        for (int sx = imax(px - reach, f.sb_min[0]); sx <= imin(px + reach, f.sb_max[0] - 1); ++sx) {
            const int tx0 = f.sb_min[0] + ((sx - f.sb_min[0]) / 16) * 16, tx1 = imin(tx0 + 16, f.sb_max[0]);
            const int tp0x = imax(iceil(rn_sub(rn_sub((float)tx0, 0.5f), f.radius[0])), f.crop_min[0]);
            const int tp1x = imin(iceil(rn_add(rn_add(rn_sub((float)tx1, 0.5f), f.radius[0]), 1.0f)), f.crop_max[0]);
            if (px < tp0x || px >= tp1x) continue;
            const uint32_t spix = (uint32_t)(sy - f.sb_min[1]) * (uint32_t)sbw + (uint32_t)(sx - f.sb_min[0]);
            if ((sx != px || sy != py) && !spill[spix]) continue;
            const uint32_t base = spix * (uint32_t)pp.s_count;
            for (int s = 0; s < pp.s_count; ++s) {
                const float2 pf = p_film[base + s];
                const float dx = rn_sub(pf.x, 0.5f), dy = rn_sub(pf.y, 0.5f);
                const int p0x = iceil(rn_sub(dx, f.radius[0])), p1x = ifloor(rn_add(dx, f.radius[0])) + 1;
                const int p0y = iceil(rn_sub(dy, f.radius[1])), p1y = ifloor(rn_add(dy, f.radius[1])) + 1;
                if (px < p0x || px >= p1x || py < p0y || py >= p1y) continue;
                const float4 L = Lbuf[base + s];
                if (L.x != L.x || L.y != L.y || L.z != L.z) flag_error(err, ERR_NAN);   // check_radiance, integrator/mod.rs:285
                // BoxFilter::evaluate == 1 for every table entry (filter/mod.rs:17-19); ray weight 1
                acc.x = rn_add(acc.x, L.x); acc.y = rn_add(acc.y, L.y); acc.z = rn_add(acc.z, L.z); acc.w = rn_add(acc.w, 1.0f);
            }
        }
    }
    *acc_io = acc;
}

// rgb_to_xyz of the accumulated sums (spectrum/mod.rs:37-43) added into Film.pixels (merge_film_tile)
FTN_HD void film_resolve_pixel(float4 a, float4* px) {
    const float X = rn_add(rn_add(rn_mul(0.412453f, a.x), rn_mul(0.357580f, a.y)), rn_mul(0.180423f, a.z));
    const float Y = rn_add(rn_add(rn_mul(0.212671f, a.x), rn_mul(0.715160f, a.y)), rn_mul(0.072169f, a.z));
    const float Z = rn_add(rn_add(rn_mul(0.019334f, a.x), rn_mul(0.119193f, a.y)), rn_mul(0.950227f, a.z));
    px->x = rn_add(px->x, X); px->y = rn_add(px->y, Y); px->z = rn_add(px->z, Z); px->w = rn_add(px->w, a.w);
}

// Film::new / sample_bounds (film.rs:49-58, 86-93); host side
inline int film_geometry(const FtnFilm* f, FilmGeom* g) {
    if (f->x_resolution < 1 || f->y_resolution < 1) return FTN_ERR_INVALID_ARGUMENT;
    if (!(f->filter_radius[0] > 0.0f) || !(f->filter_radius[1] > 0.0f)) return FTN_ERR_INVALID_ARGUMENT;
    g->xres = f->x_resolution; g->yres = f->y_resolution;
    g->crop_min[0] = (int)ceilf((float)g->xres * f->crop_window[0]); g->crop_min[1] = (int)ceilf((float)g->yres * f->crop_window[2]);
    g->crop_max[0] = (int)ceilf((float)g->xres * f->crop_window[1]); g->crop_max[1] = (int)ceilf((float)g->yres * f->crop_window[3]);
    for (int a = 0; a < 2; ++a) { g->radius[a] = f->filter_radius[a]; g->inv_radius[a] = 1.0f / f->filter_radius[a]; }
    for (int a = 0; a < 2; ++a) {
        g->sb_min[a] = (int)floorf((float)g->crop_min[a] + 0.5f - g->radius[a]);
        g->sb_max[a] = (int)ceilf((float)g->crop_max[a] - 0.5f + g->radius[a]);
    }
    if (g->crop_max[0] <= g->crop_min[0] || g->crop_max[1] <= g->crop_min[1]) return FTN_ERR_INVALID_ARGUMENT;
    return FTN_OK;
}

}  // namespace ftn
