// TEST INFRASTRUCTURE ONLY (see ref_math.h).  Shapes, primitives, the recursive
// middle-split BVH and its stack traversal, restated from the reference.
#pragma once
#include "ref_math.h"
#include <vector>
#include <memory>
#include <functional>

namespace ref {

// interaction.rs:10
static const Float SHADOW_EPSILON = 0.0001f;

// interaction.rs:12-59
struct SurfaceHit {
    Point3 p; Vec3 p_err; Float time; Vec3 n;
    Ray spawn_ray(Vec3 dir) const {  // :22-30
        Ray r; r.origin = offset_ray_origin(p, p_err, n, dir); r.dir = dir; r.t_max = INF; r.time = time;
        return r;
    }
    Ray spawn_ray_to_hit(const SurfaceHit& to) const {  // :48-58
        Point3 origin = offset_ray_origin(p, p_err, n, to.p - p);
        Point3 target = offset_ray_origin(to.p, to.p_err, to.n, origin - to.p);
        Ray r; r.origin = origin; r.dir = target - origin; r.t_max = 1.0f - SHADOW_EPSILON; r.time = time;
        return r;
    }
};

// interaction.rs:61-108 (texture differentials omitted: only ConstantTexture is in scope)
struct SurfaceInteraction {
    SurfaceHit hit;
    Float uv[2];
    Vec3 wo;
    Vec3 dpdu, dpdv;                 // geom
    Vec3 shading_n;
    Vec3 shading_dpdu, shading_dpdv; // shading_geom
    Vec3 shading_dndu = Vec3(0.0f, 0.0f, 0.0f), shading_dndv = Vec3(0.0f, 0.0f, 0.0f);   // shading_geom.dndu / dndv (specular_reflect's differentials)
    Float dudx = 0.0f, dvdx = 0.0f, dudy = 0.0f, dvdy = 0.0f;   // tex_diffs (interaction.rs:193-215), default zero
    Vec3 dpdx = Vec3(0.0f, 0.0f, 0.0f), dpdy = Vec3(0.0f, 0.0f, 0.0f);                     // tex_diffs.dpdx / dpdy
    int prim;                        // index into Scene::prims (insertion order), -1 = none
    Float b[3];                      // triangle barycentrics (oracle-side extra, for the parity tests)
};

// ---- TriangleMesh / Triangle, shapes/triangle.rs ---------------------------------------
struct TriangleMesh {             // views into the scene-wide arrays (Scene owns the storage)
    const uint32_t* vertex_indices; // this mesh's first index
    const Point3* vertices;         // world space (triangle.rs:42-44 done by the host)
    const Vec3* normals;            // nullptr = None
    const Float* uvs;               // 2 per vertex, nullptr = None
    bool flip_normals;              // reverse_orientation ^ swaps_handedness (shapes/mod.rs:27-29)
};

// triangle.rs:428-434 -- compares SIGN BITS
inline bool sign_differs(Float v1, Float v2, Float v3) {
    return sign_positive(v1) != sign_positive(v2) || sign_positive(v2) != sign_positive(v3);
}

struct TriHitCore { Float t, b0, b1, b2; };

// triangle.rs:176-268: the watertight test proper, up to the delta_t cull.
inline bool triangle_intersect_core(Point3 p0, Point3 p1, Point3 p2, const Ray& ray, TriHitCore* out) {
    Vec3 p0t = p0 - ray.origin, p1t = p1 - ray.origin, p2t = p2 - ray.origin;   // :186-188
    int kz = max_dimension(vabs(ray.dir));                                      // :191
    int kx = (kz + 1) % 3, ky = (kx + 1) % 3;
    Vec3 dir(ray.dir[kx], ray.dir[ky], ray.dir[kz]);
    p0t = Vec3(p0t[kx], p0t[ky], p0t[kz]);
    p1t = Vec3(p1t[kx], p1t[ky], p1t[kz]);
    p2t = Vec3(p2t[kx], p2t[ky], p2t[kz]);
    Float shear_x = -dir.x / dir.z, shear_y = -dir.y / dir.z, shear_z = 1.0f / dir.z;  // :203-205
    p0t.x += shear_x * p0t.z; p0t.y += shear_y * p0t.z;
    p1t.x += shear_x * p1t.z; p1t.y += shear_y * p1t.z;
    p2t.x += shear_x * p2t.z; p2t.y += shear_y * p2t.z;
    Float e0 = p1t.x * p2t.y - p1t.y * p2t.x;   // :214-216
    Float e1 = p2t.x * p0t.y - p2t.y * p0t.x;
    Float e2 = p0t.x * p1t.y - p0t.y * p1t.x;
    if (e0 == 0.0f || e1 == 0.0f || e2 == 0.0f) {  // :219-223 f64 fallback
        e0 = (Float)((double)p1t.x * (double)p2t.y - (double)p1t.y * (double)p2t.x);
        e1 = (Float)((double)p2t.x * (double)p0t.y - (double)p2t.y * (double)p0t.x);
        e2 = (Float)((double)p0t.x * (double)p1t.y - (double)p0t.y * (double)p1t.x);
    }
    if (sign_differs(e0, e1, e2)) return false;   // :227
    Float det = e0 + e1 + e2;                      // :229
    if (det == 0.0f) return false;
    p0t.z *= shear_z; p1t.z *= shear_z; p2t.z *= shear_z;   // :233-235
    Float t_scaled = e0 * p0t.z + e1 * p1t.z + e2 * p2t.z;
    if ((det < 0.0f && (t_scaled >= 0.0f || t_scaled < ray.t_max * det)) ||
        (det > 0.0f && (t_scaled <= 0.0f || t_scaled > ray.t_max * det))) return false;   // :237-242
    Float inv_det = 1.0f / det;   // :246-250
    Float b0 = e0 * inv_det, b1 = e1 * inv_det, b2 = e2 * inv_det;
    Float t = t_scaled * inv_det;
    Float max_zt = fmax_(fmax_(std::fabs(p0t.z), std::fabs(p1t.z)), std::fabs(p2t.z));   // :254-267
    Float delta_z = gamma(3) * max_zt;
    Float max_xt = fmax_(fmax_(std::fabs(p0t.x), std::fabs(p1t.x)), std::fabs(p2t.x));
    Float max_yt = fmax_(fmax_(std::fabs(p0t.y), std::fabs(p1t.y)), std::fabs(p2t.y));
    Float delta_x = gamma(5) * (max_xt + max_zt);
    Float delta_y = gamma(5) * (max_yt + max_zt);
    Float delta_e = 2.0f * (gamma(2) * max_xt * max_yt + delta_y * max_xt + delta_x * max_yt);
    Float max_e = fmax_(fmax_(std::fabs(e0), std::fabs(e1)), std::fabs(e2));
    Float delta_t = 3.0f * (gamma(3) * max_e * max_zt + delta_e * max_zt + delta_z * max_e) * std::fabs(inv_det);
    if (t <= delta_t) return false;   // :268
    out->t = t; out->b0 = b0; out->b1 = b1; out->b2 = b2;
    return true;
}

struct Triangle {
    const TriangleMesh* mesh; uint32_t tri_id;
    void verts(Point3* p0, Point3* p1, Point3* p2, uint32_t v[3]) const {
        v[0] = mesh->vertex_indices[3 * tri_id]; v[1] = mesh->vertex_indices[3 * tri_id + 1]; v[2] = mesh->vertex_indices[3 * tri_id + 2];
        *p0 = mesh->vertices[v[0]]; *p1 = mesh->vertices[v[1]]; *p2 = mesh->vertices[v[2]];
    }
    Bounds3 world_bound() const {  // :151-157
        Point3 p0, p1, p2; uint32_t v[3]; verts(&p0, &p1, &p2, v);
        return Bounds3::empty().join_point(p0).join_point(p1).join_point(p2);
    }
    Float area() const {  // :171-174
        Point3 p0, p1, p2; uint32_t v[3]; verts(&p0, &p1, &p2, v);
        return 0.5f * magnitude(cross(p1 - p0, p2 - p0));
    }
    // triangle.rs:176-393
    bool intersect(const Ray& ray, Float* t_out, SurfaceInteraction* si) const {
        Point3 p0, p1, p2; uint32_t v[3]; verts(&p0, &p1, &p2, v);
        TriHitCore h;
        if (!triangle_intersect_core(p0, p1, p2, ray, &h)) return false;
        Float b0 = h.b0, b1 = h.b1, b2 = h.b2;
        // :271-294 partial derivatives
        Float uv[3][2] = {{0.0f, 0.0f}, {1.0f, 0.0f}, {1.0f, 1.0f}};
        if (mesh->uvs) for (int i = 0; i < 3; ++i) { uv[i][0] = mesh->uvs[2 * v[i]]; uv[i][1] = mesh->uvs[2 * v[i] + 1]; }
        Float duv02[2] = {uv[0][0] - uv[2][0], uv[0][1] - uv[2][1]};
        Float duv12[2] = {uv[1][0] - uv[2][0], uv[1][1] - uv[2][1]};
        Vec3 dp02 = p0 - p2, dp12 = p1 - p2;
        Float determinant = duv02[0] * duv12[1] - duv02[1] * duv12[0];
        bool degenerate_uv = std::fabs(determinant) < 1.0e-8f;
        Vec3 dpdu, dpdv;
        Float inv_det_uv = 0.0f;
        if (degenerate_uv) {
            Vec3 ng = cross(p2 - p0, p1 - p0);
            if (magnitude2(ng) == 0.0f) return false;
            coordinate_system(normalize(ng), &dpdu, &dpdv);
        } else {
            inv_det_uv = 1.0f / determinant;
            dpdu = (duv12[1] * dp02 - duv02[1] * dp12) * inv_det_uv;
            dpdv = (-duv12[0] * dp02 + duv02[0] * dp12) * inv_det_uv;
        }
        // :297-300
        Float xs = std::fabs(b0 * p0.x) + std::fabs(b1 * p1.x) + std::fabs(b2 * p2.x);
        Float ys = std::fabs(b0 * p0.y) + std::fabs(b1 * p1.y) + std::fabs(b2 * p2.y);
        Float zs = std::fabs(b0 * p0.z) + std::fabs(b1 * p1.z) + std::fabs(b2 * p2.z);
        Vec3 p_err = gamma(7) * Vec3(xs, ys, zs);
        // :303-304
        Point3 p_hit = b0 * p0 + b1 * p1 + b2 * p2;
        Float u_hit = b0 * uv[0][0] + b1 * uv[1][0] + b2 * uv[2][0];
        Float v_hit = b0 * uv[0][1] + b1 * uv[1][1] + b2 * uv[2][1];
        Vec3 geom_normal = normalize(cross(dp02, dp12));   // :315
        si->hit.p = p_hit; si->hit.p_err = p_err; si->hit.time = ray.time; si->hit.n = geom_normal;
        si->uv[0] = u_hit; si->uv[1] = v_hit;
        si->wo = -ray.dir;
        si->dpdu = dpdu; si->dpdv = dpdv;
        si->shading_n = geom_normal; si->shading_dpdu = dpdu; si->shading_dpdv = dpdv;
        si->b[0] = b0; si->b[1] = b1; si->b[2] = b2;
        if (mesh->flip_normals) {   // :327-330
            si->hit.n = si->hit.n * -1.0f;
            si->shading_n = si->shading_n * -1.0f;
        }
        if (mesh->normals) {   // :332-391 (tangents are never supplied by the loaders)
            Vec3 ns = normalize(b0 * mesh->normals[v[0]] + b1 * mesh->normals[v[1]] + b2 * mesh->normals[v[2]]);
            Vec3 ss = normalize(si->dpdu);
            Vec3 ts = cross(ns, ss);
            if (magnitude2(ts) > 0.0f) {
                ts = normalize(ts);
                ss = cross(ts, ns);
            } else {
                coordinate_system(ns, &ts, &ss);   // `let (ts, ss) = coordinate_system(ns.0)`
            }
            si->shading_dpdu = ss; si->shading_dpdv = ts;
            {   // :357-375 dndu / dndv of the shading geometry
                Vec3 dn1 = mesh->normals[v[0]] - mesh->normals[v[2]], dn2 = mesh->normals[v[1]] - mesh->normals[v[2]];
                if (degenerate_uv) {
                    Vec3 dn = cross(mesh->normals[v[2]] - mesh->normals[v[0]], mesh->normals[v[1]] - mesh->normals[v[0]]);
                    if (magnitude2(dn) == 0.0f) { si->shading_dndu = Vec3(0.0f, 0.0f, 0.0f); si->shading_dndv = Vec3(0.0f, 0.0f, 0.0f); }
                    else coordinate_system(dn, &si->shading_dndu, &si->shading_dndv);
                } else {
                    si->shading_dndu = (duv12[1] * dn1 - duv02[1] * dn2) * inv_det_uv;
                    si->shading_dndv = (-duv12[0] * dn1 + duv02[0] * dn2) * inv_det_uv;
                }
            }
            si->shading_n = ns;
            si->hit.n = faceforward(si->hit.n, si->shading_n);   // :390
        }
        *t_out = h.t;
        return true;
    }
    // triangle.rs:395-420
    SurfaceHit sample(Float u0, Float u1) const {
        Float su0 = std::sqrt(u0);                       // sampling.rs:48-51 uniform_sample_triangle
        Float bb0 = 1.0f - su0, bb1 = u1 * su0;
        Point3 p0, p1, p2; uint32_t v[3]; verts(&p0, &p1, &p2, v);
        Vec3 sample_p = bb0 * p0 + bb1 * p1 + (1.0f - bb0 - bb1) * p2;
        Vec3 n = normalize(cross(p1 - p0, p2 - p0));
        Vec3 sample_n;
        if (mesh->normals) {
            Vec3 ns = normalize(bb0 * mesh->normals[v[0]] + bb1 * mesh->normals[v[1]] + (1.0f - bb0 - bb1) * mesh->normals[v[2]]);
            sample_n = faceforward(n, ns);
        } else if (mesh->flip_normals) sample_n = n * -1.0f;
        else sample_n = n;
        Vec3 p_abs_sum = vabs(bb0 * p0) + vabs(bb1 * p1) + vabs((1.0f - bb0 - bb1) * p2);
        SurfaceHit h; h.p = Point3(0, 0, 0) + sample_p; h.p_err = gamma(6) * p_abs_sum; h.time = 0.0f; h.n = sample_n;
        return h;
    }
};

// math.rs:36-53
inline bool quadratic(EFloat a, EFloat b, EFloat c, EFloat* t0, EFloat* t1) {
    double discrim = (double)b.v * (double)b.v - (4.0 * (double)a.v * (double)c.v);
    if (discrim < 0.0) return false;
    double root_discrim = std::sqrt(discrim);
    EFloat rd = EFloat::with_err((Float)root_discrim, MACHINE_EPSILON * (Float)root_discrim);
    EFloat q = (b.v < 0.0f) ? (-0.5f * (b - rd)) : (-0.5f * (b + rd));
    EFloat r0 = q / a, r1 = c / q;
    if (r0.v > r1.v) { *t0 = r1; *t1 = r0; } else { *t0 = r0; *t1 = r1; }
    return true;
}

// ---- Sphere, shapes/sphere.rs -----------------------------------------------------------
struct Sphere {
    Transform object_to_world, world_to_object;
    bool reverse_orientation;
    Float radius, z_min, z_max, theta_min, theta_max, phi_max;
    void init(Float r, Float zmin, Float zmax, Float phimax_deg) {  // :30-49
        radius = r;
        z_min = clampf(fmin_(zmin, zmax), -r, r);
        z_max = clampf(fmax_(zmin, zmax), -r, r);
        theta_min = std::acos(clampf(zmin / r, -1.0f, 1.0f));
        theta_max = std::acos(clampf(zmax / r, -1.0f, 1.0f));
        phi_max = clampf(phimax_deg, 0.0f, 360.0f) * (PI / 180.0f);   // f32::to_radians = x * (PI/180)
    }
    bool flip_normals() const {   // shapes/mod.rs:23-29
        return reverse_orientation ^ swaps_handedness();
    }
    bool swaps_handedness() const {   // transform.rs:126-128: 4x4 determinant < 0; affine => upper 3x3
        const Mat4& m = object_to_world.t;
        Float det = m.m[0][0] * (m.m[1][1] * m.m[2][2] - m.m[2][1] * m.m[1][2])
                  - m.m[1][0] * (m.m[0][1] * m.m[2][2] - m.m[2][1] * m.m[0][2])
                  + m.m[2][0] * (m.m[0][1] * m.m[1][2] - m.m[1][1] * m.m[0][2]);
        return det < 0.0f;
    }
    Bounds3 object_bound() const {   // :61-63
        Bounds3 b; b.min = Point3(-radius, -radius, z_min); b.max = Point3(radius, radius, z_max); return b;
    }
    Bounds3 world_bound() const { return bounds_transform(object_to_world.t, object_bound()); }   // shapes/mod.rs:17-19
    Float area() const { return phi_max * radius * (z_max - z_min); }   // :77-79

    // :83-200
    bool intersect(const Ray& world_ray, Float* t_out, SurfaceInteraction* si) const {
        Vec3 o_err, d_err;
        Ray ray = ray_tf_exact_to_err(world_to_object.t, world_ray, &o_err, &d_err);
        EFloat ox = EFloat::with_err(ray.origin.x, o_err.x), oy = EFloat::with_err(ray.origin.y, o_err.y), oz = EFloat::with_err(ray.origin.z, o_err.z);
        EFloat dx = EFloat::with_err(ray.dir.x, d_err.x), dy = EFloat::with_err(ray.dir.y, d_err.y), dz = EFloat::with_err(ray.dir.z, d_err.z);
        EFloat a = dx * dx + dy * dy + dz * dz;
        EFloat b = 2.0f * (dx * ox + dy * oy + dz * oz);
        EFloat c = ox * ox + oy * oy + oz * oz - EFloat(radius) * EFloat(radius);
        EFloat t0, t1;
        if (!quadratic(a, b, c, &t0, &t1)) return false;
        if (t0.upper_bound() > ray.t_max || t1.lower_bound() <= 0.0f) return false;
        EFloat t_shape_hit = t0;
        bool is_t1 = false;
        if (t_shape_hit.lower_bound() <= 0.0f) {
            t_shape_hit = t1; is_t1 = true;
            if (t_shape_hit.upper_bound() > ray.t_max) return false;
        }
        Point3 p_hit = ray.at(t_shape_hit.v);
        p_hit = p_hit * (radius / magnitude(p_hit - Point3(0, 0, 0)));
        if (p_hit.x == 0.0f && p_hit.y == 0.0f) p_hit.x = 1.0e-5f * radius;
        Float phi = std::atan2(p_hit.y, p_hit.x);
        if (phi < 0.0f) phi += 2.0f * PI;
        if ((z_min > -radius && p_hit.z < z_min) || (z_max < radius && p_hit.z > z_max) || phi > phi_max) {
            // `t_shape_hit == t1` compares .v only (err_float.rs:104-108)
            if (is_t1 || t_shape_hit.v == t1.v) return false;
            if (t1.upper_bound() > ray.t_max) return false;
            t_shape_hit = t1;
            p_hit = ray.at(t_shape_hit.v);
            p_hit = p_hit * (radius / magnitude(p_hit - Point3(0, 0, 0)));
            if (p_hit.x == 0.0f && p_hit.y == 0.0f) p_hit.x = 1.0e-5f * radius;
            phi = std::atan2(p_hit.y, p_hit.x);
            if (phi < 0.0f) phi += 2.0f * PI;
            if ((z_min > -radius && p_hit.z < z_min) || (z_max < radius && p_hit.z > z_max) || phi > phi_max) return false;
        }
        Float u = phi / phi_max;
        Float theta = std::acos(clampf(p_hit.z / radius, -1.0f, 1.0f));
        Float v = (theta - theta_min) / (theta_max - theta_min);
        Float z_radius = std::sqrt(p_hit.x * p_hit.x + p_hit.y * p_hit.y);
        Float inv_z_radius = 1.0f / z_radius;
        Float cos_phi = p_hit.x * inv_z_radius, sin_phi = p_hit.y * inv_z_radius;
        Vec3 dpdu(-phi_max * p_hit.y, phi_max * p_hit.x, 0.0f);
        Vec3 dpdv = (theta_max - theta_min) * Vec3(p_hit.z * cos_phi, p_hit.z * sin_phi, -radius * std::sin(theta));
        Vec3 N = normalize(cross(dpdu, dpdv));
        // :160-178 Weingarten equations (N before the orientation flip)
        Vec3 d2pduu = (-phi_max * phi_max) * Vec3(p_hit.x, p_hit.y, 0.0f);
        Vec3 d2pduv = (theta_max - theta_min) * p_hit.z * phi_max * Vec3(-sin_phi, cos_phi, 0.0f);
        Vec3 d2pdvv = -(theta_max - theta_min) * (theta_max - theta_min) * Vec3(p_hit.x, p_hit.y, p_hit.z);
        Float E = dot(dpdu, dpdu), F = dot(dpdu, dpdv), G = dot(dpdv, dpdv);
        Float e = dot(N, d2pduu), f = dot(N, d2pduv), g = dot(N, d2pdvv);
        Float inv_egf2 = 1.0f / (E * G - F * F);
        Vec3 dndu = (f * F - e * G) * inv_egf2 * dpdu + (e * F - f * E) * inv_egf2 * dpdv;
        Vec3 dndv = (g * F - f * G) * inv_egf2 * dpdu + (f * F - g * E) * inv_egf2 * dpdv;
        Vec3 p_err = gamma(5) * vabs(p_hit);
        if (reverse_orientation) N = N * -1.0f;   // :183-185 (FIXME in the reference: ignores handedness)
        // SurfaceInteraction::transform, transform.rs:374-389
        const Transform& T = object_to_world;
        si->hit.p = point_tf_err_to_err(T.t, p_hit, p_err, &si->hit.p_err);
        si->hit.n = normalize(transform_normal(T, N));
        si->hit.time = ray.time;
        si->uv[0] = u; si->uv[1] = v;
        si->wo = normalize(transform_vector(T.t, -ray.dir));
        si->dpdu = transform_vector(T.t, dpdu); si->dpdv = transform_vector(T.t, dpdv);
        si->shading_n = normalize(transform_normal(T, N));
        si->shading_dpdu = si->dpdu; si->shading_dpdv = si->dpdv;
        si->shading_dndu = transform_normal(T, dndu); si->shading_dndv = transform_normal(T, dndv);   // transform.rs:353-354
        si->b[0] = si->b[1] = si->b[2] = 0.0f;
        *t_out = t_shape_hit.v;
        return true;
    }
    // :202-218
    SurfaceHit sample(Float u0, Float u1) const {
        Float z = 1.0f - 2.0f * u0;   // sampling.rs:37-42 uniform_sample_sphere
        Float r = std::sqrt(fmax_(1.0f - z * z, 0.0f));
        Float phi = 2.0f * PI * u1;
        Vec3 us(r * std::cos(phi), r * std::sin(phi), z);
        Point3 p_obj = Point3(0, 0, 0) + radius * us;
        Vec3 n = normalize(transform_normal(object_to_world, p_obj));
        if (reverse_orientation) n = n * -1.0f;
        p_obj = p_obj * (radius / magnitude(p_obj - Point3(0, 0, 0)));
        Vec3 p_obj_err = gamma(5) * vabs(p_obj);
        SurfaceHit h;
        h.p = point_tf_err_to_err(object_to_world.t, p_obj, p_obj_err, &h.p_err);
        h.time = 0.0f; h.n = n;
        return h;
    }
};

// ---- mipmap.rs: the pyramid is built by the host (MIPMap::new, :78-143, through the `resize` crate 0.4.3, which is
// not part of the hot path); the lookups below are the path's.
struct MIPMap {
    int wrap = 0;                     // ImageWrap :15-17: 0 Repeat, 1 Black, 2 Clamp
    std::vector<int> w, h;            // per level
    std::vector<std::vector<Spectrum>> pyramid;
    int levels() const { return (int)pyramid.size(); }
    // get_texel_from_level :297-311
    Spectrum texel(int level, int s, int t) const {
        int ss = w[level], ts = h[level];
        if (wrap == 0) { s = ((s % ss) + ss) % ss; t = ((t % ts) + ts) % ts; }   // rem_euclid
        else if (wrap == 2) { s = std::min(std::max(s, 0), ss - 1); t = std::min(std::max(t, 0), ts - 1); }
        else if (s < 0 || s >= ss || t < 0 || t >= ts) return Spectrum(0.0f);
        return pyramid[level][(size_t)t * ss + s];
    }
    // triangle :265-279
    Spectrum triangle(int level, Float st0, Float st1) const {
        level = std::min(std::max(level, 0), levels() - 1);
        Float s = st0 * (Float)w[level] - 0.5f, t = st1 * (Float)h[level] - 0.5f;
        int s0 = (int)std::floor(s), t0 = (int)std::floor(t);
        Float ds = s - (Float)s0, dt = t - (Float)t0;
        return texel(level, s0, t0) * (1.0f - ds) * (1.0f - dt) + texel(level, s0, t0 + 1) * (1.0f - ds) * dt
             + texel(level, s0 + 1, t0) * ds * (1.0f - dt) + texel(level, s0 + 1, t0 + 1) * ds * dt;
    }
    // lookup_trilinear_width :245-257
    Spectrum lookup_trilinear_width(Float st0, Float st1, Float width) const {
        Float level = (Float)levels() - 1.0f + std::log2(fmax_(width, 1.0e-8f));
        if (level < 0.0f) return triangle(0, st0, st1);
        if (level >= (Float)(levels() - 1)) return texel(levels() - 1, 0, 0);
        int lf = (int)std::floor(level);
        Float delta = level - std::trunc(level);   // f32::fract
        return triangle(lf, st0, st1) * (1.0f - delta) + triangle(lf + 1, st0, st1) * delta;   // Spectrum::lerp, spectrum/mod.rs:84-86
    }
    // lookup_trilinear :259-262 -- `dst0.y` enters without abs(), as in the reference
    Spectrum lookup_trilinear(Float st0, Float st1, const Float dst0[2], const Float dst1[2]) const {
        Float width = fmax_(fmax_(std::fabs(dst0[0]), dst0[1]), fmax_(std::fabs(dst1[0]), std::fabs(dst1[1])));
        return lookup_trilinear_width(st0, st1, 2.0f * width);
    }
};

// ---- materials / lights tables ------------------------------------------------------------
// One entry of the scene's texture table: texture/{mod,checkerboard,uv,image}.rs behind UVMapping (mapping.rs:36-53)
struct TextureDef {
    int type;          // FtnTextureType
    Spectrum value, tex1, tex2; Float uv_scale[2], uv_delta[2];
    std::shared_ptr<MIPMap> image;
};
struct Material {
    int type;          // FtnMaterialType
    uint32_t ptex[FTN_PARAM_COUNT];                 // per parameter: 0 = the constant, k = (*table)[k - 1]  (loaders/constructors.rs:192-238)
    const std::vector<TextureDef>* table;
    Spectrum kd, ks, eta, k, kr, kt;
    int kd_texture;    // FtnTextureType
    Spectrum tex1, tex2; Float uv_scale[2], uv_delta[2];
    Float u_rough, v_rough, sigma;
    bool remap;
    std::shared_ptr<MIPMap> image;   // kd_texture == FTN_TEXTURE_IMAGE (texture/image.rs:8-15)
};

// primitive.rs:25-71 GeometricPrimitive<S>
struct Primitive {
    int kind;          // 0 triangle, 1 sphere
    Triangle tri;
    const Sphere* sphere;
    int material;      // -1 none
    int light;         // index into Scene::lights of its area light, -1 none
    Bounds3 world_bound() const { return kind == 0 ? tri.world_bound() : sphere->world_bound(); }
    bool shape_intersect(const Ray& r, Float* t, SurfaceInteraction* si) const {
        return kind == 0 ? tri.intersect(r, t, si) : sphere->intersect(r, t, si);
    }
    Float area() const { return kind == 0 ? tri.area() : sphere->area(); }
    SurfaceHit sample(Float u0, Float u1) const { return kind == 0 ? tri.sample(u0, u1) : sphere->sample(u0, u1); }
};

// ---- BVH, bvh.rs ------------------------------------------------------------------------
struct LinearBVHNode {   // bvh.rs:269-302, 32 bytes
    Bounds3 bounds;
    uint32_t idx;        // leaf: first_prim_idx; interior: second_child_idx
    uint16_t n_prims;    // 0 => interior
    uint8_t split_axis;
    uint8_t pad;
};

struct TraversalCounters { uint64_t nodes = 0, prims = 0; };

struct BVH {
    std::vector<int> prim_order;       // post-permutation: slot -> insertion index (bvh.rs:52 apply_permutation)
    const std::vector<Primitive>* prims = nullptr;
    Bounds3 bounds;
    std::vector<LinearBVHNode> nodes;

    struct PrimInfo { int prim_id; Bounds3 bounds; Point3 centroid; };
    struct BuildNode { Bounds3 bounds; int child[2]; int first; int n; int axis; };
    std::vector<BuildNode> build_nodes;

    // bvh.rs:27-64
    void build(const std::vector<Primitive>* prims_in) {
        prims = prims_in;
        nodes.clear(); prim_order.clear(); build_nodes.clear();
        if (prims->empty()) { bounds = Bounds3::empty(); return; }
        std::vector<PrimInfo> info(prims->size());
        for (size_t i = 0; i < prims->size(); ++i) {
            info[i].prim_id = (int)i;
            info[i].bounds = (*prims)[i].world_bound();
            info[i].centroid = info[i].bounds.centroid();
        }
        prim_order.reserve(prims->size());
        int root = recursive_build(info.data(), info.size());
        bounds = build_nodes[root].bounds;
        nodes.reserve(2 * prims->size());
        flatten(root);
        build_nodes.clear(); build_nodes.shrink_to_fit();
    }
    // bvh.rs:66-120 (SplitMethod::Middle)
    int recursive_build(PrimInfo* info, size_t n) {
        Bounds3 node_bounds = Bounds3::empty(), centroid_bounds = Bounds3::empty();
        for (size_t i = 0; i < n; ++i) {
            node_bounds = node_bounds.join(info[i].bounds);
            centroid_bounds = centroid_bounds.join_point(info[i].centroid);
        }
        if (n == 1 || centroid_bounds.is_point()) {
            BuildNode bn; bn.bounds = node_bounds; bn.child[0] = bn.child[1] = -1;
            bn.first = (int)prim_order.size(); bn.n = (int)n; bn.axis = 0;
            for (size_t i = 0; i < n; ++i) prim_order.push_back(info[i].prim_id);
            build_nodes.push_back(bn);
            return (int)build_nodes.size() - 1;
        }
        int ax = centroid_bounds.maximum_extent();
        Float midpoint = (centroid_bounds.min[ax] + centroid_bounds.max[ax]) / 2.0f;
        // partition crate 0.1.2 (absent; Cargo.lock:977): any in-place partition gives the same
        // two SETS, hence the same subtree bounds; element order inside a part is not pinned.
        PrimInfo* mid_ptr = std::partition(info, info + n, [&](const PrimInfo& p) { return p.centroid[ax] < midpoint; });
        size_t mid = (size_t)(mid_ptr - info);
        if (mid == 0 || mid == n) {   // bvh.rs:122-130 partition_equal_counts
            mid = n / 2;
            std::nth_element(info, info + mid, info + n, [&](const PrimInfo& a, const PrimInfo& b) { return a.centroid[ax] < b.centroid[ax]; });
        }
        int c0 = recursive_build(info, mid);
        int c1 = recursive_build(info + mid, n - mid);
        BuildNode bn; bn.bounds = build_nodes[c0].bounds.join(build_nodes[c1].bounds);   // bvh.rs:337-343
        bn.child[0] = c0; bn.child[1] = c1; bn.first = 0; bn.n = 0; bn.axis = ax;
        build_nodes.push_back(bn);
        return (int)build_nodes.size() - 1;
    }
    // bvh.rs:132-158
    size_t flatten(int bi) {
        const BuildNode bn = build_nodes[bi];
        LinearBVHNode ln; ln.bounds = bn.bounds; ln.pad = 0;
        if (bn.child[0] < 0) {
            ln.idx = (uint32_t)bn.first; ln.n_prims = (uint16_t)bn.n; ln.split_axis = 0;
            nodes.push_back(ln);
            return 1;
        }
        ln.idx = 0; ln.n_prims = 0; ln.split_axis = (uint8_t)bn.axis;
        nodes.push_back(ln);
        size_t my_idx = nodes.size() - 1;
        size_t first_len = flatten(bn.child[0]);
        nodes[my_idx].idx = (uint32_t)(my_idx + first_len + 1);
        size_t second_len = flatten(bn.child[1]);
        return first_len + second_len + 1;
    }

    // primitive.rs:48-54 GeometricPrimitive::intersect
    bool prim_intersect(int slot, Ray* ray, SurfaceInteraction* si) const {
        int pid = prim_order[slot];
        Float t; SurfaceInteraction tmp;
        if (!(*prims)[pid].shape_intersect(*ray, &t, &tmp)) return false;
        ray->t_max = t;
        tmp.prim = pid;
        *si = tmp;
        return true;
    }
    // bvh.rs:160-215
    bool intersect(Ray* ray, SurfaceInteraction* si, TraversalCounters* ctr = nullptr) const {
        if (nodes.empty()) return false;
        bool dir_is_neg[3] = {ray->dir.x < 0.0f, ray->dir.y < 0.0f, ray->dir.z < 0.0f};
        size_t stack[64]; int sp = 0;
        size_t cur = 0;
        bool found = false;
        for (;;) {
            const LinearBVHNode& node = nodes[cur];
            if (ctr) ctr->nodes++;
            if (node.bounds.intersect_test(*ray)) {
                if (node.n_prims > 0) {
                    for (uint32_t i = 0; i < node.n_prims; ++i) {
                        if (ctr) ctr->prims++;
                        if (prim_intersect((int)(node.idx + i), ray, si)) found = true;
                    }
                    if (sp == 0) break;
                    cur = stack[--sp];
                } else {
                    if (dir_is_neg[node.split_axis]) { stack[sp++] = cur + 1; cur = node.idx; }
                    else { stack[sp++] = node.idx; cur = cur + 1; }
                }
            } else {
                if (sp == 0) break;
                cur = stack[--sp];
            }
        }
        return found;
    }
    // bvh.rs:217-266 (Shape::intersect_test defaults to intersect().is_some(), shapes/mod.rs:35)
    bool intersect_test(const Ray& ray, TraversalCounters* ctr = nullptr) const {
        if (nodes.empty()) return false;
        bool dir_is_neg[3] = {ray.dir.x < 0.0f, ray.dir.y < 0.0f, ray.dir.z < 0.0f};
        size_t stack[64]; int sp = 0;
        size_t cur = 0;
        for (;;) {
            const LinearBVHNode& node = nodes[cur];
            if (ctr) ctr->nodes++;
            if (node.bounds.intersect_test(ray)) {
                if (node.n_prims > 0) {
                    for (uint32_t i = 0; i < node.n_prims; ++i) {
                        if (ctr) ctr->prims++;
                        Float t; SurfaceInteraction tmp;
                        if ((*prims)[prim_order[node.idx + i]].shape_intersect(ray, &t, &tmp)) return true;
                    }
                    if (sp == 0) break;
                    cur = stack[--sp];
                } else {
                    if (dir_is_neg[node.split_axis]) { stack[sp++] = cur + 1; cur = node.idx; }
                    else { stack[sp++] = node.idx; cur = cur + 1; }
                }
            } else {
                if (sp == 0) break;
                cur = stack[--sp];
            }
        }
        return false;
    }
};

}  // namespace ref
