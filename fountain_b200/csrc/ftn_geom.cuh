// Watertight triangle test, EFloat sphere test and ray/box slabs for the sm_100a traversal
// kernels.  Exactly rounded (see ftn_common.cuh): results must be bit-identical to
// fountain's src/shapes/triangle.rs:176-268 and src/shapes/sphere.rs:83-200.
#pragma once
#include "ftn_common.cuh"

namespace ftn {

// triangle.rs:428-434 -- sign BITS, so -0.0 counts as negative
FTN_HD bool sign_differs(float a, float b, float c) {
    return sign_positive(a) != sign_positive(b) || sign_positive(b) != sign_positive(c);
}

// (kx,ky,kz) is always a cyclic rotation of (0,1,2) (kx = kz+1, ky = kx+1 mod 3), so the
// permutation is two selects per component instead of indexed (branchy) component access.
FTN_HD V3 permute(V3 p, int kx, int ky, int kz) {
    (void)kx; (void)ky;
    const bool r0 = kz == 2, r1 = kz == 0;   // kz==2: (x,y,z); kz==0: (y,z,x); kz==1: (z,x,y)
    return V3(r0 ? p.x : (r1 ? p.y : p.z), r0 ? p.y : (r1 ? p.z : p.x), r0 ? p.z : (r1 ? p.x : p.y));
}

// Per-ray constants of the watertight test (triangle.rs:191-205: the "TODO: cache shear
// coefficients in ray" the reference leaves open).  Same values, computed once per ray.
struct RayShear {
    int kx, ky, kz;
    float sx, sy, sz;
};
FTN_HD RayShear make_ray_shear(V3 d) {
    RayShear s;
    s.kz = max_dimension(x_abs(d));
    s.kx = (s.kz + 1) % 3;
    s.ky = (s.kx + 1) % 3;
    const V3 dp = permute(d, s.kx, s.ky, s.kz);
    float dx = dp.x, dy = dp.y, dz = dp.z;
    s.sx = rn_div(-dx, dz);
    s.sy = rn_div(-dy, dz);
    s.sz = rn_div(1.0f, dz);
    return s;
}

struct TriHit { float t, b0, b1, b2; };

// triangle.rs:176-268.  `t_max` is the ray's current t_max (shrinks during closest-hit search).
FTN_HD bool triangle_intersect(V3 p0, V3 p1, V3 p2, V3 ro, const RayShear& rs, float t_max, TriHit* out) {
    V3 p0t = permute(x_sub(p0, ro), rs.kx, rs.ky, rs.kz);
    V3 p1t = permute(x_sub(p1, ro), rs.kx, rs.ky, rs.kz);
    V3 p2t = permute(x_sub(p2, ro), rs.kx, rs.ky, rs.kz);
    p0t.x = rn_add(p0t.x, rn_mul(rs.sx, p0t.z)); p0t.y = rn_add(p0t.y, rn_mul(rs.sy, p0t.z));
    p1t.x = rn_add(p1t.x, rn_mul(rs.sx, p1t.z)); p1t.y = rn_add(p1t.y, rn_mul(rs.sy, p1t.z));
    p2t.x = rn_add(p2t.x, rn_mul(rs.sx, p2t.z)); p2t.y = rn_add(p2t.y, rn_mul(rs.sy, p2t.z));
    float e0 = rn_sub(rn_mul(p1t.x, p2t.y), rn_mul(p1t.y, p2t.x));
    float e1 = rn_sub(rn_mul(p2t.x, p0t.y), rn_mul(p2t.y, p0t.x));
    float e2 = rn_sub(rn_mul(p0t.x, p1t.y), rn_mul(p0t.y, p1t.x));
    if (e0 == 0.0f || e1 == 0.0f || e2 == 0.0f) {   // :219-223, f64 retry
        e0 = (float)rn_dsub(rn_dmul((double)p1t.x, (double)p2t.y), rn_dmul((double)p1t.y, (double)p2t.x));
        e1 = (float)rn_dsub(rn_dmul((double)p2t.x, (double)p0t.y), rn_dmul((double)p2t.y, (double)p0t.x));
        e2 = (float)rn_dsub(rn_dmul((double)p0t.x, (double)p1t.y), rn_dmul((double)p0t.y, (double)p1t.x));
    }
    if (sign_differs(e0, e1, e2)) return false;
    float det = rn_add(rn_add(e0, e1), e2);
    if (det == 0.0f) return false;
    p0t.z = rn_mul(p0t.z, rs.sz); p1t.z = rn_mul(p1t.z, rs.sz); p2t.z = rn_mul(p2t.z, rs.sz);
    float t_scaled = rn_add(rn_add(rn_mul(e0, p0t.z), rn_mul(e1, p1t.z)), rn_mul(e2, p2t.z));
    float tmd = rn_mul(t_max, det);
    if ((det < 0.0f && (t_scaled >= 0.0f || t_scaled < tmd)) ||
        (det > 0.0f && (t_scaled <= 0.0f || t_scaled > tmd))) return false;
    float inv_det = rn_div(1.0f, det);
    float b0 = rn_mul(e0, inv_det), b1 = rn_mul(e1, inv_det), b2 = rn_mul(e2, inv_det);
    float t = rn_mul(t_scaled, inv_det);
    // :254-268 conservative t > 0
    float max_zt = fmaxf(fmaxf(fabsf(p0t.z), fabsf(p1t.z)), fabsf(p2t.z));
    float delta_z = rn_mul(gamma_n(3), max_zt);
    float max_xt = fmaxf(fmaxf(fabsf(p0t.x), fabsf(p1t.x)), fabsf(p2t.x));
    float max_yt = fmaxf(fmaxf(fabsf(p0t.y), fabsf(p1t.y)), fabsf(p2t.y));
    float delta_x = rn_mul(gamma_n(5), rn_add(max_xt, max_zt));
    float delta_y = rn_mul(gamma_n(5), rn_add(max_yt, max_zt));
    float delta_e = rn_mul(2.0f, rn_add(rn_add(rn_mul(rn_mul(gamma_n(2), max_xt), max_yt), rn_mul(delta_y, max_xt)), rn_mul(delta_x, max_yt)));
    float max_e = fmaxf(fmaxf(fabsf(e0), fabsf(e1)), fabsf(e2));
    float delta_t = rn_mul(rn_mul(3.0f, rn_add(rn_add(rn_mul(rn_mul(gamma_n(3), max_e), max_zt), rn_mul(delta_e, max_zt)), rn_mul(delta_z, max_e))), fabsf(inv_det));
    if (t <= delta_t) return false;
    out->t = t; out->b0 = b0; out->b1 = b1; out->b2 = b2;
    return true;
}

// ---- EFloat, err_float.rs:33-193 ------------------------------------------------------------------
struct EF { float v, lo, hi; };
FTN_HD EF ef(float v) { EF e; e.v = v; e.lo = v; e.hi = v; return e; }
FTN_HD EF ef_err(float v, float err) {
    if (err == 0.0f) return ef(v);
    EF e; e.v = v; e.lo = next_float_down(rn_sub(v, err)); e.hi = next_float_up(rn_add(v, err)); return e;
}
FTN_HD EF ef_add(EF a, EF b) { EF r; r.v = rn_add(a.v, b.v); r.lo = next_float_down(rn_add(a.lo, b.lo)); r.hi = next_float_up(rn_add(a.hi, b.hi)); return r; }
FTN_HD EF ef_sub(EF a, EF b) { EF r; r.v = rn_sub(a.v, b.v); r.lo = next_float_down(rn_sub(a.lo, b.lo)); r.hi = next_float_up(rn_sub(a.hi, b.hi)); return r; }
FTN_HD EF ef_mul(EF a, EF b) {
    float p1 = rn_mul(a.lo, b.lo), p2 = rn_mul(a.hi, b.lo), p3 = rn_mul(a.lo, b.hi), p4 = rn_mul(a.hi, b.hi);
    EF r; r.v = rn_mul(a.v, b.v);
    r.lo = next_float_down(fminf(fminf(p1, p2), fminf(p3, p4)));
    r.hi = next_float_up(fmaxf(fmaxf(p1, p2), fmaxf(p3, p4)));
    return r;
}
FTN_HD EF ef_div(EF a, EF b) {
    EF r; r.v = rn_div(a.v, b.v);
    if (b.lo < 0.0f && b.hi > 0.0f) { r.lo = -FTN_INF; r.hi = FTN_INF; return r; }
    float d1 = rn_div(a.lo, b.lo), d2 = rn_div(a.hi, b.lo), d3 = rn_div(a.lo, b.hi), d4 = rn_div(a.hi, b.hi);
    r.lo = next_float_down(fminf(fminf(d1, d2), fminf(d3, d4)));
    r.hi = next_float_up(fmaxf(fmaxf(d1, d2), fmaxf(d3, d4)));
    return r;
}
FTN_HD EF ef_neg(EF a) { EF r; r.v = -a.v; r.lo = -a.hi; r.hi = -a.lo; return r; }

// math.rs:36-53
FTN_HD bool ef_quadratic(EF a, EF b, EF c, EF* t0, EF* t1) {
    double discrim = rn_dsub(rn_dmul((double)b.v, (double)b.v), rn_dmul(rn_dmul(4.0, (double)a.v), (double)c.v));
    if (discrim < 0.0) return false;
    double root = sqrt(discrim);
    EF rd = ef_err((float)root, rn_mul(FTN_MACHINE_EPS, (float)root));
    EF q = (b.v < 0.0f) ? ef_mul(ef(-0.5f), ef_sub(b, rd)) : ef_mul(ef(-0.5f), ef_add(b, rd));
    EF r0 = ef_div(q, a), r1 = ef_div(c, q);
    if (r0.v > r1.v) { *t0 = r1; *t1 = r0; } else { *t0 = r0; *t1 = r1; }
    return true;
}

// shapes/sphere.rs:16-58 after construction
struct SphereData {
    M4 o2w, w2o;
    float radius, z_min, z_max, theta_min, theta_max, phi_max;
    int reverse_orientation;
    int material;      // -1 none
    int light;         // index into the light table, -1 none
    float emit[3];
    float area;        // sphere.rs:77-79
};

struct SphereHit {
    float t;
    V3 p, p_err, n;    // world space (SurfaceHit after SurfaceInteraction::transform)
    V3 wo;             // normalised, world
    V3 dpdu, dpdv;     // world (dpdv only feeds texture differentials)
    V3 ns;             // shading normal (== transformed geometric normal)
    float u, v;        // phi / phi_max, (theta - theta_min) / (theta_max - theta_min), sphere.rs:146-148
};

// sphere.rs:83-200 + SurfaceInteraction::transform (transform.rs:374-389)
FTN_HD bool sphere_intersect(const SphereData& s, const RayF& wray, SphereHit* h) {
    V3 oe, de;
    RayF ray = ray_transform_err(s.w2o, wray, &oe, &de);
    EF ox = ef_err(ray.o.x, oe.x), oy = ef_err(ray.o.y, oe.y), oz = ef_err(ray.o.z, oe.z);
    EF dx = ef_err(ray.d.x, de.x), dy = ef_err(ray.d.y, de.y), dz = ef_err(ray.d.z, de.z);
    EF a = ef_add(ef_add(ef_mul(dx, dx), ef_mul(dy, dy)), ef_mul(dz, dz));
    EF b = ef_mul(ef(2.0f), ef_add(ef_add(ef_mul(dx, ox), ef_mul(dy, oy)), ef_mul(dz, oz)));
    EF c = ef_sub(ef_add(ef_add(ef_mul(ox, ox), ef_mul(oy, oy)), ef_mul(oz, oz)), ef_mul(ef(s.radius), ef(s.radius)));
    EF t0, t1;
    if (!ef_quadratic(a, b, c, &t0, &t1)) return false;
    if (t0.hi > ray.t_max || t1.lo <= 0.0f) return false;
    EF th = t0;
    if (th.lo <= 0.0f) { th = t1; if (th.hi > ray.t_max) return false; }
    V3 p = x_add(ray.o, x_scale(ray.d, th.v));
    p = x_scale(p, rn_div(s.radius, x_len(p)));
    if (p.x == 0.0f && p.y == 0.0f) p.x = rn_mul(1.0e-5f, s.radius);
    float phi = atan2f(p.y, p.x);
    if (phi < 0.0f) phi = rn_add(phi, rn_mul(2.0f, FTN_PI));
    if ((s.z_min > -s.radius && p.z < s.z_min) || (s.z_max < s.radius && p.z > s.z_max) || phi > s.phi_max) {
        if (th.v == t1.v) return false;
        if (t1.hi > ray.t_max) return false;
        th = t1;
        p = x_add(ray.o, x_scale(ray.d, th.v));
        p = x_scale(p, rn_div(s.radius, x_len(p)));
        if (p.x == 0.0f && p.y == 0.0f) p.x = rn_mul(1.0e-5f, s.radius);
        phi = atan2f(p.y, p.x);
        if (phi < 0.0f) phi = rn_add(phi, rn_mul(2.0f, FTN_PI));
        if ((s.z_min > -s.radius && p.z < s.z_min) || (s.z_max < s.radius && p.z > s.z_max) || phi > s.phi_max) return false;
    }
    float theta = acosf(clampf(rn_div(p.z, s.radius), -1.0f, 1.0f));
    float z_radius = rn_sqrt(rn_add(rn_mul(p.x, p.x), rn_mul(p.y, p.y)));
    float inv_zr = rn_div(1.0f, z_radius);
    float cos_phi = rn_mul(p.x, inv_zr), sin_phi = rn_mul(p.y, inv_zr);
    V3 dpdu = V3(rn_mul(-s.phi_max, p.y), rn_mul(s.phi_max, p.x), 0.0f);
    float dth = rn_sub(s.theta_max, s.theta_min);
    V3 dpdv = x_scale(V3(rn_mul(p.z, cos_phi), rn_mul(p.z, sin_phi), rn_mul(-s.radius, sinf(theta))), dth);
    V3 N = x_normalize(x_cross(dpdu, dpdv));
    V3 p_err = x_scale(x_abs(p), gamma_n(5));
    if (s.reverse_orientation) N = x_scale(N, -1.0f);
    h->p = point_tf_err_to_err(s.o2w, p, p_err, &h->p_err);
    h->n = x_normalize(transform_normal_inv(s.w2o, N));
    h->ns = h->n;
    h->wo = x_normalize(transform_vector(s.o2w, x_neg(ray.d)));
    h->dpdu = transform_vector(s.o2w, dpdu);
    h->dpdv = transform_vector(s.o2w, dpdv);
    h->u = rn_div(phi, s.phi_max); h->v = rn_div(rn_sub(theta, s.theta_min), dth);   // sphere.rs:146-148
    h->t = th.v;
    return true;
}

// ---- ray / box ---------------------------------------------------------------------------------------
// The reference's slab test (bounds.rs:214-233) recomputes 1/d per node and widens t_far by
// 1+2*gamma(3).  The traversal kernels hoist 1/d and o/d out of the loop and use FMAs, which
// rounds differently, so they widen by a larger factor on BOTH ends: the set of boxes entered
// is a superset of the reference's, and hit/miss is decided by the (exact) primitive tests.
struct RayBox { V3 inv_d; V3 o_inv_d; };   // t = p * inv_d - o_inv_d
FTN_HD float safe_inv(float d) {
    // 1/0 = +-inf is fine for slabs as long as 0*inf NaNs are ignored by fmin/fmax (they are).
    return 1.0f / d;
}

}  // namespace ftn
