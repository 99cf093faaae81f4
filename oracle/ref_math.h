// TEST INFRASTRUCTURE ONLY -- CPU restatement of akofke/fountain's arithmetic.
// Nothing under oracle/ is linked into, imported by or called from the product
// (fountain_b200/ + libfountain_gpu.so).  Only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs may use it.
//
// The reference (pure Rust, nightly 2020 + git deps) cannot be built here, so this
// is a line-by-line restatement; each function cites the reference file:line it
// follows.  Vector arithmetic the reference takes from cgmath 0.17.0 (absent from
// /root/reference, Cargo.lock:173) is restated from its published source:
//   dot(a,b)      = (a.x*b.x + a.y*b.y) + a.z*b.z           (mul_element_wise().sum())
//   cross(a,b)    = (a.y*b.z - a.z*b.y, a.z*b.x - a.x*b.z, a.x*b.y - a.y*b.x)
//   normalize(v)  = v * (1.0 / sqrt(dot(v,v)))               (normalize_to(1))
//   M * v4        = M[0]*v.x + M[1]*v.y + M[2]*v.z + M[3]*v.w (column-major)
//   transform_point = (M * (p,1)).xyz * (1 / w)
// Compile with -ffp-contract=off: Rust never contracts a*b+c.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <algorithm>

namespace ref {

typedef float Float;
static const Float PI = 3.14159265358979323846f;       // std::f32::consts::PI
static const Float FRAC_1_PI = 0.318309886183790671538f;
static const Float FRAC_PI_2 = 1.57079632679489661923f;
static const Float FRAC_PI_4 = 0.785398163397448309616f;
static const Float INF = std::numeric_limits<float>::infinity();

// Rust f32::max / f32::min ignore a NaN operand (like fmaxf/fminf).
inline Float fmax_(Float a, Float b) { return std::fmax(a, b); }
inline Float fmin_(Float a, Float b) { return std::fmin(a, b); }
// Rust f32::clamp (feature clamp): max then min with plain comparisons; NaN stays NaN.
inline Float clampf(Float x, Float lo, Float hi) {
    Float r = x;
    if (r < lo) r = lo;
    if (r > hi) r = hi;
    return r;
}
inline uint32_t f2u(Float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
inline Float u2f(uint32_t u) { Float f; std::memcpy(&f, &u, 4); return f; }
inline bool sign_positive(Float f) { return (f2u(f) >> 31) == 0; }  // f32::is_sign_positive

// ---- err_float.rs:5-30 ---------------------------------------------------------------
static const Float MACHINE_EPSILON = std::numeric_limits<float>::epsilon() * 0.5f;
inline Float gamma(int n) {  // err_float.rs:7-10, evaluated in f32
    Float nf = (Float)n;
    return (nf * MACHINE_EPSILON) / (1.0f - nf * MACHINE_EPSILON);
}
inline Float next_float_up(Float v) {  // err_float.rs:12-20
    if (v == INF) return v;
    if (v == -0.0f) v = 0.0f;
    uint32_t bits = f2u(v);
    bits = (v >= 0.0f) ? bits + 1 : bits - 1;
    return u2f(bits);
}
inline Float next_float_down(Float v) {  // err_float.rs:22-30
    if (v == -INF) return v;
    if (v == 0.0f) v = -0.0f;
    uint32_t bits = f2u(v);
    bits = (v >= 0.0f) ? bits - 1 : bits + 1;
    return u2f(bits);
}

// ---- EFloat, err_float.rs:33-193 -----------------------------------------------------
struct EFloat {
    Float v, low, high;
    EFloat() : v(0), low(0), high(0) {}
    explicit EFloat(Float v_) : v(v_), low(v_), high(v_) {}
    static EFloat with_err(Float v, Float err) {  // :47-57
        if (err == 0.0f) return with_bounds(v, v, v);
        return with_bounds(v, next_float_down(v - err), next_float_up(v + err));
    }
    static EFloat with_bounds(Float v, Float low, Float high) {
        EFloat e; e.v = v; e.low = low; e.high = high; return e;
    }
    Float upper_bound() const { return high; }
    Float lower_bound() const { return low; }
};
inline EFloat operator+(EFloat a, EFloat b) {  // :116-125
    return EFloat::with_bounds(a.v + b.v, next_float_down(a.low + b.low), next_float_up(a.high + b.high));
}
inline EFloat operator-(EFloat a, EFloat b) {  // :127-136 (low-low, high-high, as written)
    return EFloat::with_bounds(a.v - b.v, next_float_down(a.low - b.low), next_float_up(a.high - b.high));
}
inline EFloat operator*(EFloat a, EFloat b) {  // :138-158
    Float p1 = a.low * b.low, p2 = a.high * b.low, p3 = a.low * b.high, p4 = a.high * b.high;
    Float lo = next_float_down(fmin_(fmin_(p1, p2), fmin_(p3, p4)));
    Float hi = next_float_up(fmax_(fmax_(p1, p2), fmax_(p3, p4)));
    return EFloat::with_bounds(a.v * b.v, lo, hi);
}
inline EFloat operator/(EFloat a, EFloat b) {  // :160-185
    Float v = a.v / b.v;
    if (b.low < 0.0f && b.high > 0.0f) return EFloat::with_bounds(v, -INF, INF);
    Float d1 = a.low / b.low, d2 = a.high / b.low, d3 = a.low / b.high, d4 = a.high / b.high;
    Float lo = next_float_down(fmin_(fmin_(d1, d2), fmin_(d3, d4)));
    Float hi = next_float_up(fmax_(fmax_(d1, d2), fmax_(d3, d4)));
    return EFloat::with_bounds(v, lo, hi);
}
inline EFloat operator-(EFloat a) { return EFloat::with_bounds(-a.v, -a.high, -a.low); }  // :187-193
inline EFloat operator*(Float a, EFloat b) { return EFloat(a) * b; }                       // :211-217

// ---- vectors (cgmath) ------------------------------------------------------------------
struct Vec3 {
    Float x, y, z;
    Vec3() : x(0), y(0), z(0) {}
    Vec3(Float x_, Float y_, Float z_) : x(x_), y(y_), z(z_) {}
    Float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
    Float& operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }
};
typedef Vec3 Point3;
inline Vec3 operator+(Vec3 a, Vec3 b) { return Vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline Vec3 operator-(Vec3 a, Vec3 b) { return Vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline Vec3 operator-(Vec3 a) { return Vec3(-a.x, -a.y, -a.z); }
inline Vec3 operator*(Vec3 a, Float s) { return Vec3(a.x * s, a.y * s, a.z * s); }
inline Vec3 operator*(Float s, Vec3 a) { return Vec3(s * a.x, s * a.y, s * a.z); }
inline Vec3 operator/(Vec3 a, Float s) { return Vec3(a.x / s, a.y / s, a.z / s); }
inline Float dot(Vec3 a, Vec3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline Vec3 cross(Vec3 a, Vec3 b) {
    return Vec3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
inline Float magnitude2(Vec3 a) { return dot(a, a); }
inline Float magnitude(Vec3 a) { return std::sqrt(dot(a, a)); }
inline Vec3 normalize(Vec3 a) { return a * (1.0f / magnitude(a)); }
inline Vec3 vabs(Vec3 a) { return Vec3(std::fabs(a.x), std::fabs(a.y), std::fabs(a.z)); }
inline Float abs_dot(Vec3 a, Vec3 b) { return std::fabs(dot(a, b)); }  // math.rs:32

// geometry/mod.rs:45-51
inline int max_dimension(Vec3 v) {
    if (v.x > v.y) { return (v.x > v.z) ? 0 : 2; }
    else { return (v.y > v.z) ? 1 : 2; }
}
// geometry/mod.rs:53-62
inline void coordinate_system(Vec3 v1, Vec3* v2, Vec3* v3) {
    if (std::fabs(v1.x) > std::fabs(v1.y)) *v2 = normalize(Vec3(-v1.z, 0.0f, v1.x));
    else *v2 = normalize(Vec3(0.0f, v1.z, -v1.y));
    *v3 = cross(v1, *v2);
}
// geometry/mod.rs:64-70
inline Vec3 faceforward(Vec3 v1, Vec3 v2) { return (dot(v1, v2) < 0.0f) ? -v1 : v1; }
// geometry/mod.rs:23-34
inline Float spherical_theta(Vec3 v) { return std::acos(clampf(v.z, -1.0f, 1.0f)); }
inline Float spherical_phi(Vec3 v) {
    Float p = std::atan2(v.y, v.x);
    return (p < 0.0f) ? p + (2.0f * PI) : p;
}
// math.rs:74-80
inline Vec3 spherical_direction(Float sin_theta, Float cos_theta, Float phi) {
    return Vec3(sin_theta * std::cos(phi), sin_theta * std::sin(phi), cos_theta);
}

// geometry/mod.rs:72-85
inline Point3 offset_ray_origin(Point3 p, Vec3 p_err, Vec3 n, Vec3 dir) {
    Float d = dot(vabs(n), p_err);
    Vec3 offset = d * n;
    if (dot(dir, n) < 0.0f) offset = -offset;
    Point3 po = p + offset;
    for (int i = 0; i < 3; ++i) {
        if (offset[i] > 0.0f) po[i] = next_float_up(po[i]);
        else if (offset[i] < 0.0f) po[i] = next_float_down(po[i]);
    }
    return po;
}

// ---- Ray, geometry/mod.rs:87-106 ---------------------------------------------------------
struct Ray {
    Point3 origin; Vec3 dir; Float t_max; Float time;
    Point3 at(Float t) const { return origin + (dir * t); }
};

// ---- Matrix4 / Transform, geometry/transform.rs ------------------------------------------
struct Mat4 {
    Float m[4][4];  // m[c][r], column-major like cgmath
};
inline Mat4 mat_from_flat(const Float* f) {  // transform.rs:34-42 (Matrix4::new takes columns)
    Mat4 r;
    for (int c = 0; c < 4; ++c) for (int k = 0; k < 4; ++k) r.m[c][k] = f[4 * c + k];
    return r;
}
struct Transform { Mat4 t, invt; };

inline Point3 transform_point(const Mat4& M, Point3 p) {  // cgmath Matrix4::transform_point
    Float x = M.m[0][0] * p.x + M.m[1][0] * p.y + M.m[2][0] * p.z + M.m[3][0] * 1.0f;
    Float y = M.m[0][1] * p.x + M.m[1][1] * p.y + M.m[2][1] * p.z + M.m[3][1] * 1.0f;
    Float z = M.m[0][2] * p.x + M.m[1][2] * p.y + M.m[2][2] * p.z + M.m[3][2] * 1.0f;
    Float w = M.m[0][3] * p.x + M.m[1][3] * p.y + M.m[2][3] * p.z + M.m[3][3] * 1.0f;
    Float iw = 1.0f / w;
    return Point3(x * iw, y * iw, z * iw);
}
inline Vec3 transform_vector(const Mat4& M, Vec3 v) {  // cgmath Matrix4::transform_vector
    Float x = M.m[0][0] * v.x + M.m[1][0] * v.y + M.m[2][0] * v.z + M.m[3][0] * 0.0f;
    Float y = M.m[0][1] * v.x + M.m[1][1] * v.y + M.m[2][1] * v.z + M.m[3][1] * 0.0f;
    Float z = M.m[0][2] * v.x + M.m[1][2] * v.y + M.m[2][2] * v.z + M.m[3][2] * 0.0f;
    return Vec3(x, y, z);
}
inline Vec3 transform_normal(const Transform& T, Vec3 n) {  // transform.rs:134-140
    const Mat4& I = T.invt;
    Float x = I.m[0][0] * n.x + I.m[1][0] * n.y + I.m[2][0] * n.z;
    Float y = I.m[0][1] * n.x + I.m[1][1] * n.y + I.m[2][1] * n.z;
    Float z = I.m[0][2] * n.x + I.m[1][2] * n.y + I.m[2][2] * n.z;
    return Vec3(x, y, z);
}
inline Transform inverse(const Transform& T) { Transform r; r.t = T.invt; r.invt = T.t; return r; }

// Point3f::tf_exact_to_err, transform.rs:231-245
inline Point3 point_tf_exact_to_err(const Mat4& m, Point3 p, Vec3* err) {
    Point3 pt = transform_point(m, p);
    Float xs = std::fabs(m.m[0][0] * p.x) + std::fabs(m.m[1][0] * p.y) + std::fabs(m.m[2][0] * p.z) + std::fabs(m.m[3][0]);
    Float ys = std::fabs(m.m[0][1] * p.x) + std::fabs(m.m[1][1] * p.y) + std::fabs(m.m[2][1] * p.z) + std::fabs(m.m[3][1]);
    Float zs = std::fabs(m.m[0][2] * p.x) + std::fabs(m.m[1][2] * p.y) + std::fabs(m.m[2][2] * p.z) + std::fabs(m.m[3][2]);
    *err = Vec3(xs, ys, zs) * gamma(3);
    return pt;
}
// Point3f::tf_err_to_err, transform.rs:247-268
inline Point3 point_tf_err_to_err(const Mat4& m, Point3 p, Vec3 perr, Vec3* err) {
    Point3 pt = transform_point(m, p);
    Float g3 = gamma(3);
    Float xerr = (g3 + 1.0f) *
        (std::fabs(m.m[0][0]) * perr.x + std::fabs(m.m[1][0]) * perr.y + std::fabs(m.m[2][0]) * perr.z) +
        g3 * (std::fabs(m.m[0][0] * p.x) + std::fabs(m.m[1][0] * p.y) + std::fabs(m.m[2][0] * p.z) + std::fabs(m.m[3][0]));
    Float yerr = (g3 + 1.0f) *
        (std::fabs(m.m[0][1]) * perr.x + std::fabs(m.m[1][1]) * perr.y + std::fabs(m.m[2][1]) * perr.z) +
        g3 * (std::fabs(m.m[0][1] * p.x) + std::fabs(m.m[1][1] * p.y) + std::fabs(m.m[2][1] * p.z) + std::fabs(m.m[3][1]));
    Float zerr = (g3 + 1.0f) *
        (std::fabs(m.m[0][2] * perr.x) + std::fabs(m.m[1][2] * perr.y) + std::fabs(m.m[2][2] * perr.z)) +
        g3 * (std::fabs(m.m[0][2] * p.x) + std::fabs(m.m[1][2] * p.y) + std::fabs(m.m[2][2] * p.z) + std::fabs(m.m[3][2]));
    *err = Vec3(xerr, yerr, zerr);
    return pt;
}
// Vec3f::tf_exact_to_err, transform.rs:184-198
inline Vec3 vec_tf_exact_to_err(const Mat4& m, Vec3 v, Vec3* err) {
    Vec3 vt = transform_vector(m, v);
    Float xs = std::fabs(m.m[0][0] * v.x) + std::fabs(m.m[1][0] * v.y) + std::fabs(m.m[2][0] * v.z);
    Float ys = std::fabs(m.m[0][1] * v.x) + std::fabs(m.m[1][1] * v.y) + std::fabs(m.m[2][1] * v.z);
    Float zs = std::fabs(m.m[0][2] * v.x) + std::fabs(m.m[1][2] * v.y) + std::fabs(m.m[2][2] * v.z);
    *err = Vec3(xs, ys, zs) * gamma(3);
    return vt;
}
// Ray::tf_exact_to_err, transform.rs:287-303
inline Ray ray_tf_exact_to_err(const Mat4& m, const Ray& r, Vec3* o_err, Vec3* d_err) {
    Point3 ot = point_tf_exact_to_err(m, r.origin, o_err);
    Vec3 dt_ = vec_tf_exact_to_err(m, r.dir, d_err);
    Float tmax = r.t_max;
    Float len_sq = magnitude2(dt_);
    if (len_sq > 0.0f) {
        Float dt = dot(vabs(dt_), *o_err) / len_sq;
        ot = ot + dt_ * dt;
        tmax -= dt;
    }
    Ray out; out.origin = ot; out.dir = dt_; out.t_max = tmax; out.time = r.time;
    return out;
}
// Ray::transform, transform.rs:306-322
inline Ray ray_transform(const Mat4& m, const Ray& r) {
    Vec3 o_err;
    Point3 ot = point_tf_exact_to_err(m, r.origin, &o_err);
    Vec3 dir = transform_vector(m, r.dir);
    Float t_max = r.t_max;
    Float len_sq = magnitude2(dir);
    if (len_sq > 0.0f) {
        Float dt = dot(vabs(dir), o_err) / len_sq;
        ot = ot + dir * dt;
        t_max -= dt;
    }
    Ray out; out.origin = ot; out.dir = dir; out.t_max = t_max; out.time = r.time;
    return out;
}

// ---- Bounds3f, geometry/bounds.rs --------------------------------------------------------
struct Bounds3 {
    Point3 min, max;
    static Bounds3 empty() {  // bounds.rs:125-127: (f32::MAX, f32::MIN) from num::Bounded
        Float M = std::numeric_limits<float>::max();
        Bounds3 b; b.min = Point3(M, M, M); b.max = Point3(-M, -M, -M); return b;
    }
    Bounds3 join(const Bounds3& o) const {  // bounds.rs:129-143 (f32::min/max)
        Bounds3 b;
        b.min = Point3(fmin_(min.x, o.min.x), fmin_(min.y, o.min.y), fmin_(min.z, o.min.z));
        b.max = Point3(fmax_(max.x, o.max.x), fmax_(max.y, o.max.y), fmax_(max.z, o.max.z));
        return b;
    }
    Bounds3 join_point(Point3 p) const {  // bounds.rs:145-159
        Bounds3 b;
        b.min = Point3(fmin_(min.x, p.x), fmin_(min.y, p.y), fmin_(min.z, p.z));
        b.max = Point3(fmax_(max.x, p.x), fmax_(max.y, p.y), fmax_(max.z, p.z));
        return b;
    }
    Vec3 diagonal() const { return max - min; }                       // bounds.rs:165-167
    Point3 centroid() const { return min + (diagonal() / 2.0f); }     // bounds.rs:161-163
    int maximum_extent() const {                                      // bounds.rs:169-178
        Vec3 d = diagonal();
        if (d.x > d.y && d.x > d.z) return 0;
        else if (d.y > d.z) return 1;
        else return 2;
    }
    bool is_point() const { return max.x == min.x && max.y == min.y && max.z == min.z; }  // :180-182
    Vec3 offset(Point3 p) const {  // bounds.rs:200-206
        Vec3 o = p - min;
        if (max.x > min.x) o.x /= max.x - min.x;
        if (max.y > min.y) o.y /= max.y - min.y;
        if (max.z > min.z) o.z /= max.z - min.z;
        return o;
    }
    void bounding_sphere(Point3* center, Float* radius) const {  // bounds.rs:208-212
        *center = Point3(0, 0, 0) + ((min + max) / 2.0f);
        *radius = magnitude(max - *center);  // center.distance(max) = (max - center).magnitude()
    }
    // bounds.rs:214-233
    bool intersect_test(const Ray& ray, Float* t0_out = nullptr, Float* t1_out = nullptr) const {
        Float t0 = 0.0f, t1 = ray.t_max;
        for (int i = 0; i < 3; ++i) {
            Float inv_ray_dir = 1.0f / ray.dir[i];
            Float t_near = (min[i] - ray.origin[i]) * inv_ray_dir;
            Float t_far = (max[i] - ray.origin[i]) * inv_ray_dir;
            if (t_near > t_far) std::swap(t_near, t_far);
            t_far *= 1.0f + 2.0f * gamma(3);
            t0 = fmax_(t0, t_near);
            t1 = fmin_(t1, t_far);
            if (t0 > t1) return false;
        }
        if (t0_out) *t0_out = t0;
        if (t1_out) *t1_out = t1;
        return true;
    }
};
inline Bounds3 bounds_transform(const Mat4& m, const Bounds3& b) {  // transform.rs:276-283, corner order bounds.rs:184-196
    Bounds3 r = Bounds3::empty();
    for (int ix = 0; ix < 2; ++ix) for (int iy = 0; iy < 2; ++iy) for (int iz = 0; iz < 2; ++iz) {
        Point3 c(ix ? b.max.x : b.min.x, iy ? b.max.y : b.min.y, iz ? b.max.z : b.min.z);
        r = r.join_point(transform_point(m, c));
    }
    return r;
}

// ---- Spectrum (3 x f32 RGB), spectrum/mod.rs ----------------------------------------------
struct Spectrum {
    Float c[3];
    Spectrum() { c[0] = c[1] = c[2] = 0.0f; }
    explicit Spectrum(Float v) { c[0] = c[1] = c[2] = v; }
    Spectrum(Float r, Float g, Float b) { c[0] = r; c[1] = g; c[2] = b; }
    bool is_black() const { return c[0] == 0.0f && c[1] == 0.0f && c[2] == 0.0f; }   // :78-80
    bool has_nans() const { return std::isnan(c[0]) || std::isnan(c[1]) || std::isnan(c[2]); }
    Float max_component_value() const {  // :100-102, max_by(total_cmp): last maximum under the IEEE total order
        Float m = c[0];
        for (int i = 1; i < 3; ++i) {
            // total_cmp on finite non-negative throughput values reduces to >=
            int32_t a = (int32_t)f2u(c[i]), b = (int32_t)f2u(m);
            a ^= (int32_t)(((uint32_t)(a >> 31)) >> 1);
            b ^= (int32_t)(((uint32_t)(b >> 31)) >> 1);
            if (a >= b) m = c[i];
        }
        return m;
    }
    Float luminance() const { return c[0] * 0.212671f + c[1] * 0.715160f + c[2] * 0.072169f; }  // :104-107
    Spectrum clamp_positive() const { return Spectrum(clampf(c[0], 0.0f, INF), clampf(c[1], 0.0f, INF), clampf(c[2], 0.0f, INF)); }
};
inline Spectrum operator+(Spectrum a, Spectrum b) { return Spectrum(a.c[0] + b.c[0], a.c[1] + b.c[1], a.c[2] + b.c[2]); }
inline Spectrum operator-(Spectrum a, Spectrum b) { return Spectrum(a.c[0] - b.c[0], a.c[1] - b.c[1], a.c[2] - b.c[2]); }
inline Spectrum operator*(Spectrum a, Spectrum b) { return Spectrum(a.c[0] * b.c[0], a.c[1] * b.c[1], a.c[2] * b.c[2]); }
inline Spectrum operator/(Spectrum a, Spectrum b) { return Spectrum(a.c[0] / b.c[0], a.c[1] / b.c[1], a.c[2] / b.c[2]); }
inline Spectrum operator*(Spectrum a, Float s) { return Spectrum(a.c[0] * s, a.c[1] * s, a.c[2] * s); }
inline Spectrum operator*(Float s, Spectrum a) { return Spectrum(s * a.c[0], s * a.c[1], s * a.c[2]); }
inline Spectrum operator/(Spectrum a, Float s) { return Spectrum(a.c[0] / s, a.c[1] / s, a.c[2] / s); }
inline Spectrum operator+(Spectrum a, Float s) { return Spectrum(a.c[0] + s, a.c[1] + s, a.c[2] + s); }
inline Spectrum operator-(Spectrum a, Float s) { return Spectrum(a.c[0] - s, a.c[1] - s, a.c[2] - s); }
inline Spectrum operator-(Float s, Spectrum a) { return Spectrum(s - a.c[0], s - a.c[1], s - a.c[2]); }
inline Spectrum ssqrt(Spectrum a) { return Spectrum(std::sqrt(a.c[0]), std::sqrt(a.c[1]), std::sqrt(a.c[2])); }

// spectrum/mod.rs:28-43
inline void xyz_to_rgb(const Float xyz[3], Float rgb[3]) {
    rgb[0] = 3.240479f * xyz[0] - 1.537150f * xyz[1] - 0.498535f * xyz[2];
    rgb[1] = -0.969256f * xyz[0] + 1.875991f * xyz[1] + 0.041556f * xyz[2];
    rgb[2] = 0.055648f * xyz[0] - 0.204043f * xyz[1] + 1.057311f * xyz[2];
}
inline void rgb_to_xyz(const Float rgb[3], Float xyz[3]) {
    xyz[0] = 0.412453f * rgb[0] + 0.357580f * rgb[1] + 0.180423f * rgb[2];
    xyz[1] = 0.212671f * rgb[0] + 0.715160f * rgb[1] + 0.072169f * rgb[2];
    xyz[2] = 0.019334f * rgb[0] + 0.119193f * rgb[1] + 0.950227f * rgb[2];
}

// ---- morton.rs:3-36 ---------------------------------------------------------------------
inline uint32_t to_fixed_point(Float val) { return (uint32_t)std::trunc(val * 1024.0f); }
inline uint32_t expand_bits(uint32_t val) {
    val = (val * 0x00010001u) & 0xFF0000FFu;
    val = (val * 0x00000101u) & 0x0F00F00Fu;
    val = (val * 0x00000011u) & 0xC30C30C3u;
    val = (val * 0x00000005u) & 0x49249249u;
    return val;
}
inline uint32_t morton3(Float x, Float y, Float z) {
    uint32_t xx = expand_bits(to_fixed_point(x));
    uint32_t yy = expand_bits(to_fixed_point(y));
    uint32_t zz = expand_bits(to_fixed_point(z));
    return (xx << 2) | (yy << 1) | zz;
}

}  // namespace ref
