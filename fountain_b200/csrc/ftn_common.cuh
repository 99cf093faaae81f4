// fountain_b200 -- device/host math primitives for the sm_100a wavefront path tracer.
//
// Exact-arithmetic policy.  Everything that decides hit / miss / t / spawned ray origins
// must round exactly like the reference (akofke/fountain is Rust: every f32 op rounds once,
// a*b+c is never fused).  nvcc contracts a*b+c into FFMA by default, so those code paths
// use the rn_* helpers below (__fmul_rn/__fadd_rn/... are never contracted).  Shading code
// (BSDFs, Fresnel, env lookups) uses plain operators and may be contracted: its parity
// tolerance is stated in tests/ (transcendentals already differ from glibc at the ulp level).
//
// The functions are FTN_HD so that tests/hostsim/ can compile the very same code with g++
// and run it on the CPU against the oracle *as a test harness*.  That harness is not part of
// libfountain_gpu.so; the product has no CPU execution path.
#pragma once
#include <stdint.h>
#include <math.h>
#include <float.h>

#if defined(__CUDACC__)
#define FTN_HD __host__ __device__ __forceinline__
#define FTN_D __device__ __forceinline__
#ifndef FTN_HD_COLD
// side paths that only the kernel variants needing them instantiate (k_shade<.., IMG = true>); inlined there: a real
// call cost 6-9 % on the textured scene (profiles/r01_ab_vote_ldg256.txt), and the plain variants never see the code
#define FTN_HD_COLD static __host__ __device__ __forceinline__
#endif
#else
#define FTN_HD inline
#define FTN_D inline
#define FTN_HD_COLD inline
#include <string.h>
#endif

namespace ftn {

#if defined(__CUDA_ARCH__)
FTN_HD float rn_mul(float a, float b) { return __fmul_rn(a, b); }
FTN_HD float rn_add(float a, float b) { return __fadd_rn(a, b); }
FTN_HD float rn_sub(float a, float b) { return __fsub_rn(a, b); }
FTN_HD float rn_div(float a, float b) { return __fdiv_rn(a, b); }
FTN_HD float rn_sqrt(float a) { return __fsqrt_rn(a); }
FTN_HD double rn_dmul(double a, double b) { return __dmul_rn(a, b); }
FTN_HD double rn_dsub(double a, double b) { return __dsub_rn(a, b); }
FTN_HD uint32_t f2u(float f) { return __float_as_uint(f); }
FTN_HD float u2f(uint32_t u) { return __uint_as_float(u); }
#else
// host (hostsim harness, compiled with -ffp-contract=off)
FTN_HD float rn_mul(float a, float b) { return a * b; }
FTN_HD float rn_add(float a, float b) { return a + b; }
FTN_HD float rn_sub(float a, float b) { return a - b; }
FTN_HD float rn_div(float a, float b) { return a / b; }
FTN_HD float rn_sqrt(float a) { return sqrtf(a); }
FTN_HD double rn_dmul(double a, double b) { return a * b; }
FTN_HD double rn_dsub(double a, double b) { return a - b; }
FTN_HD uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
FTN_HD float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
#endif

#define FTN_PI 3.14159265358979323846f
#define FTN_INV_PI 0.318309886183790671538f
#define FTN_PI_2 1.57079632679489661923f
#define FTN_PI_4 0.785398163397448309616f
#define FTN_INF (u2f(0x7F800000u))

// err_float.rs:5-10 gamma(n) evaluated in f32; constants checked in tests against the oracle.
#define FTN_MACHINE_EPS 5.9604644775390625e-08f   /* f32::EPSILON * 0.5 = 2^-24 */
// constexpr: folded at compile time in IEEE f32 (each op rounds once; no a*b+c shape to contract),
// so no division is executed per ray / per triangle.
constexpr float gamma_c(int n) { return ((float)n * FTN_MACHINE_EPS) / (1.0f - (float)n * FTN_MACHINE_EPS); }
template <int N> struct GammaK { static constexpr float value = gamma_c(N); };
#define gamma_n(n) (ftn::GammaK<(n)>::value)

FTN_HD bool sign_positive(float f) { return (f2u(f) >> 31) == 0u; }   // f32::is_sign_positive

// Rust f32::clamp: plain comparisons, NaN passes through.
FTN_HD float clampf(float x, float lo, float hi) { float r = x; if (r < lo) r = lo; if (r > hi) r = hi; return r; }

// err_float.rs:12-30.  next_float_down(+-0) yields NaN in the reference (it tests `v >= 0.0`
// after mapping 0.0 to -0.0); kept bit-for-bit.
FTN_HD float next_float_up(float v) {
    if (v == FTN_INF) return v;
    if (v == -0.0f) v = 0.0f;
    uint32_t bits = f2u(v);
    bits = (v >= 0.0f) ? bits + 1u : bits - 1u;
    return u2f(bits);
}
FTN_HD float next_float_down(float v) {
    if (v == -FTN_INF) return v;
    if (v == 0.0f) v = -0.0f;
    uint32_t bits = f2u(v);
    bits = (v >= 0.0f) ? bits - 1u : bits + 1u;
    return u2f(bits);
}

struct V3 {
    float x, y, z;
    FTN_HD V3() {}
    FTN_HD V3(float a, float b, float c) : x(a), y(b), z(c) {}
    FTN_HD float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};
FTN_HD V3 v3(float a, float b, float c) { return V3(a, b, c); }
FTN_HD V3 v3s(float a) { return V3(a, a, a); }

// ---- exactly-rounded vector ops (cgmath operation order, see oracle/ref_math.h) -------------
FTN_HD V3 x_add(V3 a, V3 b) { return V3(rn_add(a.x, b.x), rn_add(a.y, b.y), rn_add(a.z, b.z)); }
FTN_HD V3 x_sub(V3 a, V3 b) { return V3(rn_sub(a.x, b.x), rn_sub(a.y, b.y), rn_sub(a.z, b.z)); }
FTN_HD V3 x_scale(V3 a, float s) { return V3(rn_mul(a.x, s), rn_mul(a.y, s), rn_mul(a.z, s)); }
FTN_HD V3 x_neg(V3 a) { return V3(-a.x, -a.y, -a.z); }
FTN_HD V3 x_abs(V3 a) { return V3(fabsf(a.x), fabsf(a.y), fabsf(a.z)); }
FTN_HD float x_dot(V3 a, V3 b) { return rn_add(rn_add(rn_mul(a.x, b.x), rn_mul(a.y, b.y)), rn_mul(a.z, b.z)); }
FTN_HD V3 x_cross(V3 a, V3 b) {
    return V3(rn_sub(rn_mul(a.y, b.z), rn_mul(a.z, b.y)),
              rn_sub(rn_mul(a.z, b.x), rn_mul(a.x, b.z)),
              rn_sub(rn_mul(a.x, b.y), rn_mul(a.y, b.x)));
}
FTN_HD float x_len2(V3 a) { return x_dot(a, a); }
FTN_HD float x_len(V3 a) { return rn_sqrt(x_dot(a, a)); }
FTN_HD V3 x_normalize(V3 a) { return x_scale(a, rn_div(1.0f, x_len(a))); }   // v * (1 / |v|)
FTN_HD float x_abs_dot(V3 a, V3 b) { return fabsf(x_dot(a, b)); }

// geometry/mod.rs:45-51
FTN_HD int max_dimension(V3 v) {
    if (v.x > v.y) { return (v.x > v.z) ? 0 : 2; }
    else { return (v.y > v.z) ? 1 : 2; }
}
// geometry/mod.rs:53-62
FTN_HD void coordinate_system(V3 v1, V3* v2, V3* v3_) {
    if (fabsf(v1.x) > fabsf(v1.y)) *v2 = x_normalize(V3(-v1.z, 0.0f, v1.x));
    else *v2 = x_normalize(V3(0.0f, v1.z, -v1.y));
    *v3_ = x_cross(v1, *v2);
}
// geometry/mod.rs:64-70
FTN_HD V3 faceforward(V3 v1, V3 v2) { return (x_dot(v1, v2) < 0.0f) ? x_neg(v1) : v1; }

// geometry/mod.rs:72-85
FTN_HD V3 offset_ray_origin(V3 p, V3 p_err, V3 n, V3 dir) {
    float d = x_dot(x_abs(n), p_err);
    V3 off = V3(rn_mul(d, n.x), rn_mul(d, n.y), rn_mul(d, n.z));
    if (x_dot(dir, n) < 0.0f) off = x_neg(off);
    V3 po = x_add(p, off);
    if (off.x > 0.0f) po.x = next_float_up(po.x); else if (off.x < 0.0f) po.x = next_float_down(po.x);
    if (off.y > 0.0f) po.y = next_float_up(po.y); else if (off.y < 0.0f) po.y = next_float_down(po.y);
    if (off.z > 0.0f) po.z = next_float_up(po.z); else if (off.z < 0.0f) po.z = next_float_down(po.z);
    return po;
}

// ---- morton.rs:3-36 (+ the 1023 clamp documented in oracle_capi.cpp) ---------------------------
FTN_HD uint32_t expand_bits(uint32_t v) {
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}
FTN_HD uint32_t to_fixed_point_clamped(float v) {
    float s = truncf(rn_mul(v, 1024.0f));
    uint32_t u = (s > 0.0f) ? (uint32_t)s : 0u;   // Rust `as u32` saturates
    return u > 1023u ? 1023u : u;
}
FTN_HD uint32_t morton3_clamped(float x, float y, float z) {
    return (expand_bits(to_fixed_point_clamped(x)) << 2) | (expand_bits(to_fixed_point_clamped(y)) << 1) |
           expand_bits(to_fixed_point_clamped(z));
}

// ---- counter-based sampler (FTN_SAMPLER_COUNTER; same definition in oracle/ref_render.h) -------
FTN_HD uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
FTN_HD uint64_t sampler_seed_key(uint64_t seed) { return mix64(seed * 0x9E3779B97F4A7C15ull + 0xD1B54A32D192ED03ull); }
FTN_HD uint64_t sampler_sample_key(uint64_t seed_key, uint64_t sample_index) { return seed_key ^ (sample_index * 0x9E3779B97F4A7C15ull); }
FTN_HD float sampler_uniform(uint64_t sample_key, uint32_t dim) {
    uint64_t z = mix64(sample_key + (uint64_t)dim * 0xC2B2AE3D27D4EB4Full);
    return (float)((uint32_t)(z >> 32) >> 8) * (1.0f / 16777216.0f);   // (u32 >> 8) * 2^-24, as rand 0.6.5
}
enum { DIM_CAMERA = 5, DIM_PER_BOUNCE = 8 };

// ---- 4x4 column-major matrices (cgmath layout: m[4*c + r]) ---------------------------------------
struct M4 { float m[16]; };
FTN_HD float m4(const M4& M, int c, int r) { return M.m[4 * c + r]; }
// cgmath Matrix4::transform_point: (M * (p,1)).xyz * (1/w)
FTN_HD V3 transform_point(const M4& M, V3 p) {
    float x = rn_add(rn_add(rn_add(rn_mul(m4(M, 0, 0), p.x), rn_mul(m4(M, 1, 0), p.y)), rn_mul(m4(M, 2, 0), p.z)), m4(M, 3, 0));
    float y = rn_add(rn_add(rn_add(rn_mul(m4(M, 0, 1), p.x), rn_mul(m4(M, 1, 1), p.y)), rn_mul(m4(M, 2, 1), p.z)), m4(M, 3, 1));
    float z = rn_add(rn_add(rn_add(rn_mul(m4(M, 0, 2), p.x), rn_mul(m4(M, 1, 2), p.y)), rn_mul(m4(M, 2, 2), p.z)), m4(M, 3, 2));
    float w = rn_add(rn_add(rn_add(rn_mul(m4(M, 0, 3), p.x), rn_mul(m4(M, 1, 3), p.y)), rn_mul(m4(M, 2, 3), p.z)), m4(M, 3, 3));
    float iw = rn_div(1.0f, w);
    return V3(rn_mul(x, iw), rn_mul(y, iw), rn_mul(z, iw));
}
// cgmath Matrix4::transform_vector: + M[3]*0 (exact no-op for finite M)
FTN_HD V3 transform_vector(const M4& M, V3 v) {
    float x = rn_add(rn_add(rn_add(rn_mul(m4(M, 0, 0), v.x), rn_mul(m4(M, 1, 0), v.y)), rn_mul(m4(M, 2, 0), v.z)), rn_mul(m4(M, 3, 0), 0.0f));
    float y = rn_add(rn_add(rn_add(rn_mul(m4(M, 0, 1), v.x), rn_mul(m4(M, 1, 1), v.y)), rn_mul(m4(M, 2, 1), v.z)), rn_mul(m4(M, 3, 1), 0.0f));
    float z = rn_add(rn_add(rn_add(rn_mul(m4(M, 0, 2), v.x), rn_mul(m4(M, 1, 2), v.y)), rn_mul(m4(M, 2, 2), v.z)), rn_mul(m4(M, 3, 2), 0.0f));
    return V3(x, y, z);
}
// transform.rs:134-140: normals by the transpose of the inverse (pass the INVERSE matrix)
FTN_HD V3 transform_normal_inv(const M4& I, V3 n) {
    float x = rn_add(rn_add(rn_mul(m4(I, 0, 0), n.x), rn_mul(m4(I, 1, 0), n.y)), rn_mul(m4(I, 2, 0), n.z));
    float y = rn_add(rn_add(rn_mul(m4(I, 0, 1), n.x), rn_mul(m4(I, 1, 1), n.y)), rn_mul(m4(I, 2, 1), n.z));
    float z = rn_add(rn_add(rn_mul(m4(I, 0, 2), n.x), rn_mul(m4(I, 1, 2), n.y)), rn_mul(m4(I, 2, 2), n.z));
    return V3(x, y, z);
}
// Point3f::tf_exact_to_err, transform.rs:231-245
FTN_HD V3 point_tf_exact_to_err(const M4& m, V3 p, V3* err) {
    V3 pt = transform_point(m, p);
    float g3 = gamma_n(3);
    float xs = rn_add(rn_add(rn_add(fabsf(rn_mul(m4(m, 0, 0), p.x)), fabsf(rn_mul(m4(m, 1, 0), p.y))), fabsf(rn_mul(m4(m, 2, 0), p.z))), fabsf(m4(m, 3, 0)));
    float ys = rn_add(rn_add(rn_add(fabsf(rn_mul(m4(m, 0, 1), p.x)), fabsf(rn_mul(m4(m, 1, 1), p.y))), fabsf(rn_mul(m4(m, 2, 1), p.z))), fabsf(m4(m, 3, 1)));
    float zs = rn_add(rn_add(rn_add(fabsf(rn_mul(m4(m, 0, 2), p.x)), fabsf(rn_mul(m4(m, 1, 2), p.y))), fabsf(rn_mul(m4(m, 2, 2), p.z))), fabsf(m4(m, 3, 2)));
    *err = V3(rn_mul(xs, g3), rn_mul(ys, g3), rn_mul(zs, g3));
    return pt;
}
// Point3f::tf_err_to_err, transform.rs:247-268
FTN_HD V3 point_tf_err_to_err(const M4& m, V3 p, V3 pe, V3* err) {
    V3 pt = transform_point(m, p);
    float g3 = gamma_n(3), g31 = rn_add(g3, 1.0f);
    float out[3];
    for (int r = 0; r < 3; ++r) {
        float a = rn_add(rn_add(rn_mul(fabsf(m4(m, 0, r)), pe.x), rn_mul(fabsf(m4(m, 1, r)), pe.y)), rn_mul(fabsf(m4(m, 2, r)), pe.z));
        float b = rn_add(rn_add(rn_add(fabsf(rn_mul(m4(m, 0, r), p.x)), fabsf(rn_mul(m4(m, 1, r), p.y))), fabsf(rn_mul(m4(m, 2, r), p.z))), fabsf(m4(m, 3, r)));
        out[r] = rn_add(rn_mul(g31, a), rn_mul(g3, b));
    }
    *err = V3(out[0], out[1], out[2]);
    return pt;
}
// Vec3f::tf_exact_to_err, transform.rs:184-198
FTN_HD V3 vec_tf_exact_to_err(const M4& m, V3 v, V3* err) {
    V3 vt = transform_vector(m, v);
    float g3 = gamma_n(3);
    float xs = rn_add(rn_add(fabsf(rn_mul(m4(m, 0, 0), v.x)), fabsf(rn_mul(m4(m, 1, 0), v.y))), fabsf(rn_mul(m4(m, 2, 0), v.z)));
    float ys = rn_add(rn_add(fabsf(rn_mul(m4(m, 0, 1), v.x)), fabsf(rn_mul(m4(m, 1, 1), v.y))), fabsf(rn_mul(m4(m, 2, 1), v.z)));
    float zs = rn_add(rn_add(fabsf(rn_mul(m4(m, 0, 2), v.x)), fabsf(rn_mul(m4(m, 1, 2), v.y))), fabsf(rn_mul(m4(m, 2, 2), v.z)));
    *err = V3(rn_mul(xs, g3), rn_mul(ys, g3), rn_mul(zs, g3));
    return vt;
}

struct RayF { V3 o, d; float t_max, time; };

// Ray::transform, transform.rs:306-322 (camera) and Ray::tf_exact_to_err :287-303 (sphere): the
// same arithmetic; the latter also returns the direction error.
FTN_HD RayF ray_transform_err(const M4& m, const RayF& r, V3* o_err, V3* d_err) {
    V3 ot = point_tf_exact_to_err(m, r.o, o_err);
    V3 dt_ = vec_tf_exact_to_err(m, r.d, d_err);
    float tmax = r.t_max;
    float len_sq = x_len2(dt_);
    if (len_sq > 0.0f) {
        float dt = rn_div(x_dot(x_abs(dt_), *o_err), len_sq);
        ot = x_add(ot, x_scale(dt_, dt));
        tmax = rn_sub(tmax, dt);
    }
    RayF out; out.o = ot; out.d = dt_; out.t_max = tmax; out.time = r.time;
    return out;
}

}  // namespace ftn
