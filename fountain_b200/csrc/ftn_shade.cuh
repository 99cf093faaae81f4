// Shading-side device functions: sampling routines, Lambert / Trowbridge-Reitz lobes, Fresnel,
// the Bsdf frame, the infinite (env-map) and sphere area lights, the thin-lens camera.
// Reference rows: SURVEY.md 8a a10-a18.  Geometry that feeds ray origins uses the exact ops of
// ftn_common.cuh; BSDF / Fresnel / pdf arithmetic uses plain operators (FMA contraction allowed,
// tolerance stated in tests/test_gpu_render.py).
#pragma once
#include "ftn_scene.h"
#include "ftn_geom.cuh"

namespace ftn {

// ---- tiny RGB helper -------------------------------------------------------------------------------
FTN_HD V3 operator+(V3 a, V3 b) { return V3(a.x + b.x, a.y + b.y, a.z + b.z); }
FTN_HD V3 operator-(V3 a, V3 b) { return V3(a.x - b.x, a.y - b.y, a.z - b.z); }
FTN_HD V3 operator*(V3 a, V3 b) { return V3(a.x * b.x, a.y * b.y, a.z * b.z); }
FTN_HD V3 operator/(V3 a, V3 b) { return V3(a.x / b.x, a.y / b.y, a.z / b.z); }
FTN_HD V3 operator*(V3 a, float s) { return V3(a.x * s, a.y * s, a.z * s); }
FTN_HD V3 operator*(float s, V3 a) { return V3(s * a.x, s * a.y, s * a.z); }
FTN_HD V3 operator/(V3 a, float s) { return V3(a.x / s, a.y / s, a.z / s); }
FTN_HD V3 operator+(V3 a, float s) { return V3(a.x + s, a.y + s, a.z + s); }
FTN_HD V3 operator-(V3 a, float s) { return V3(a.x - s, a.y - s, a.z - s); }
FTN_HD V3 operator-(V3 a) { return V3(-a.x, -a.y, -a.z); }
FTN_HD V3 vsqrt(V3 a) { return V3(sqrtf(a.x), sqrtf(a.y), sqrtf(a.z)); }
FTN_HD bool is_black(V3 a) { return a.x == 0.0f && a.y == 0.0f && a.z == 0.0f; }          // spectrum/mod.rs:78-80
FTN_HD bool has_nans(V3 a) { return a.x != a.x || a.y != a.y || a.z != a.z; }
FTN_HD float max_component(V3 a) { return fmaxf(fmaxf(a.x, a.y), a.z); }                     // :100-102 (finite inputs)
FTN_HD float dot(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
FTN_HD V3 cross(V3 a, V3 b) { return V3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
FTN_HD V3 normalize(V3 a) { return a * (1.0f / sqrtf(dot(a, a))); }
FTN_HD float abs_dot(V3 a, V3 b) { return fabsf(dot(a, b)); }

// sin / cos of the SAMPLING code (directions drawn from random numbers; compared with the oracle under the image
// tolerance, not bit for bit): the SFU forms on the device (abs error ~2^-21 on [0, 2 pi]), libm on the host harness
#if defined(__CUDA_ARCH__) && !defined(FTN_SHADE_PRECISE)
FTN_HD void f_sincos(float x, float* s, float* c) { __sincosf(x, s, c); }
FTN_HD float f_sin(float x) { return __sinf(x); }
#else
FTN_HD void f_sincos(float x, float* s, float* c) { *s = sinf(x); *c = cosf(x); }
FTN_HD float f_sin(float x) { return sinf(x); }
#endif

// ---- sampling.rs -------------------------------------------------------------------------------------
FTN_HD void concentric_sample_disk(float u0, float u1, float* dx, float* dy) {   // :5-19
    const float ox = 2.0f * u0 - 1.0f, oy = 2.0f * u1 - 1.0f;
    if (ox == 0.0f && oy == 0.0f) { *dx = 0.0f; *dy = 0.0f; return; }
    float theta, r;
    if (fabsf(ox) > fabsf(oy)) { theta = FTN_PI_4 * (oy / ox); r = ox; }
    else { theta = FTN_PI_2 - FTN_PI_4 * (ox / oy); r = oy; }
    float st, ct; f_sincos(theta, &st, &ct);
    *dx = r * ct; *dy = r * st;
}
FTN_HD V3 cosine_sample_hemisphere(float u0, float u1) {   // :21-25
    float dx, dy; concentric_sample_disk(u0, u1, &dx, &dy);
    return V3(dx, dy, sqrtf(fmaxf(0.0f, 1.0f - dx * dx - dy * dy)));
}
FTN_HD float power_heuristic1(float f, float g) { return (f * f) / (f * f + g * g); }   // :53-57 with nf = ng = 1

// ---- level-0 MIPMap lookup, mipmap.rs:245-312, ImageWrap::Repeat -------------------------------------
FTN_HD V3 env_texel(const EnvLightData& e, int s, int t) {
    // Euclidean s mod w.  Lookups arrive with s in [-1, w] (one texel beyond the image): one conditional add
    // instead of an integer division; anything further out takes the general form.
    int ss, tt;
    if (s >= -e.w && s < 2 * e.w) ss = s < 0 ? s + e.w : (s >= e.w ? s - e.w : s);
    else { ss = s % e.w; if (ss < 0) ss += e.w; }
    if (t >= -e.h && t < 2 * e.h) tt = t < 0 ? t + e.h : (t >= e.h ? t - e.h : t);
    else { tt = t % e.h; if (tt < 0) tt += e.h; }
    const F4 v = ld4(e.texels + (size_t)tt * e.w + ss);
    return V3(v.x, v.y, v.z);
}
FTN_HD V3 env_triangle0(const EnvLightData& e, float sx, float ty) {   // :265-279
    const float s = sx * (float)e.w - 0.5f, t = ty * (float)e.h - 0.5f;
    const float fs = floorf(s), ft = floorf(t);
    const int s0 = (int)fs, t0 = (int)ft;
    const float ds = s - fs, dt = t - ft;
    return env_texel(e, s0, t0) * (1.0f - ds) * (1.0f - dt) + env_texel(e, s0, t0 + 1) * (1.0f - ds) * dt
         + env_texel(e, s0 + 1, t0) * ds * (1.0f - dt) + env_texel(e, s0 + 1, t0 + 1) * ds * dt;
}
// lookup_trilinear_width for the widths the env light uses (0 and 1/max(w,h)): level 0 only
// (SURVEY section 5 note 1); a 1x1 map with width = 1 returns the texel itself.
FTN_HD V3 env_lookup_width(const EnvLightData& e, float sx, float ty, float width) {
    const float level = (float)e.levels - 1.0f + log2f(fmaxf(width, 1.0e-8f));
    if (level < 0.0f) return env_triangle0(e, sx, ty);
    if (level >= (float)(e.levels - 1)) return env_texel(e, 0, 0);
    return env_triangle0(e, sx, ty);
}

// InfiniteAreaLight::compute_distribution, infinite.rs:63-77:
//   func[j * width + i] = luminance(lookup(u = i/width, v = j/height, 1/max)) * sin(pi (j+.5)/height)
// with (height, width) = (map.w, map.h) exactly as the reference destructures them (nu = width).
FTN_HD float env_func_value(const EnvLightData& env, int k) {
    const int width = env.nu, height = env.nv;
    const int i = k % width, j = k / width;
    const float v = (float)j / (float)height;
    const float sin_theta = sinf(FTN_PI * ((float)j + 0.5f) / (float)height);
    const float u = (float)i / (float)width;
    const float filter = 1.0f / (float)(width > height ? width : height);
    const V3 c = env_lookup_width(env, u, v, filter);
    return (c.x * 0.212671f + c.y * 0.715160f + c.z * 0.072169f) * sin_theta;   // Spectrum::luminance, spectrum/mod.rs:104-107
}
// Distribution1D::new for one row (sampling.rs:84-107): the running f32 sum is sequential by
// definition, so one thread owns one row.
FTN_HD void dist_row_build(const float* f, int nu, float* c, float* integral) {
    float run = 0.0f;
    c[0] = 0.0f;
    const float nf = (float)nu;
    for (int i = 1; i <= nu; ++i) { run = rn_add(run, rn_div(f[i - 1], nf)); c[i] = run; }
    const float total = run;
    *integral = total;
    if (total == 0.0f) { for (int i = 1; i <= nu; ++i) c[i] = rn_div((float)i, nf); }
    else { for (int i = 1; i <= nu; ++i) c[i] = rn_div(c[i], total); }
}

// sampling.rs:66-81 over a device array: index of the last cdf entry <= u, clamped to [0, size-2]
// the partition point of the search: number of entries of cdf[first .. first + len) that are <= u, plus first
FTN_HD int search_partition(const float* cdf, int first, int len, float u) {
    while (len > 0) {
        const int half = len >> 1, middle = first + half;
        if (cdf[middle] <= u) { first = middle + 1; len -= half + 1; }
        else len = half;
    }
    return first;
}
FTN_HD int search_sorted_le(const float* cdf, int size, float u) {
    int r = search_partition(cdf, 0, size, u) - 1;
    if (r < 0) r = 0;
    if (r > size - 2) r = size - 2;
    return r;
}
// Guide table of one cdf row (n + 1 entries): entry g in [0, n] = partition point of the key g / n.
FTN_HD uint32_t env_guide_entry(const float* cdf, int n, int g) { return (uint32_t)search_partition(cdf, 0, n + 1, (float)g / (float)n); }
// search_sorted_le through the guide: floor(u n) is within one of the true bucket whatever the rounding of u * n, and the
// partition point is monotone in u, so it lies between the guide entries of buckets g - 1 and g + 2.
FTN_HD int search_sorted_le_guided(const float* cdf, const uint32_t* guide, int n, float u) {
    const float gf = u * (float)n;
    int g = (gf > 0.0f) ? (int)fminf(gf, 2.0e9f) : 0;
    if (g > n - 1) g = n - 1;
    const int lo = (int)guide[g > 0 ? g - 1 : 0], hi = (int)guide[g + 2 < n ? g + 2 : n];
    int r = search_partition(cdf, lo, hi - lo, u) - 1;
    if (r < 0) r = 0;
    if (r > n - 1) r = n - 1;
    return r;
}
// Distribution1D::sample_continuous, sampling.rs:121-134
FTN_HD void dist1d_sample(const float* func, const float* cdf, int n, float integral, float u, float* x, float* pdf, int* idx, const uint32_t* guide = nullptr) {
    const int i = guide ? search_sorted_le_guided(cdf, guide, n, u) : search_sorted_le(cdf, n + 1, u);
    float du = u - cdf[i];
    const float w = cdf[i + 1] - cdf[i];
    if (w > 0.0f) du /= w;
    *pdf = func[i] / integral;
    *x = ((float)i + du) / (float)n;
    *idx = i;
}
// Distribution2D::sample_continuous / pdf, sampling.rs:163-179
FTN_HD void env_dist_sample(const EnvLightData& e, float u0, float u1, float* d0, float* d1, float* pdf) {
    float pdf1, pdf0; int v, dummy;
    dist1d_sample(e.cond_integral, e.marg_cdf, e.nv, e.marg_integral, u1, d1, &pdf1, &v, e.marg_guide);
    dist1d_sample(e.cond_func + (size_t)v * e.nu, e.cond_cdf + (size_t)v * (e.nu + 1), e.nu, e.cond_integral[v], u0, d0, &pdf0, &dummy,
                  e.cond_guide ? e.cond_guide + (size_t)v * (e.nu + 1) : nullptr);
    *pdf = pdf0 * pdf1;
}
FTN_HD float env_dist_pdf(const EnvLightData& e, float px, float py) {
    const float fu = px * (float)e.nu, fv = py * (float)e.nv;
    int iu = (fu > 0.0f) ? (int)fminf(fu, 2.0e9f) : 0; if (iu > e.nu - 1) iu = e.nu - 1;
    int iv = (fv > 0.0f) ? (int)fminf(fv, 2.0e9f) : 0; if (iv > e.nv - 1) iv = e.nv - 1;
    return e.cond_func[(size_t)iv * e.nu + iu] / e.marg_integral;
}

FTN_HD float spherical_theta(V3 v) { return acosf(clampf(v.z, -1.0f, 1.0f)); }   // geometry/mod.rs:23-25
FTN_HD float spherical_phi(V3 v) { const float p = atan2f(v.y, v.x); return (p < 0.0f) ? p + (2.0f * FTN_PI) : p; }   // :27-34

// InfiniteAreaLight::environment_emitted_radiance, infinite.rs:156-164
FTN_HD V3 env_emitted(const EnvLightData& e, V3 dir) {
    const V3 w = normalize(transform_vector(e.w2l, dir));
    return env_lookup_width(e, spherical_phi(w) * (1.0f / (2.0f * FTN_PI)), spherical_theta(w) * FTN_INV_PI, 0.0f);
}
// InfiniteAreaLight::pdf_incident_radiance, infinite.rs:142-154
FTN_HD float env_pdf(const EnvLightData& e, V3 wi_w) {
    const V3 wi = transform_vector(e.w2l, wi_w);
    const float theta = spherical_theta(wi), phi = spherical_phi(wi);
    const float st = sinf(theta);
    if (st == 0.0f) return 0.0f;
    return env_dist_pdf(e, phi * (1.0f / (2.0f * FTN_PI)), theta * FTN_INV_PI) / (2.0f * FTN_PI * FTN_PI * st);
}
// InfiniteAreaLight::sample_incident_radiance, infinite.rs:99-140.  Returns false where the
// reference hits unimplemented!() (map_pdf == 0).
FTN_HD bool env_sample(const EnvLightData& e, float u0, float u1, V3* wi, float* pdf, V3* radiance) {
    float uvx, uvy, map_pdf;
    env_dist_sample(e, u0, u1, &uvx, &uvy, &map_pdf);
    if (map_pdf == 0.0f) return false;
    const float theta = uvy * FTN_PI, phi = uvx * 2.0f * FTN_PI;
    float st, ct, sp, cp; f_sincos(theta, &st, &ct); f_sincos(phi, &sp, &cp);
    *wi = transform_vector(e.l2w, V3(st * cp, st * sp, ct));
    *pdf = (st == 0.0f) ? 0.0f : map_pdf / (2.0f * FTN_PI * FTN_PI * st);
    *radiance = env_lookup_width(e, uvx, uvy, 0.0f);
    return true;
}

// ---- fresnel.rs ------------------------------------------------------------------------------------------
FTN_HD float fresnel_dielectric(float cos_i, float eta_i, float eta_t) {   // :4-22
    cos_i = clampf(cos_i, -1.0f, 1.0f);
    if (!(cos_i > 0.0f)) { const float t = eta_i; eta_i = eta_t; eta_t = t; cos_i = fabsf(cos_i); }
    const float sin_i = sqrtf(fmaxf(1.0f - cos_i * cos_i, 0.0f));
    const float sin_t = eta_i / eta_t * sin_i;
    if (sin_t >= 1.0f) return 1.0f;
    const float cos_t = sqrtf(fmaxf(1.0f - sin_t * sin_t, 0.0f));
    const float rpar = ((eta_t * cos_i) - (eta_i * cos_t)) / ((eta_t * cos_i) + (eta_i * cos_t));
    const float rper = ((eta_i * cos_i) - (eta_t * cos_t)) / ((eta_i * cos_i) + (eta_t * cos_t));
    return (rpar * rpar + rper * rper) / 2.0f;
}
FTN_HD V3 fresnel_conductor(float cos_i, V3 eta_t, V3 k) {   // :25-48 with eta_i = 1 (metal.rs:54)
    cos_i = clampf(cos_i, -1.0f, 1.0f);
    const V3 one = v3s(1.0f);
    const V3 eta = eta_t / one, eta_k = k / one;
    const float cos2 = cos_i * cos_i, sin2 = 1.0f - cos2;
    const V3 eta2 = eta * eta, eta_k2 = eta_k * eta_k;
    const V3 t0 = eta2 - eta_k2 - sin2;
    const V3 a2plusb2 = vsqrt(t0 * t0 + 4.0f * eta2 * eta_k2);
    const V3 t1 = a2plusb2 + cos2;
    const V3 a = vsqrt(0.5f * (a2plusb2 + t0));
    const V3 t2 = (2.0f * cos_i) * a;
    const V3 Rs = (t1 - t2) / (t1 + t2);
    const V3 t3 = cos2 * a2plusb2 + sin2 * sin2;
    const V3 t4 = t2 * sin2;
    const V3 Rp = Rs * (t3 - t4) / (t3 + t4);
    return 0.5f * (Rp + Rs);
}

// ---- reflection/mod.rs trig helpers :24-86 -------------------------------------------------------------------
FTN_HD float cos2_theta(V3 w) { return w.z * w.z; }
FTN_HD float abs_cos_theta(V3 w) { return fabsf(w.z); }
FTN_HD float sin2_theta(V3 w) { return fmaxf(0.0f, 1.0f - cos2_theta(w)); }
FTN_HD float sin_theta(V3 w) { return sqrtf(sin2_theta(w)); }
FTN_HD float tan_theta(V3 w) { return sin_theta(w) / w.z; }
FTN_HD float tan2_theta(V3 w) { return sin2_theta(w) / cos2_theta(w); }
FTN_HD float cos_phi(V3 w) { const float s = sin_theta(w); return (s == 0.0f) ? 1.0f : clampf(w.x / s, -1.0f, 1.0f); }
FTN_HD float sin_phi(V3 w) { const float s = sin_theta(w); return (s == 0.0f) ? 0.0f : clampf(w.y / s, -1.0f, 1.0f); }
FTN_HD bool same_hemisphere(V3 a, V3 b) { return sign_positive(a.z) == sign_positive(b.z); }
FTN_HD V3 reflect(V3 wo, V3 n) { return -wo + 2.0f * dot(wo, n) * n; }
FTN_HD bool is_inf(float f) { return fabsf(f) == FTN_INF; }

enum { BXDF_REFLECTION = 1, BXDF_TRANSMISSION = 2, BXDF_DIFFUSE = 4, BXDF_GLOSSY = 8, BXDF_SPECULAR = 16, BXDF_ALL = 31 };

// One lobe of a Bsdf (bsdf.rs holds up to 8 `dyn BxDF`; the in-scope materials produce at most 2).
// KIND is a compile-time property of (material class, lobe slot): 0 LambertianReflection,
// 1 MicrofacetReflection<TrowbridgeReitz, F>, 2 SpecularReflection<FresnelNoOp> (mirror), 3 OrenNayar
// (a, b ride in ax, ay), -1 no such lobe.  Shading runs one kernel launch per
// material class over its queue, so the class is a template argument and the matte kernel
// contains no microfacet / Fresnel code at all (it used 167 registers when the lobe kind was a
// run-time field).
struct Lobe {
    V3 r;
    float ax, ay;
    V3 eta, k;
};
template <int MAT, int I> struct LobeKind { static constexpr int value = -1; };
template <> struct LobeKind<FTN_MATERIAL_MATTE, 0> { static constexpr int value = 0; };
template <> struct LobeKind<FTN_MATERIAL_METAL, 0> { static constexpr int value = 1; };
template <> struct LobeKind<FTN_MATERIAL_PLASTIC, 0> { static constexpr int value = 0; };
template <> struct LobeKind<FTN_MATERIAL_PLASTIC, 1> { static constexpr int value = 1; };
template <> struct LobeKind<FTN_MATERIAL_MIRROR, 0> { static constexpr int value = 2; };
template <> struct LobeKind<FTN_CLASS_OREN_NAYAR, 0> { static constexpr int value = 3; };
template <> struct LobeKind<FTN_MATERIAL_GLASS, 0> { static constexpr int value = 1; };   // MicrofacetReflection (Kr), glass.rs:76-79
template <> struct LobeKind<FTN_MATERIAL_GLASS, 1> { static constexpr int value = 4; };   // MicrofacetTransmission (Kt), glass.rs:88-91
// Fresnel of the microfacet lobe: 0 FresnelConductor{1, eta, k} (metal.rs:52-56), 1 FresnelDielectric{1.5, 1.0} (plastic.rs:34),
// 2 FresnelDielectric{1, eta} with the material's index in Lobe::eta.x (glass.rs:70)
template <int MAT> struct LobeFresnel { static constexpr int value = (MAT == FTN_MATERIAL_PLASTIC) ? 1 : (MAT == FTN_MATERIAL_GLASS) ? 2 : 0; };

template <int KIND> FTN_HD constexpr int lobe_type() {
    return (KIND == 0 || KIND == 3) ? (BXDF_REFLECTION | BXDF_DIFFUSE) : KIND == 1 ? (BXDF_REFLECTION | BXDF_GLOSSY)
         : KIND == 4 ? (BXDF_TRANSMISSION | BXDF_GLOSSY) : (BXDF_REFLECTION | BXDF_SPECULAR);
}
template <int KIND> FTN_HD constexpr bool lobe_matches(int flags) { return KIND >= 0 && (flags & lobe_type<KIND>()) == lobe_type<KIND>(); }

// microfacet.rs:135-160
FTN_HD float tr_d(const Lobe& l, V3 wh) {
    const float t2 = tan2_theta(wh);
    if (is_inf(t2)) return 0.0f;
    const float cos4 = cos2_theta(wh) * cos2_theta(wh);
    const float cp = cos_phi(wh), sp = sin_phi(wh);
    const float e = ((cp * cp) / (l.ax * l.ax) + (sp * sp) / (l.ay * l.ay)) * t2;
    return 1.0f / (FTN_PI * l.ax * l.ay * cos4 * (1.0f + e) * (1.0f + e));
}
FTN_HD float tr_lambda(const Lobe& l, V3 w) {
    const float att = fabsf(tan_theta(w));
    if (is_inf(att)) return 0.0f;
    const float cp = cos_phi(w), sp = sin_phi(w);
    const float alpha = sqrtf((cp * cp) * l.ax * l.ax + (sp * sp) * l.ay * l.ay);
    const float a2t2 = (alpha * att) * (alpha * att);
    return (-1.0f + sqrtf(1.0f + a2t2)) / 2.0f;
}
FTN_HD float tr_pdf(const Lobe& l, V3 wh) { return tr_d(l, wh) * abs_cos_theta(wh); }   // microfacet.rs:28-31
// microfacet.rs:162-186 (full-NDF sampling)
FTN_HD V3 tr_sample_wh(const Lobe& l, V3 wo, float u0, float u1) {
    float cos_t, phi;
    if (l.ax == l.ay) {
        const float tan2 = (l.ax * l.ax) * u0 / (1.0f - u0);
        cos_t = 1.0f / sqrtf(1.0f + tan2);
        phi = 2.0f * FTN_PI * u1;
    } else {
        phi = atanf(l.ay / l.ax * tanf(2.0f * FTN_PI * u1 + 0.5f * FTN_PI));
        if (u1 > 0.5f) phi += FTN_PI;
        const float sp = sinf(phi), cp = cosf(phi);
        const float alpha2 = 1.0f / ((cp * cp) / (l.ax * l.ax) + (sp * sp) / (l.ay * l.ay));
        const float tan2 = alpha2 * u0 / (1.0f - u0);
        cos_t = 1.0f / sqrtf(1.0f + tan2);
    }
    const float sin_t = sqrtf(fmaxf(0.0f, 1.0f - cos_t * cos_t));
    float sphi, cphi; f_sincos(phi, &sphi, &cphi);
    const V3 wh = V3(sin_t * cphi, sin_t * sphi, cos_t);   // spherical_direction, math.rs:74-80
    return same_hemisphere(wo, wh) ? wh : -wh;
}
template <int FRESNEL> FTN_HD V3 lobe_fresnel(const Lobe& l, float cos_i) {
    if (FRESNEL == 0) return fresnel_conductor(fabsf(cos_i), l.eta, l.k);   // fresnel.rs:70-73
    if (FRESNEL == 2) return v3s(fresnel_dielectric(cos_i, 1.0f, l.eta.x));
    return v3s(fresnel_dielectric(cos_i, 1.5f, 1.0f));
}
// refract, reflection/mod.rs:70-78
FTN_HD bool refract(V3 wi, V3 n, float eta, V3* wt) {
    const float cos_i = dot(n, wi);
    const float sin2_i = fmaxf(0.0f, 1.0f - cos_i * cos_i);
    const float sin2_t = eta * eta * sin2_i;
    if (sin2_t >= 1.0f) return false;
    const float cos_t = sqrtf(1.0f - sin2_t);
    *wt = eta * -wi + (eta * cos_i - cos_t) * n;
    return true;
}
// MicrofacetTransmission::get_eta with eta_a = 1, eta_b = the material's index (reflection/mod.rs:376-378)
FTN_HD float mt_eta(const Lobe& l, V3 wo) { return (wo.z > 0.0f) ? l.eta.x / 1.0f : 1.0f / l.eta.x; }

template <int KIND, int FRESNEL> FTN_HD V3 lobe_f(const Lobe& l, V3 wo, V3 wi) {
    if (KIND == 0) return l.r * FTN_INV_PI;   // reflection/mod.rs:159-161
    if (KIND == 2) return v3s(0.0f);           // reflection/mod.rs:181-183
    if (KIND == 4) {                           // MicrofacetTransmission::f, reflection/mod.rs:386-404 (TransportMode::Radiance)
        if (same_hemisphere(wo, wi)) return v3s(0.0f);
        const float cos_o = wo.z, cos_i = wi.z;
        if (cos_o == 0.0f || cos_i == 0.0f) return v3s(0.0f);
        const float eta = mt_eta(l, wo);
        V3 wh = normalize(wo + wi * eta);
        if (wh.z < 0.0f) wh = -wh;
        const float F = fresnel_dielectric(dot(wo, wh), 1.0f, l.eta.x);
        const float sqrt_denom = dot(wo, wh) + eta * dot(wi, wh);
        const float factor = 1.0f / eta;
        const float G = 1.0f / (1.0f + tr_lambda(l, wo) + tr_lambda(l, wi));
        return (v3s(1.0f) - v3s(F)) * l.r *
               fabsf(tr_d(l, wh) * G * (eta * eta) * abs_dot(wi, wh) * abs_dot(wo, wh) * (factor * factor) / (cos_i * cos_o * (sqrt_denom * sqrt_denom)));
    }
    if (KIND == 3) {                           // OrenNayar::f, reflection/mod.rs:274-296
        const float sin_i = sin_theta(wi), sin_o = sin_theta(wo);
        float max_cos = 0.0f;
        if (sin_i > 1.0e-4f && sin_o > 1.0e-4f) {
            const float d_cos = cos_phi(wi) * cos_phi(wo) + sin_phi(wi) * sin_phi(wo);
            max_cos = fmaxf(0.0f, d_cos);
        }
        float sin_alpha, tan_beta;
        if (abs_cos_theta(wi) > abs_cos_theta(wo)) { sin_alpha = sin_o; tan_beta = sin_i / abs_cos_theta(wi); }
        else { sin_alpha = sin_i; tan_beta = sin_o / abs_cos_theta(wo); }
        return l.r * FTN_INV_PI * (l.ax + (l.ay * max_cos * sin_alpha * tan_beta));
    }
    const float cos_o = abs_cos_theta(wo), cos_i = abs_cos_theta(wi);   // :318-336
    V3 wh = wi + wo;
    if (cos_i == 0.0f || cos_o == 0.0f || (wh.x == 0.0f && wh.y == 0.0f && wh.z == 0.0f)) return v3s(0.0f);
    wh = normalize(wh);
    const V3 whf = (wh.z < 0.0f) ? -wh : wh;   // faceforward(wh, (0,0,1)): dot = wh.z
    const V3 F = lobe_fresnel<FRESNEL>(l, dot(wi, whf));
    const float G = 1.0f / (1.0f + tr_lambda(l, wo) + tr_lambda(l, wi));   // microfacet.rs:21-23
    return l.r * tr_d(l, wh) * G * F / (4.0f * cos_i * cos_o);
}
template <int KIND> FTN_HD float lobe_pdf(const Lobe& l, V3 wo, V3 wi) {
    if (KIND == 0 || KIND == 3) return same_hemisphere(wo, wi) ? abs_cos_theta(wi) * FTN_INV_PI : 0.0f;   // DefaultSampleF :140-146
    if (KIND == 2) return 0.0f;                // :194-196
    if (KIND == 4) {                           // MicrofacetTransmission::pdf, reflection/mod.rs:426-435
        if (same_hemisphere(wo, wi)) return 0.0f;
        const float eta = mt_eta(l, wo);
        const V3 wh = normalize(wo + wi * eta);
        const float sqrt_denom = dot(wo, wh) + eta * dot(wi, wh);
        const float dwh_dwi = fabsf(((eta * eta) * dot(wi, wh)) / (sqrt_denom * sqrt_denom));
        return tr_pdf(l, wh) * dwh_dwi;
    }
    if (!same_hemisphere(wo, wi)) return 0.0f;   // :354-360
    const V3 wh = normalize(wo + wi);
    return tr_pdf(l, wh) / (4.0f * dot(wo, wh));
}
struct ScatterSample { V3 f, wi; float pdf; int type; };
template <int KIND, int FRESNEL> FTN_HD bool lobe_sample_f(const Lobe& l, V3 wo, float u0, float u1, ScatterSample* s) {
    if (KIND == 0 || KIND == 3) {   // DefaultSampleF :131-138
        V3 wi = cosine_sample_hemisphere(u0, u1);
        if (wo.z < 0.0f) wi.z *= -1.0f;
        s->pdf = lobe_pdf<KIND>(l, wo, wi); s->f = lobe_f<KIND, FRESNEL>(l, wo, wi); s->wi = wi; s->type = lobe_type<KIND>();
        return true;
    }
    if (KIND == 2) {   // reflection/mod.rs:185-192; FresnelNoOp evaluates to 1
        const V3 wi = V3(-wo.x, -wo.y, wo.z);
        s->pdf = 1.0f; s->f = (v3s(1.0f) * l.r) / abs_cos_theta(wi); s->wi = wi; s->type = lobe_type<KIND>();
        return true;
    }
    if (KIND == 4) {   // MicrofacetTransmission::sample_f, reflection/mod.rs:406-424
        if (wo.z == 0.0f) return false;
        const V3 wh = tr_sample_wh(l, wo, u0, u1);
        if (dot(wo, wh) < 0.0f) return false;
        const float eta = mt_eta(l, -wo);   // "NOTE: this inverts the eta fraction"
        V3 wi;
        if (!refract(wo, wh, eta, &wi)) return false;
        s->f = lobe_f<KIND, FRESNEL>(l, wo, wi); s->wi = wi; s->pdf = lobe_pdf<KIND>(l, wo, wi); s->type = lobe_type<KIND>();
        return true;
    }
    const V3 wh = tr_sample_wh(l, wo, u0, u1);   // :338-352
    const V3 wi = reflect(wo, wh);
    if (!same_hemisphere(wo, wi)) return false;
    s->pdf = tr_pdf(l, wh) / (4.0f * dot(wo, wh));
    s->f = lobe_f<KIND, FRESNEL>(l, wo, wi); s->wi = wi; s->type = lobe_type<KIND>();
    return true;
}

// reflection/bsdf.rs:8-148.  `on[i]`: lobe slot i of this material class is present at run time
// (matte with Kd == 0 or plastic with Kd / Ks == 0 drop the lobe, matte.rs:41, plastic.rs:28-33).
struct Bsdf {
    V3 ns, ng, ss, ts;
    Lobe l0, l1;
    bool on0, on1;
};
FTN_HD void bsdf_init(Bsdf* b, V3 ns, V3 ng, V3 shading_dpdu) {   // :31-46
    b->ns = ns; b->ng = ng;
    b->ss = x_normalize(shading_dpdu);
    b->ts = x_normalize(x_cross(ns, b->ss));
    b->on0 = false; b->on1 = false;
}
template <int MAT> FTN_HD int bsdf_num_components(const Bsdf& b, int flags) {
    return (int)(b.on0 && lobe_matches<LobeKind<MAT, 0>::value>(flags)) + (int)(b.on1 && lobe_matches<LobeKind<MAT, 1>::value>(flags));
}
FTN_HD V3 bsdf_to_local(const Bsdf& b, V3 v) { return V3(dot(v, b.ss), dot(v, b.ts), dot(v, b.ns)); }
FTN_HD V3 bsdf_to_world(const Bsdf& b, V3 v) {
    return V3(b.ss.x * v.x + b.ts.x * v.y + b.ns.x * v.z, b.ss.y * v.x + b.ts.y * v.y + b.ns.y * v.z, b.ss.z * v.x + b.ts.z * v.y + b.ns.z * v.z);
}
// Sum of f over the matching lobes that pass the reflect/transmit gate (bsdf.rs:72-80): a REFLECTION lobe counts when wi
// and wo lie on the same side of the geometric normal, a TRANSMISSION lobe (rough glass only) when they do not.
template <int KIND> FTN_HD constexpr bool lobe_gate(bool refl) { return ((lobe_type<KIND < 0 ? 0 : KIND>() & BXDF_TRANSMISSION) != 0) ? !refl : refl; }
template <int MAT> FTN_HD V3 bsdf_sum_f(const Bsdf& b, V3 wo, V3 wi, bool refl, int flags) {
    constexpr int K0 = LobeKind<MAT, 0>::value, K1 = LobeKind<MAT, 1>::value, FR = LobeFresnel<MAT>::value;
    V3 sum = v3s(0.0f);
    if (K0 >= 0 && b.on0 && lobe_matches<K0>(flags) && lobe_gate<K0>(refl)) sum = sum + lobe_f<K0 < 0 ? 0 : K0, FR>(b.l0, wo, wi);
    if (K1 >= 0 && b.on1 && lobe_matches<K1>(flags) && lobe_gate<K1>(refl)) sum = sum + lobe_f<K1 < 0 ? 0 : K1, FR>(b.l1, wo, wi);
    return sum;
}
template <int MAT> FTN_HD V3 bsdf_f(const Bsdf& b, V3 wo_w, V3 wi_w, int flags) {   // :67-82
    const V3 wi = bsdf_to_local(b, wi_w), wo = bsdf_to_local(b, wo_w);
    if (wo.z == 0.0f) return v3s(0.0f);
    const bool refl = dot(wi_w, b.ng) * dot(wo_w, b.ng) > 0.0f;
    return bsdf_sum_f<MAT>(b, wo, wi, refl, flags);
}
template <int MAT> FTN_HD float bsdf_pdf(const Bsdf& b, V3 wo_w, V3 wi_w, int flags) {   // :131-144
    constexpr int K0 = LobeKind<MAT, 0>::value, K1 = LobeKind<MAT, 1>::value;
    const V3 wo = bsdf_to_local(b, wo_w), wi = bsdf_to_local(b, wi_w);
    if (wo.z == 0.0f) return 0.0f;
    float p = 0.0f; int nm = 0;
    if (K0 >= 0 && b.on0 && lobe_matches<K0>(flags)) { p += lobe_pdf<K0 < 0 ? 0 : K0>(b.l0, wo, wi); ++nm; }
    if (K1 >= 0 && b.on1 && lobe_matches<K1>(flags)) { p += lobe_pdf<K1 < 0 ? 0 : K1>(b.l1, wo, wi); ++nm; }
    return nm > 0 ? p / (float)nm : 0.0f;
}
template <int MAT> FTN_HD bool bsdf_sample_f(const Bsdf& b, V3 wo_w, float u0, float u1, int flags, ScatterSample* out) {   // :85-129
    constexpr int K0 = LobeKind<MAT, 0>::value, K1 = LobeKind<MAT, 1>::value, FR = LobeFresnel<MAT>::value;
    const bool m0 = K0 >= 0 && b.on0 && lobe_matches<K0>(flags), m1 = K1 >= 0 && b.on1 && lobe_matches<K1>(flags);
    const int nm = (int)m0 + (int)m1;
    if (nm == 0) return false;
    const float matching = (float)nm;
    const int comp = (int)fminf(floorf(u0 * matching), matching - 1.0f);
    // the comp-th matching lobe: slot 0 if it matches and comp == 0, otherwise slot 1
    const bool pick0 = m0 && comp == 0;
    const float ur0 = u0 * matching - (float)comp;
    const V3 wo = bsdf_to_local(b, wo_w);
    ScatterSample s;
    bool ok;
    if (pick0) ok = lobe_sample_f<K0 < 0 ? 0 : K0, FR>(b.l0, wo, ur0, u1, &s);
    else ok = lobe_sample_f<K1 < 0 ? 0 : K1, FR>(b.l1, wo, ur0, u1, &s);
    if (!ok) return false;
    if (s.pdf == 0.0f) return false;
    const V3 wi = s.wi;
    const V3 wi_w = bsdf_to_world(b, wi);
    float pdf = s.pdf;
    const bool specular = (s.type & BXDF_SPECULAR) != 0;   // bsdf.rs:106-127: a specular sample keeps its own pdf and f
    if (nm > 1) {
        if (!specular) {
            if (pick0) pdf += lobe_pdf<K1 < 0 ? 0 : K1>(b.l1, wo, wi);
            else pdf += lobe_pdf<K0 < 0 ? 0 : K0>(b.l0, wo, wi);
        }
        pdf /= matching;
    }
    if (specular) out->f = s.f;
    else {
        const bool refl = dot(wi_w, b.ng) * dot(wo_w, b.ng) > 0.0f;
        out->f = bsdf_sum_f<MAT>(b, wo, wi, refl, flags);
    }
    out->wi = wi_w; out->pdf = pdf; out->type = s.type;
    return true;
}

// ---- MIPMap lookups of an image texture (mipmap.rs:245-311) -----------------------------------------------------
// SurfaceInteraction::tex_diffs (interaction.rs:193-215); all zero when there is no differential
struct TexDiffs { float dudx, dvdx, dudy, dvdy; };

FTN_HD int max_i(int a, int b) { return a > b ? a : b; }
FTN_HD int min_i(int a, int b) { return a < b ? a : b; }
struct MipLevel { const F4* texels; int w, h; };
template <class TEX> FTN_HD MipLevel mip_level(const TEX& m, int level) {
    size_t off = 0;
    for (int l = 0; l < level; ++l) off += (size_t)max_i(1, m.img_w >> l) * (size_t)max_i(1, m.img_h >> l);
    MipLevel lv; lv.texels = m.image + off; lv.w = max_i(1, m.img_w >> level); lv.h = max_i(1, m.img_h >> level);
    return lv;
}
// get_texel_from_level, mipmap.rs:297-311
FTN_HD V3 mip_texel(const MipLevel& lv, int wrap, int s, int t) {
    if (wrap == FTN_WRAP_REPEAT) { s %= lv.w; if (s < 0) s += lv.w; t %= lv.h; if (t < 0) t += lv.h; }   // rem_euclid
    else if (wrap == FTN_WRAP_CLAMP) { s = min_i(max_i(s, 0), lv.w - 1); t = min_i(max_i(t, 0), lv.h - 1); }
    else if (s < 0 || s >= lv.w || t < 0 || t >= lv.h) return v3s(0.0f);
    const F4 v = ld4(lv.texels + (size_t)t * lv.w + s);
    return V3(v.x, v.y, v.z);
}
// triangle, mipmap.rs:265-279: the four texels around the continuous coordinate
template <class TEX> FTN_HD V3 mip_triangle(const TEX& m, int level, float st0, float st1) {
    const MipLevel lv = mip_level(m, min_i(max_i(level, 0), m.img_levels - 1));
    const float s = st0 * (float)lv.w - 0.5f, t = st1 * (float)lv.h - 0.5f;
    const float fs = floorf(s), ft = floorf(t);
    const int s0 = (int)fs, t0 = (int)ft;
    const float ds = s - fs, dt = t - ft;
    return mip_texel(lv, m.img_wrap, s0, t0) * ((1.0f - ds) * (1.0f - dt)) + mip_texel(lv, m.img_wrap, s0, t0 + 1) * ((1.0f - ds) * dt)
         + mip_texel(lv, m.img_wrap, s0 + 1, t0) * (ds * (1.0f - dt)) + mip_texel(lv, m.img_wrap, s0 + 1, t0 + 1) * (ds * dt);
}
// lookup_trilinear + lookup_trilinear_width, mipmap.rs:245-262 (`dst0.y` without abs(), as there)
template <class TEX> FTN_HD_COLD V3 mip_lookup_trilinear(const TEX& m, float st0, float st1, float dsdx, float dtdx, float dsdy, float dtdy) {
    const float width = 2.0f * fmaxf(fmaxf(fabsf(dsdx), dtdx), fmaxf(fabsf(dsdy), fabsf(dtdy)));
    const float level = (float)m.img_levels - 1.0f + log2f(fmaxf(width, 1.0e-8f));
    if (level < 0.0f) return mip_triangle(m, 0, st0, st1);
    if (level >= (float)(m.img_levels - 1)) return mip_texel(mip_level(m, m.img_levels - 1), m.img_wrap, 0, 0);
    const float lf = floorf(level), delta = level - lf;
    return mip_triangle(m, (int)lf, st0, st1) * (1.0f - delta) + mip_triangle(m, (int)lf + 1, st0, st1) * delta;
}

// Material::compute_scattering_functions: matte.rs:36-52, metal.rs:38-65, plastic.rs:24-48, mirror.rs:21-30
// Kd through its texture: constant, checkerboard (AAMethod::None, checkerboard.rs:50-64), uv (uv.rs:18-23) or image
// (image.rs:30-33); st = scale * uv + delta (mapping.rs:40-52), explicitly rounded so that floor() sees the oracle's values
FTN_HD V3 material_kd(const MaterialData& m, float u, float v, const TexDiffs& td) {
    if (m.kd_texture == 0) return V3(m.kd[0], m.kd[1], m.kd[2]);
    const float s = rn_add(rn_mul(m.uv_scale[0], u), m.uv_delta[0]), t = rn_add(rn_mul(m.uv_scale[1], v), m.uv_delta[1]);
    if (m.kd_texture == 1) return (((int)floorf(s) + (int)floorf(t)) % 2 == 0) ? V3(m.tex1[0], m.tex1[1], m.tex1[2]) : V3(m.tex2[0], m.tex2[1], m.tex2[2]);
    if (m.kd_texture == FTN_TEXTURE_IMAGE)   // dst_dx = (su dudx, sv dvdx), dst_dy = (su dudy, sv dvdy), mapping.rs:43-44
        return mip_lookup_trilinear(m, s, t, m.uv_scale[0] * td.dudx, m.uv_scale[1] * td.dvdx, m.uv_scale[0] * td.dudy, m.uv_scale[1] * td.dvdy);
    return V3(rn_sub(s, floorf(s)), rn_sub(t, floorf(t)), 0.0f);
}
// A texture-table entry: the same four textures as the inline slot (texture/{mod,checkerboard,uv,image}.rs)
template <bool IMG> FTN_HD V3 texture_eval(const TextureData& t, float u, float v, const TexDiffs& td) {
    if (t.type == FTN_TEXTURE_CONSTANT) return V3(t.v[0], t.v[1], t.v[2]);
    const float s = rn_add(rn_mul(t.uv_scale[0], u), t.uv_delta[0]), tt = rn_add(rn_mul(t.uv_scale[1], v), t.uv_delta[1]);
    if (t.type == FTN_TEXTURE_CHECKERBOARD) return (((int)floorf(s) + (int)floorf(tt)) % 2 == 0) ? V3(t.tex1[0], t.tex1[1], t.tex1[2]) : V3(t.tex2[0], t.tex2[1], t.tex2[2]);
    if (IMG && t.type == FTN_TEXTURE_IMAGE)    // the mip lookup is compiled only into the shaders of scenes that hold an image texture
        return mip_lookup_trilinear(t, s, tt, t.uv_scale[0] * td.dudx, t.uv_scale[1] * td.dvdx, t.uv_scale[0] * td.dudy, t.uv_scale[1] * td.dvdy);
    return V3(rn_sub(s, floorf(s)), rn_sub(tt, floorf(tt)), 0.0f);
}
// a material parameter: its texture-table entry when it names one, its constant otherwise (loaders/constructors.rs:192-238)
template <bool IMG> FTN_HD V3 mat_spectrum(const SceneView& sc, const MaterialData& m, int P, V3 constant, float u, float v, const TexDiffs& td) {
    const uint32_t id = m.ptex[P];
    return id ? texture_eval<IMG>(sc.textures[id - 1u], u, v, td) : constant;
}
template <bool IMG> FTN_HD float mat_float(const SceneView& sc, const MaterialData& m, int P, float constant, float u, float v, const TexDiffs& td) {
    const uint32_t id = m.ptex[P];
    return id ? texture_eval<IMG>(sc.textures[id - 1u], u, v, td).x : constant;
}
// Kd (matte, plastic) / Kr (mirror): table entry, else the inline slot, else the constant
template <bool IMG> FTN_HD V3 mat_kd(const SceneView& sc, const MaterialData& m, int P, float u, float v, const TexDiffs& td) {
    const uint32_t id = m.ptex[P];
    return id ? texture_eval<IMG>(sc.textures[id - 1u], u, v, td) : material_kd(m, u, v, td);
}
// TrowbridgeReitzDistribution::roughness_to_alpha, microfacet.rs:40-45 (per hit only when the roughness is textured)
FTN_HD float roughness_to_alpha_dev(float roughness) {
    const float x = logf(fmaxf(roughness, 1.0e-3f));
    return 1.62142f + 0.819955f * x + 0.1734f * x * x + 0.0171201f * x * x * x + 0.000640711f * x * x * x * x;
}
// the lobe's alphas: the values remapped at scene creation, or the textured roughnesses evaluated (and remapped) at this hit
template <bool IMG> FTN_HD void mat_alphas(const SceneView& sc, const MaterialData& m, float u, float v, const TexDiffs& td, float* ax, float* ay) {
    *ax = m.alpha_x; *ay = m.alpha_y;
    if (m.ptex[FTN_PARAM_UROUGHNESS] | m.ptex[FTN_PARAM_VROUGHNESS]) {
        float ur = mat_float<IMG>(sc, m, FTN_PARAM_UROUGHNESS, m.u_rough, u, v, td), vr = mat_float<IMG>(sc, m, FTN_PARAM_VROUGHNESS, m.v_rough, u, v, td);
        if (m.remap) { ur = roughness_to_alpha_dev(ur); vr = roughness_to_alpha_dev(vr); }
        *ax = ur; *ay = vr;
    }
}
FTN_HD V3 clamp_positive(V3 c) { return V3(clampf(c.x, 0.0f, FTN_INF), clampf(c.y, 0.0f, FTN_INF), clampf(c.z, 0.0f, FTN_INF)); }

// `unsupported` is set where the reference would hit todo!() (a glass whose textured alphas are both 0 at this hit, glass.rs:64-67)
template <int MAT, bool IMG> FTN_HD void material_bsdf(const SceneView& sc, const MaterialData& m, float u, float v, const TexDiffs& td, Bsdf* b, bool* unsupported) {
    if (MAT == FTN_MATERIAL_MATTE) {
        const V3 r = clamp_positive(mat_kd<IMG>(sc, m, FTN_PARAM_KD, u, v, td));
        if (!is_black(r)) { b->on0 = true; b->l0.r = r; }
    } else if (MAT == FTN_MATERIAL_METAL) {
        b->on0 = true;
        b->l0.r = v3s(1.0f);
        mat_alphas<IMG>(sc, m, u, v, td, &b->l0.ax, &b->l0.ay);
        b->l0.eta = mat_spectrum<IMG>(sc, m, FTN_PARAM_ETA, V3(m.eta[0], m.eta[1], m.eta[2]), u, v, td);
        b->l0.k = mat_spectrum<IMG>(sc, m, FTN_PARAM_K, V3(m.k[0], m.k[1], m.k[2]), u, v, td);
    } else if (MAT == FTN_CLASS_OREN_NAYAR) {   // matte.rs:45-49: OrenNayar::new(r, Deg(sigma)); (a, b) precomputed at scene creation unless sigma is textured
        const V3 r = clamp_positive(mat_kd<IMG>(sc, m, FTN_PARAM_KD, u, v, td));
        if (!is_black(r)) {
            b->on0 = true; b->l0.r = r; b->l0.ax = m.alpha_x; b->l0.ay = m.alpha_y;
            if (m.ptex[FTN_PARAM_SIGMA]) {      // reflection/mod.rs:259-267; sigma == 0 gives a = 1, b = 0: Lambert's r / pi exactly
                const float sigma = clampf(mat_float<IMG>(sc, m, FTN_PARAM_SIGMA, m.sigma, u, v, td), 0.0f, 90.0f);
                const float sr = sigma * (float)(3.14159265358979323846 / 180.0), s2 = sr * sr;
                b->l0.ax = (sigma == 0.0f) ? 1.0f : 1.0f - (s2 / (2.0f * (s2 + 0.33f)));
                b->l0.ay = (sigma == 0.0f) ? 0.0f : 0.45f * s2 / (s2 + 0.09f);
            }
        }
    } else if (MAT == FTN_MATERIAL_GLASS) {   // glass.rs:52-96, the non-specular branch (rough glass)
        const V3 r = clamp_positive(mat_spectrum<IMG>(sc, m, FTN_PARAM_KR, V3(m.kd[0], m.kd[1], m.kd[2]), u, v, td));
        const V3 t = clamp_positive(mat_spectrum<IMG>(sc, m, FTN_PARAM_KT, V3(m.ks[0], m.ks[1], m.ks[2]), u, v, td));
        const float eta = mat_float<IMG>(sc, m, FTN_PARAM_INDEX, m.eta[0], u, v, td);
        float ax, ay; mat_alphas<IMG>(sc, m, u, v, td, &ax, &ay);
        if (ax == 0.0f && ay == 0.0f) { *unsupported = true; return; }
        if (!is_black(r)) { b->on0 = true; b->l0.r = r; b->l0.ax = ax; b->l0.ay = ay; b->l0.eta = v3s(eta); b->l0.k = v3s(0.0f); }
        if (!is_black(t)) { b->on1 = true; b->l1.r = t; b->l1.ax = ax; b->l1.ay = ay; b->l1.eta = v3s(eta); b->l1.k = v3s(0.0f); }
    } else if (MAT == FTN_MATERIAL_MIRROR) {   // mirror.rs:21-30; Kr (constant or textured) is carried in MaterialData::kd and its texture fields
        const V3 r = clamp_positive(mat_kd<IMG>(sc, m, FTN_PARAM_KR, u, v, td));
        if (!is_black(r)) { b->on0 = true; b->l0.r = r; }
    } else {                                    // plastic.rs:24-48
        const V3 kd = mat_kd<IMG>(sc, m, FTN_PARAM_KD, u, v, td), ks = mat_spectrum<IMG>(sc, m, FTN_PARAM_KS, V3(m.ks[0], m.ks[1], m.ks[2]), u, v, td);
        if (!is_black(kd)) { b->on0 = true; b->l0.r = kd; }
        if (!is_black(ks)) {
            b->on1 = true; b->l1.r = ks; b->l1.eta = v3s(0.0f); b->l1.k = v3s(0.0f);
            float ax, ay; mat_alphas<IMG>(sc, m, u, v, td, &ax, &ay);
            b->l1.ax = ax; b->l1.ay = ax;
        }
    }
}

// ---- surface reconstruction -------------------------------------------------------------------------------------
// SurfaceHit (interaction.rs:12-18) + what Bsdf::new needs (shading normal, shading dpdu).
struct Surface {
    V3 p, p_err, n;      // hit.p, hit.p_err, hit.n
    V3 ns, sdpdu;        // shading_n, shading_geom.dpdu
    V3 wo;               // SurfaceInteraction.wo
    float u, v;          // SurfaceInteraction.uv (texture lookups)
    int material;        // -1 none
    int light;           // area light index, -1 none
};

// Triangle::intersect tail, triangle.rs:270-392, from the slot's vertices and the barycentrics
// the traversal found (exact ops: p, p_err and n feed spawned ray origins).
FTN_HD void triangle_surface_at(const SceneView& sc, V3 p0, V3 p1, V3 p2, uint32_t prim, const MeshData& mesh, const TriHit& h, V3 ray_d, Surface* s) {
    const uint32_t v0 = sc.idx[3 * (size_t)prim], v1 = sc.idx[3 * (size_t)prim + 1], v2 = sc.idx[3 * (size_t)prim + 2];
    float uv[3][2] = {{0.0f, 0.0f}, {1.0f, 0.0f}, {1.0f, 1.0f}};   // triangle.rs:131-143
    if (sc.uv) {
        uv[0][0] = sc.uv[2 * v0]; uv[0][1] = sc.uv[2 * v0 + 1]; uv[1][0] = sc.uv[2 * v1]; uv[1][1] = sc.uv[2 * v1 + 1];
        uv[2][0] = sc.uv[2 * v2]; uv[2][1] = sc.uv[2 * v2 + 1];
    }
    // uv of the hit, triangle.rs:299-300: b0 uv0 + b1 uv1 + b2 uv2
    s->u = rn_add(rn_add(rn_mul(h.b0, uv[0][0]), rn_mul(h.b1, uv[1][0])), rn_mul(h.b2, uv[2][0]));
    s->v = rn_add(rn_add(rn_mul(h.b0, uv[0][1]), rn_mul(h.b1, uv[1][1])), rn_mul(h.b2, uv[2][1]));
    const float duv02x = rn_sub(uv[0][0], uv[2][0]), duv02y = rn_sub(uv[0][1], uv[2][1]);
    const float duv12x = rn_sub(uv[1][0], uv[2][0]), duv12y = rn_sub(uv[1][1], uv[2][1]);
    const V3 dp02 = x_sub(p0, p2), dp12 = x_sub(p1, p2);
    const float determinant = rn_sub(rn_mul(duv02x, duv12y), rn_mul(duv02y, duv12x));
    V3 dpdu, dpdv;
    if (fabsf(determinant) < 1.0e-8f) {
        // degenerate uv; a zero-area triangle cannot reach here (det == 0 rejects it in the hit test)
        const V3 ng = x_cross(x_sub(p2, p0), x_sub(p1, p0));
        coordinate_system(x_normalize(ng), &dpdu, &dpdv);
    } else {
        const float inv = rn_div(1.0f, determinant);
        dpdu = x_scale(x_sub(x_scale(dp02, duv12y), x_scale(dp12, duv02y)), inv);
    }
    const float xs = rn_add(rn_add(fabsf(rn_mul(h.b0, p0.x)), fabsf(rn_mul(h.b1, p1.x))), fabsf(rn_mul(h.b2, p2.x)));
    const float ys = rn_add(rn_add(fabsf(rn_mul(h.b0, p0.y)), fabsf(rn_mul(h.b1, p1.y))), fabsf(rn_mul(h.b2, p2.y)));
    const float zs = rn_add(rn_add(fabsf(rn_mul(h.b0, p0.z)), fabsf(rn_mul(h.b1, p1.z))), fabsf(rn_mul(h.b2, p2.z)));
    const float g7 = gamma_n(7);
    s->p_err = V3(rn_mul(g7, xs), rn_mul(g7, ys), rn_mul(g7, zs));
    s->p = x_add(x_add(x_scale(p0, h.b0), x_scale(p1, h.b1)), x_scale(p2, h.b2));
    V3 n = x_normalize(x_cross(dp02, dp12));
    V3 ns = n;
    if (mesh.flags & FTN_MESH_FLIP_NORMALS) { n = x_scale(n, -1.0f); ns = x_scale(ns, -1.0f); }
    V3 sdpdu = dpdu;
    if (sc.nrm) {
        const V3 n0 = V3(sc.nrm[3 * v0], sc.nrm[3 * v0 + 1], sc.nrm[3 * v0 + 2]);
        const V3 n1 = V3(sc.nrm[3 * v1], sc.nrm[3 * v1 + 1], sc.nrm[3 * v1 + 2]);
        const V3 n2 = V3(sc.nrm[3 * v2], sc.nrm[3 * v2 + 1], sc.nrm[3 * v2 + 2]);
        ns = x_normalize(x_add(x_add(x_scale(n0, h.b0), x_scale(n1, h.b1)), x_scale(n2, h.b2)));
        V3 ss = x_normalize(dpdu);
        V3 ts = x_cross(ns, ss);
        if (x_len2(ts) > 0.0f) { ts = x_normalize(ts); ss = x_cross(ts, ns); }
        else { coordinate_system(ns, &ts, &ss); }
        sdpdu = ss;
        n = faceforward(n, ns);
    }
    s->n = n; s->ns = ns; s->sdpdu = sdpdu;
    s->wo = x_neg(ray_d);
    s->material = mesh.material;
    s->light = mesh.light_base >= 0 ? mesh.light_base + (int)(prim - mesh.first_tri) : -1;   // its own DiffuseAreaLight (primitive.rs:60-66)
}
FTN_HD void triangle_surface(const SceneView& sc, uint32_t slot, const TriHit& h, V3 ray_d, Surface* s) {
    F4 a, b, c; load_tri(sc.bvh, slot, &a, &b, &c);
    triangle_surface_at(sc, V3(a.x, a.y, a.z), V3(b.x, b.y, b.z), V3(c.x, c.y, c.z), f2u(a.w), sc.meshes[f2u(b.w)], h, ray_d, s);
}

FTN_HD void sphere_surface(const SphereData& sd, const SphereHit& h, Surface* s) {
    s->p = h.p; s->p_err = h.p_err; s->n = h.n; s->ns = h.ns; s->sdpdu = h.dpdu; s->wo = h.wo; s->u = h.u; s->v = h.v;
    s->material = sd.material; s->light = sd.light;
}

// SurfaceHit::spawn_ray, interaction.rs:22-30
FTN_HD V3 spawn_origin(const Surface& s, V3 dir) { return offset_ray_origin(s.p, s.p_err, s.n, dir); }

// ---- sphere area light (light/diffuse.rs, shapes/mod.rs:43-66, sphere.rs:202-218) ---------------------
struct ShapeSample { V3 p, p_err, n; };
FTN_HD ShapeSample sphere_sample(const SphereData& sd, float u0, float u1) {
    const float z = rn_sub(1.0f, rn_mul(2.0f, u0));   // uniform_sample_sphere, sampling.rs:37-42
    const float r = rn_sqrt(fmaxf(rn_sub(1.0f, rn_mul(z, z)), 0.0f));
    const float phi = rn_mul(rn_mul(2.0f, FTN_PI), u1);
    V3 p_obj = x_scale(V3(rn_mul(r, cosf(phi)), rn_mul(r, sinf(phi)), z), sd.radius);
    V3 n = x_normalize(transform_normal_inv(sd.w2o, p_obj));
    if (sd.reverse_orientation) n = x_scale(n, -1.0f);
    p_obj = x_scale(p_obj, rn_div(sd.radius, x_len(p_obj)));
    const V3 pe = x_scale(x_abs(p_obj), gamma_n(5));
    ShapeSample s;
    s.p = point_tf_err_to_err(sd.o2w, p_obj, pe, &s.p_err);
    s.n = n;
    return s;
}
// Shape::pdf_from_ref, shapes/mod.rs:55-66
FTN_HD float sphere_pdf_from_ref(const SphereData& sd, const Surface& ref, V3 wi) {
    RayF ray; ray.o = spawn_origin(ref, wi); ray.d = wi; ray.t_max = FTN_INF; ray.time = 0.0f;
    SphereHit h;
    if (!sphere_intersect(sd, ray, &h)) return 0.0f;
    const V3 d = x_sub(ref.p, h.p);
    return rn_div(x_len2(d), rn_mul(x_abs_dot(h.n, x_neg(wi)), sd.area));
}

// ---- triangle area light (light/diffuse.rs over shapes/triangle.rs:171-174, 395-420 and the Shape defaults shapes/mod.rs:41-66) ----
struct TriLightGeom { V3 p0, p1, p2; uint32_t prim, v0, v1, v2; MeshData mesh; };
FTN_HD TriLightGeom tri_light_geom(const SceneView& sc, const LightData& light) {
    const uint32_t prim = (uint32_t)light.sphere;
    TriLightGeom g; g.prim = prim;
    g.v0 = sc.idx[3 * (size_t)prim]; g.v1 = sc.idx[3 * (size_t)prim + 1]; g.v2 = sc.idx[3 * (size_t)prim + 2];
    g.p0 = V3(sc.pos[3 * (size_t)g.v0], sc.pos[3 * (size_t)g.v0 + 1], sc.pos[3 * (size_t)g.v0 + 2]);
    g.p1 = V3(sc.pos[3 * (size_t)g.v1], sc.pos[3 * (size_t)g.v1 + 1], sc.pos[3 * (size_t)g.v1 + 2]);
    g.p2 = V3(sc.pos[3 * (size_t)g.v2], sc.pos[3 * (size_t)g.v2 + 1], sc.pos[3 * (size_t)g.v2 + 2]);
    g.mesh = sc.meshes[light.mesh];
    return g;
}
// Triangle::sample, triangle.rs:395-420 (uniform_sample_triangle, sampling.rs:48-51)
FTN_HD ShapeSample triangle_sample(const SceneView& sc, const TriLightGeom& g, float u0, float u1) {
    const float su0 = rn_sqrt(u0);
    const float b0 = rn_sub(1.0f, su0), b1 = rn_mul(u1, su0), b2 = rn_sub(rn_sub(1.0f, b0), b1);
    const V3 q0 = x_scale(g.p0, b0), q1 = x_scale(g.p1, b1), q2 = x_scale(g.p2, b2);
    V3 n = x_normalize(x_cross(x_sub(g.p1, g.p0), x_sub(g.p2, g.p0)));
    if (sc.nrm) {
        const V3 n0 = V3(sc.nrm[3 * (size_t)g.v0], sc.nrm[3 * (size_t)g.v0 + 1], sc.nrm[3 * (size_t)g.v0 + 2]);
        const V3 n1 = V3(sc.nrm[3 * (size_t)g.v1], sc.nrm[3 * (size_t)g.v1 + 1], sc.nrm[3 * (size_t)g.v1 + 2]);
        const V3 n2 = V3(sc.nrm[3 * (size_t)g.v2], sc.nrm[3 * (size_t)g.v2 + 1], sc.nrm[3 * (size_t)g.v2 + 2]);
        const V3 ns = x_normalize(x_add(x_add(x_scale(n0, b0), x_scale(n1, b1)), x_scale(n2, b2)));
        n = faceforward(n, ns);
    } else if (g.mesh.flags & FTN_MESH_FLIP_NORMALS) n = x_scale(n, -1.0f);
    ShapeSample s;
    s.p = x_add(x_add(q0, q1), q2);                                              // Point3(0,0,0) + sample_p
    s.p_err = x_scale(x_add(x_add(x_abs(q0), x_abs(q1)), x_abs(q2)), gamma_n(6));
    s.n = n;
    return s;
}
// Shape::pdf_from_ref (shapes/mod.rs:55-66) over Triangle::intersect and Triangle::area
FTN_HD float triangle_pdf_from_ref(const SceneView& sc, const TriLightGeom& g, const Surface& ref, V3 wi) {
    const V3 o = spawn_origin(ref, wi);
    TriHit th;
    if (!triangle_intersect(g.p0, g.p1, g.p2, o, make_ray_shear(wi), FTN_INF, &th)) return 0.0f;
    Surface hs;
    triangle_surface_at(sc, g.p0, g.p1, g.p2, g.prim, g.mesh, th, wi, &hs);
    const float area = rn_mul(0.5f, x_len(x_cross(x_sub(g.p1, g.p0), x_sub(g.p2, g.p0))));
    const V3 d = x_sub(ref.p, hs.p);
    return rn_div(x_len2(d), rn_mul(x_abs_dot(hs.n, x_neg(wi)), area));
}

// ---- thin-lens perspective camera, camera/mod.rs:117-143 (exact ops: primary rays match the oracle) -----
FTN_HD RayF camera_ray(const FtnCamera& cam, float fx, float fy, float lx, float ly, float tu) {
    M4 r2c, c2w;
    for (int i = 0; i < 16; ++i) { r2c.m[i] = cam.raster_to_camera[i]; c2w.m[i] = cam.camera_to_world[i]; }
    const V3 pc = transform_point(r2c, V3(fx, fy, 0.0f));
    RayF ray;
    ray.o = V3(0.0f, 0.0f, 0.0f);
    ray.d = x_normalize(pc);
    ray.time = rn_add(rn_mul(rn_sub(1.0f, tu), cam.shutter_open), rn_mul(tu, cam.shutter_close));   // lerp, math.rs:21-23
    ray.t_max = FTN_INF;
    if (cam.lens_radius > 0.0f) {
        float dx, dy;
        {   // concentric_sample_disk with exact ops
            const float ox = rn_sub(rn_mul(2.0f, lx), 1.0f), oy = rn_sub(rn_mul(2.0f, ly), 1.0f);
            if (ox == 0.0f && oy == 0.0f) { dx = 0.0f; dy = 0.0f; }
            else {
                float theta, r;
                if (fabsf(ox) > fabsf(oy)) { theta = rn_mul(FTN_PI_4, rn_div(oy, ox)); r = ox; }
                else { theta = rn_sub(FTN_PI_2, rn_mul(FTN_PI_4, rn_div(ox, oy))); r = oy; }
                dx = rn_mul(r, cosf(theta)); dy = rn_mul(r, sinf(theta));
            }
        }
        const float plx = rn_mul(cam.lens_radius, dx), ply = rn_mul(cam.lens_radius, dy);
        const float ft = rn_div(cam.focal_distance, ray.d.z);
        const V3 pf = x_add(ray.o, x_scale(ray.d, ft));
        ray.o = V3(plx, ply, 0.0f);
        ray.d = x_normalize(x_sub(pf, ray.o));
    }
    V3 oe, de;
    return ray_transform_err(c2w, ray, &oe, &de);
}

// The two offset rays of generate_ray_differential (camera/mod.rs:145-205), transformed like points / vectors
// (transform.rs:324-338) and pulled towards the main ray by `scale` (geometry/mod.rs:125-132; the integrator passes
// 1 / sqrt(spp), integrator/mod.rs:249-251).  With a lens BOTH offsets use dx_camera (:178 repeats :172), as there.
struct RayDiff { V3 rx_o, rx_d, ry_o, ry_d; };
FTN_HD RayDiff camera_differential(const FtnCamera& cam, const RayF& main_ray, float fx, float fy, float lx, float ly, float scale) {
    M4 r2c, c2w;
    for (int i = 0; i < 16; ++i) { r2c.m[i] = cam.raster_to_camera[i]; c2w.m[i] = cam.camera_to_world[i]; }
    const V3 pc = transform_point(r2c, V3(fx, fy, 0.0f)), o0 = transform_point(r2c, V3(0.0f, 0.0f, 0.0f));
    const V3 dxc = transform_point(r2c, V3(1.0f, 0.0f, 0.0f)) - o0, dyc = transform_point(r2c, V3(0.0f, 1.0f, 0.0f)) - o0;
    RayDiff df;
    if (cam.lens_radius > 0.0f) {
        float dx = 0.0f, dy = 0.0f;
        const float ox = 2.0f * lx - 1.0f, oy = 2.0f * ly - 1.0f;   // concentric_sample_disk
        if (!(ox == 0.0f && oy == 0.0f)) {
            float theta, r;
            if (fabsf(ox) > fabsf(oy)) { theta = FTN_PI_4 * (oy / ox); r = ox; }
            else { theta = FTN_PI_2 - FTN_PI_4 * (ox / oy); r = oy; }
            dx = r * cosf(theta); dy = r * sinf(theta);
        }
        const V3 pl = V3(cam.lens_radius * dx, cam.lens_radius * dy, 0.0f);
        const V3 ddx = normalize(pc + dxc);
        const V3 pfx = ddx * (cam.focal_distance / ddx.z);
        df.rx_o = pl; df.rx_d = normalize(pfx - pl);
        df.ry_o = pl; df.ry_d = df.rx_d;
    } else {
        df.rx_o = v3s(0.0f); df.ry_o = v3s(0.0f);
        df.rx_d = normalize(pc + dxc); df.ry_d = normalize(pc + dyc);
    }
    df.rx_o = transform_point(c2w, df.rx_o); df.ry_o = transform_point(c2w, df.ry_o);
    df.rx_d = transform_vector(c2w, df.rx_d); df.ry_d = transform_vector(c2w, df.ry_d);
    df.rx_o = main_ray.o + (df.rx_o - main_ray.o) * scale; df.ry_o = main_ray.o + (df.ry_o - main_ray.o) * scale;
    df.rx_d = main_ray.d + (df.rx_d - main_ray.d) * scale; df.ry_d = main_ray.d + (df.ry_d - main_ray.d) * scale;
    return df;
}

// math.rs:56-72 solve_linear_system_2x2, A = from_cols((a00, a01), (a10, a11))
FTN_HD bool solve_2x2(float a00, float a01, float a10, float a11, float b0, float b1, float* x0, float* x1) {
    const float det = a00 * a11 - a10 * a01;
    if (fabsf(det) < 1.0e-10f) return false;
    *x0 = (a11 * b0 - a10 * b1) / det;
    *x1 = (a00 * b1 - a01 * b0) / det;
    return !(*x0 != *x0 || *x1 != *x1);
}
FTN_HD float v3_at(V3 v, int i) { return i == 0 ? v.x : (i == 1 ? v.y : v.z); }

// SurfaceInteraction::compute_tex_differentials, interaction.rs:124-176 (None -> zeros, :117).  dpdx / dpdy (optional) are
// TextureDifferentials::dpdx / dpdy, which specular_reflect needs (integrator/mod.rs:61-62); zero like the rest when a solve fails.
FTN_HD TexDiffs tex_differentials(V3 p, V3 n, V3 dpdu, V3 dpdv, const RayDiff& df, V3* dpdx_out = nullptr, V3* dpdy_out = nullptr) {
    TexDiffs td; td.dudx = td.dvdx = td.dudy = td.dvdy = 0.0f;
    if (dpdx_out) { *dpdx_out = v3s(0.0f); *dpdy_out = v3s(0.0f); }
    const float d = dot(n, p);
    const float tx = -(dot(n, df.rx_o) - d) / dot(n, df.rx_d), ty = -(dot(n, df.ry_o) - d) / dot(n, df.ry_d);
    const V3 dpdx = (df.rx_o + df.rx_d * tx) - p, dpdy = (df.ry_o + df.ry_d * ty) - p;
    int d0, d1;
    if (fabsf(n.x) > fabsf(n.y) && fabsf(n.x) > fabsf(n.z)) { d0 = 1; d1 = 2; }
    else if (fabsf(n.y) > fabsf(n.z)) { d0 = 0; d1 = 2; }
    else { d0 = 0; d1 = 1; }
    float dudx, dvdx, dudy, dvdy;
    if (!solve_2x2(v3_at(dpdu, d0), v3_at(dpdu, d1), v3_at(dpdv, d0), v3_at(dpdv, d1), v3_at(dpdx, d0), v3_at(dpdx, d1), &dudx, &dvdx)) return td;
    if (!solve_2x2(v3_at(dpdu, d0), v3_at(dpdu, d1), v3_at(dpdv, d0), v3_at(dpdv, d1), v3_at(dpdy, d0), v3_at(dpdy, d1), &dudy, &dvdy)) return td;
    td.dudx = dudx; td.dvdx = dvdx; td.dudy = dudy; td.dvdy = dvdy;
    if (dpdx_out) { *dpdx_out = dpdx; *dpdy_out = dpdy; }
    return td;
}

// DiffGeom::dpdu / dpdv of a triangle (triangle.rs:258-295), for the texture differentials only
FTN_HD void triangle_dpduv(const SceneView& sc, uint32_t slot, V3* dpdu, V3* dpdv) {
    F4 a, b, c; load_tri(sc.bvh, slot, &a, &b, &c);
    const V3 p0 = V3(a.x, a.y, a.z), p1 = V3(b.x, b.y, b.z), p2 = V3(c.x, c.y, c.z);
    const uint32_t prim = f2u(a.w);
    const uint32_t v0 = sc.idx[3 * (size_t)prim], v1 = sc.idx[3 * (size_t)prim + 1], v2 = sc.idx[3 * (size_t)prim + 2];
    float uv[3][2] = {{0.0f, 0.0f}, {1.0f, 0.0f}, {1.0f, 1.0f}};
    if (sc.uv) {
        uv[0][0] = sc.uv[2 * v0]; uv[0][1] = sc.uv[2 * v0 + 1]; uv[1][0] = sc.uv[2 * v1]; uv[1][1] = sc.uv[2 * v1 + 1];
        uv[2][0] = sc.uv[2 * v2]; uv[2][1] = sc.uv[2 * v2 + 1];
    }
    const float duv02x = uv[0][0] - uv[2][0], duv02y = uv[0][1] - uv[2][1], duv12x = uv[1][0] - uv[2][0], duv12y = uv[1][1] - uv[2][1];
    const V3 dp02 = p0 - p2, dp12 = p1 - p2;
    const float determinant = rn_sub(rn_mul(duv02x, duv12y), rn_mul(duv02y, duv12x));
    if (fabsf(determinant) < 1.0e-8f) { coordinate_system(x_normalize(x_cross(x_sub(p2, p0), x_sub(p1, p0))), dpdu, dpdv); return; }
    const float inv = 1.0f / determinant;
    *dpdu = (dp02 * duv12y - dp12 * duv02y) * inv;
    *dpdv = (dp12 * duv02x - dp02 * duv12x) * inv;
}

// shading_geom.dndu / dndv of a triangle (triangle.rs:357-375): zero without vertex normals (:308-313); only specular_reflect's
// ray differentials read them (integrator/mod.rs:65-66)
FTN_HD void triangle_dnduv(const SceneView& sc, uint32_t slot, V3* dndu, V3* dndv) {
    *dndu = v3s(0.0f); *dndv = v3s(0.0f);
    if (!sc.nrm) return;
    const uint32_t prim = f2u(ld4(sc.bvh.tris + (size_t)FTN_TRI_F4 * (size_t)slot).w);
    const uint32_t v0 = sc.idx[3 * (size_t)prim], v1 = sc.idx[3 * (size_t)prim + 1], v2 = sc.idx[3 * (size_t)prim + 2];
    float uv[3][2] = {{0.0f, 0.0f}, {1.0f, 0.0f}, {1.0f, 1.0f}};
    if (sc.uv) {
        uv[0][0] = sc.uv[2 * v0]; uv[0][1] = sc.uv[2 * v0 + 1]; uv[1][0] = sc.uv[2 * v1]; uv[1][1] = sc.uv[2 * v1 + 1];
        uv[2][0] = sc.uv[2 * v2]; uv[2][1] = sc.uv[2 * v2 + 1];
    }
    const V3 n0 = V3(sc.nrm[3 * v0], sc.nrm[3 * v0 + 1], sc.nrm[3 * v0 + 2]);
    const V3 n1 = V3(sc.nrm[3 * v1], sc.nrm[3 * v1 + 1], sc.nrm[3 * v1 + 2]);
    const V3 n2 = V3(sc.nrm[3 * v2], sc.nrm[3 * v2 + 1], sc.nrm[3 * v2 + 2]);
    const float duv02x = uv[0][0] - uv[2][0], duv02y = uv[0][1] - uv[2][1], duv12x = uv[1][0] - uv[2][0], duv12y = uv[1][1] - uv[2][1];
    const float determinant = rn_sub(rn_mul(duv02x, duv12y), rn_mul(duv02y, duv12x));
    if (fabsf(determinant) < 1.0e-8f) {
        const V3 dn = cross(n2 - n0, n1 - n0);
        if (dot(dn, dn) != 0.0f) coordinate_system(dn, dndu, dndv);   // `coordinate_system(dn)`: dn is NOT normalised there (:364)
        return;
    }
    const float inv = 1.0f / determinant;
    const V3 dn1 = n0 - n2, dn2 = n1 - n2;
    *dndu = (dn1 * duv12y - dn2 * duv02y) * inv;
    *dndv = (dn2 * duv02x - dn1 * duv12x) * inv;
}
// dndu / dndv of a sphere hit: the Weingarten equations of sphere.rs:160-178 on the object-space hit point, carried to world
// space like normals (DiffGeom::transform, transform.rs:353-354).  The object-space point is recovered from the world-space
// hit (tolerance-tested code, like everything that only feeds texture filtering).
FTN_HD void sphere_dnduv(const SphereData& s, V3 p_world, V3* dndu, V3* dndv) {
    V3 p = transform_point(s.w2o, p_world);
    p = p * (s.radius / sqrtf(dot(p, p)));
    if (p.x == 0.0f && p.y == 0.0f) p.x = 1.0e-5f * s.radius;
    const float theta = acosf(clampf(p.z / s.radius, -1.0f, 1.0f));
    const float inv_zr = 1.0f / sqrtf(p.x * p.x + p.y * p.y);
    const float cos_phi = p.x * inv_zr, sin_phi = p.y * inv_zr;
    const float dth = s.theta_max - s.theta_min;
    const V3 dpdu = V3(-s.phi_max * p.y, s.phi_max * p.x, 0.0f);
    const V3 dpdv = V3(p.z * cos_phi, p.z * sin_phi, -s.radius * sinf(theta)) * dth;
    const V3 d2pduu = V3(p.x, p.y, 0.0f) * (-s.phi_max * s.phi_max);
    const V3 d2pduv = V3(-sin_phi, cos_phi, 0.0f) * (dth * p.z * s.phi_max);
    const V3 d2pdvv = p * (-dth * dth);
    const float E = dot(dpdu, dpdu), F = dot(dpdu, dpdv), G = dot(dpdv, dpdv);
    const V3 N = normalize(cross(dpdu, dpdv));
    const float e = dot(N, d2pduu), f = dot(N, d2pduv), g = dot(N, d2pdvv);
    const float inv_egf2 = 1.0f / (E * G - F * F);
    const V3 du = dpdu * ((f * F - e * G) * inv_egf2) + dpdv * ((e * F - f * E) * inv_egf2);
    const V3 dv = dpdu * ((g * F - f * G) * inv_egf2) + dpdv * ((f * F - g * E) * inv_egf2);
    *dndu = transform_normal_inv(s.w2o, du); *dndv = transform_normal_inv(s.w2o, dv);
}

}  // namespace ftn
