// TEST INFRASTRUCTURE ONLY (see ref_math.h).  Reflection models, Fresnel, sampling
// routines, 1D/2D distributions and the level-0 env-map lookup, restated from the reference.
#pragma once
#include "ref_scene.h"

namespace ref {

// ---- sampling.rs ---------------------------------------------------------------------------
inline void concentric_sample_disk(Float u0, Float u1, Float* dx, Float* dy) {  // :5-19
    Float ox = 2.0f * u0 - 1.0f, oy = 2.0f * u1 - 1.0f;
    if (ox == 0.0f && oy == 0.0f) { *dx = 0.0f; *dy = 0.0f; return; }
    Float theta, r;
    if (std::fabs(ox) > std::fabs(oy)) { theta = FRAC_PI_4 * (oy / ox); r = ox; }
    else { theta = FRAC_PI_2 - FRAC_PI_4 * (ox / oy); r = oy; }
    *dx = r * std::cos(theta); *dy = r * std::sin(theta);
}
inline Vec3 cosine_sample_hemisphere(Float u0, Float u1) {  // :21-25
    Float dx, dy; concentric_sample_disk(u0, u1, &dx, &dy);
    Float z = std::sqrt(fmax_(0.0f, 1.0f - dx * dx - dy * dy));
    return Vec3(dx, dy, z);
}
inline Float power_heuristic(uint32_t nf, Float f_pdf, uint32_t ng, Float g_pdf) {  // :53-57
    Float f = (Float)nf * f_pdf, g = (Float)ng * g_pdf;
    return (f * f) / (f * f + g * g);
}

// sampling.rs:66-81
template <class F> inline size_t search_sorted(size_t size, F key) {
    size_t first = 0, len = size;
    while (len > 0) {
        size_t half = len >> 1, middle = first + half;
        if (key(middle)) { first = middle + 1; len -= half + 1; }
        else len = half;
    }
    size_t r = first - 1;            // usize; first >= 1 because cdf[0] = 0 <= u (release build would wrap)
    size_t hi = size - 2;
    if (r > hi) r = hi;              // .clamp(0, size - 2)
    return r;
}

struct Distribution1D {  // sampling.rs:59-135
    std::vector<Float> func, cdf;
    Float func_integral;
    void init(const Float* f, size_t n) {
        func.assign(f, f + n);
        cdf.assign(n + 1, 0.0f);
        for (size_t i = 1; i < n + 1; ++i) cdf[i] = cdf[i - 1] + (func[i - 1] / (Float)n);
        func_integral = cdf[n];
        if (func_integral == 0.0f) { for (size_t i = 1; i < n + 1; ++i) cdf[i] = (Float)i / (Float)n; }
        else { for (size_t i = 1; i < n + 1; ++i) cdf[i] /= func_integral; }
    }
    void sample_continuous(Float u, Float* x, Float* pdf, size_t* idx_out) const {  // :121-134
        size_t idx = search_sorted(cdf.size(), [&](size_t i) { return cdf[i] <= u; });
        Float du = u - cdf[idx];
        if (cdf[idx + 1] - cdf[idx] > 0.0f) du /= cdf[idx + 1] - cdf[idx];
        *pdf = func[idx] / func_integral;
        *x = ((Float)idx + du) / (Float)func.size();
        *idx_out = idx;
    }
};
struct Distribution2D {  // sampling.rs:137-180
    std::vector<Distribution1D> cond;
    Distribution1D marginal;
    void init(const Float* func, size_t nu, size_t nv) {
        cond.resize(nv);
        std::vector<Float> mf(nv);
        for (size_t v = 0; v < nv; ++v) { cond[v].init(func + v * nu, nu); mf[v] = cond[v].func_integral; }
        marginal.init(mf.data(), nv);
    }
    void sample_continuous(Float u0, Float u1, Float* d0, Float* d1, Float* pdf) const {  // :163-169
        Float pdf1, pdf0; size_t v_idx, dummy;
        marginal.sample_continuous(u1, d1, &pdf1, &v_idx);
        cond[v_idx].sample_continuous(u0, d0, &pdf0, &dummy);
        *pdf = pdf0 * pdf1;
    }
    Float pdf(Float px, Float py) const {  // :171-179  (`as usize` saturates: negative/NaN -> 0)
        size_t u_len = cond[0].func.size(), v_len = marginal.func.size();
        Float fu = px * (Float)u_len, fv = py * (Float)v_len;
        size_t iu = (fu > 0.0f) ? (size_t)fu : 0; if (iu > u_len - 1) iu = u_len - 1;
        size_t iv = (fv > 0.0f) ? (size_t)fv : 0; if (iv > v_len - 1) iv = v_len - 1;
        return cond[iv].func[iu] / marginal.func_integral;
    }
};

// ---- level-0 MIPMap lookups, mipmap.rs:245-312 (ImageWrap::Repeat) --------------------------
// InfiniteAreaLight only ever reaches `triangle(0, st)` or texel(levels-1,0,0) of a 1x1 map
// (SURVEY section 5 note 1); levels >= 1 (built by the absent `resize` crate) never contribute.
struct EnvMap {
    int w, h;
    std::vector<Float> texels;   // rgb
    Spectrum texel(int s, int t) const {  // :297-312 rem_euclid wrap
        int ss = ((s % w) + w) % w, tt = ((t % h) + h) % h;
        const Float* p = &texels[3 * ((size_t)tt * w + ss)];
        return Spectrum(p[0], p[1], p[2]);
    }
    Spectrum triangle0(Float sx, Float ty) const {  // :265-279
        Float s = sx * (Float)w - 0.5f, t = ty * (Float)h - 0.5f;
        int s0 = (int)std::floor(s), t0 = (int)std::floor(t);
        Float ds = s - (Float)s0, dt = t - (Float)t0;
        return texel(s0, t0) * (1.0f - ds) * (1.0f - dt)
             + texel(s0, t0 + 1) * (1.0f - ds) * dt
             + texel(s0 + 1, t0) * ds * (1.0f - dt)
             + texel(s0 + 1, t0 + 1) * ds * dt;
    }
    int levels() const { int m = std::max(w, h); int l = 0; while ((1 << (l + 1)) <= m) ++l; return 1 + l; }  // :103
    // lookup_trilinear_width, :245-257, for the two widths the env light uses.
    Spectrum lookup_width(Float sx, Float ty, Float width) const {
        Float level = (Float)levels() - 1.0f + std::log2(fmax_(width, 1.0e-8f));
        if (level < 0.0f) return triangle0(sx, ty);
        if (level >= (Float)(levels() - 1)) return texel(0, 0);   // only reachable here for a 1x1 map
        // power-of-two map, filter = 1/max(w,h): level == 0, delta == 0 => (1-0)*tri(0) + 0*tri(1)
        return triangle0(sx, ty);
    }
};

// ---- fresnel.rs -------------------------------------------------------------------------------
inline Float fresnel_dielectric(Float cos_theta_i, Float eta_i, Float eta_t) {  // :4-22
    cos_theta_i = clampf(cos_theta_i, -1.0f, 1.0f);
    bool entering = cos_theta_i > 0.0f;
    if (!entering) { std::swap(eta_i, eta_t); cos_theta_i = std::fabs(cos_theta_i); }
    Float sin_theta_i = std::sqrt(fmax_(1.0f - cos_theta_i * cos_theta_i, 0.0f));
    Float sin_theta_t = eta_i / eta_t * sin_theta_i;
    if (sin_theta_t >= 1.0f) return 1.0f;
    Float cos_theta_t = std::sqrt(fmax_(1.0f - sin_theta_t * sin_theta_t, 0.0f));
    Float r_parallel = ((eta_t * cos_theta_i) - (eta_i * cos_theta_t)) / ((eta_t * cos_theta_i) + (eta_i * cos_theta_t));
    Float r_perp = ((eta_i * cos_theta_i) - (eta_t * cos_theta_t)) / ((eta_i * cos_theta_i) + (eta_t * cos_theta_t));
    return (r_parallel * r_parallel + r_perp * r_perp) / 2.0f;
}
inline Spectrum fresnel_conductor(Float cos_theta_i, Spectrum eta_i, Spectrum eta_t, Spectrum k) {  // :25-48
    cos_theta_i = clampf(cos_theta_i, -1.0f, 1.0f);
    Spectrum eta = eta_t / eta_i, eta_k = k / eta_i;
    Float cos2 = cos_theta_i * cos_theta_i, sin2 = 1.0f - cos2;
    Spectrum eta2 = eta * eta, eta_k2 = eta_k * eta_k;
    Spectrum t0 = eta2 - eta_k2 - sin2;
    Spectrum a2plusb2 = ssqrt(t0 * t0 + 4.0f * eta2 * eta_k2);
    Spectrum t1 = a2plusb2 + cos2;
    Spectrum a = ssqrt(0.5f * (a2plusb2 + t0));
    Spectrum t2 = 2.0f * cos_theta_i * a;
    Spectrum Rs = (t1 - t2) / (t1 + t2);
    Spectrum t3 = cos2 * a2plusb2 + sin2 * sin2;
    Spectrum t4 = t2 * sin2;
    Spectrum Rp = Rs * (t3 - t4) / (t3 + t4);
    return 0.5f * (Rp + Rs);
}

// ---- reflection/mod.rs trig helpers :24-68 -----------------------------------------------------
inline Float cos_theta(Vec3 w) { return w.z; }
inline Float cos2_theta(Vec3 w) { return w.z * w.z; }
inline Float abs_cos_theta(Vec3 w) { return std::fabs(w.z); }
inline Float sin2_theta(Vec3 w) { return fmax_(0.0f, 1.0f - cos2_theta(w)); }
inline Float sin_theta(Vec3 w) { return std::sqrt(sin2_theta(w)); }
inline Float tan_theta(Vec3 w) { return sin_theta(w) / cos_theta(w); }
inline Float tan2_theta(Vec3 w) { return sin2_theta(w) / cos2_theta(w); }
inline Float cos_phi(Vec3 w) { Float s = sin_theta(w); return (s == 0.0f) ? 1.0f : clampf(w.x / s, -1.0f, 1.0f); }
inline Float sin_phi(Vec3 w) { Float s = sin_theta(w); return (s == 0.0f) ? 0.0f : clampf(w.y / s, -1.0f, 1.0f); }
inline Float cos2_phi(Vec3 w) { return cos_phi(w) * cos_phi(w); }
inline Float sin2_phi(Vec3 w) { return sin_phi(w) * sin_phi(w); }
inline Vec3 reflect(Vec3 wo, Vec3 n) { return -wo + 2.0f * dot(wo, n) * n; }                 // :80-82
inline bool same_hemisphere(Vec3 a, Vec3 b) { return sign_positive(a.z) == sign_positive(b.z); }  // :84-86

// reflection/mod.rs:14-22
enum { BXDF_REFLECTION = 1, BXDF_TRANSMISSION = 2, BXDF_DIFFUSE = 4, BXDF_GLOSSY = 8, BXDF_SPECULAR = 16, BXDF_ALL = 31 };

// microfacet.rs:40-45 / 125-127
inline Float roughness_to_alpha(Float roughness) {
    Float rough = fmax_(roughness, 1.0e-3f);
    Float x = std::log(rough);
    return 1.62142f + 0.819955f * x + 0.1734f * x * x + 0.0171201f * x * x * x + 0.000640711f * x * x * x * x;
}

struct ScatterSample { Spectrum f; Vec3 wi; Float pdf; int sampled_type; };

// One lobe: Lambert (reflection/mod.rs:116-162) or Torrance-Sparrow with Trowbridge-Reitz
// (reflection/mod.rs:301-361, microfacet.rs:119-186) and a conductor or dielectric Fresnel.
struct BxDF {
    int kind;            // 0 lambert, 1 microfacet, 2 specular reflection with FresnelNoOp (reflection/mod.rs:165-197), 3 Oren-Nayar (:252-296),
                         // 4 MicrofacetTransmission (:365-436; t in `r`, eta_a / eta_b in d_eta_i / d_eta_t, TransportMode::Radiance)
    Float on_a, on_b;    // Oren-Nayar coefficients
    Spectrum r;
    Float alpha_x, alpha_y;
    int fresnel;         // 0 conductor, 1 dielectric
    Spectrum eta_i, eta_t, k;
    Float d_eta_i, d_eta_t;

    int get_type() const {
        return (kind == 0 || kind == 3) ? (BXDF_REFLECTION | BXDF_DIFFUSE) : kind == 1 ? (BXDF_REFLECTION | BXDF_GLOSSY)
             : kind == 4 ? (BXDF_TRANSMISSION | BXDF_GLOSSY) : (BXDF_REFLECTION | BXDF_SPECULAR);
    }
    Float mt_get_eta(Vec3 wo) const { return (cos_theta(wo) > 0.0f) ? d_eta_t / d_eta_i : d_eta_i / d_eta_t; }   // :376-378
    bool matches(int flags) const { return (flags & get_type()) == get_type(); }

    Spectrum fresnel_eval(Float cos_i) const {
        if (fresnel == 0) return fresnel_conductor(std::fabs(cos_i), eta_i, eta_t, k);   // fresnel.rs:70-73
        return Spectrum(fresnel_dielectric(cos_i, d_eta_i, d_eta_t));                     // fresnel.rs:90-94
    }
    Float tr_d(Vec3 wh) const {  // microfacet.rs:135-146
        Float t2 = tan2_theta(wh);
        if (std::isinf(t2)) return 0.0f;
        Float cos4 = cos2_theta(wh) * cos2_theta(wh);
        Float e = (cos2_phi(wh) / (alpha_x * alpha_x) + sin2_phi(wh) / (alpha_y * alpha_y)) * t2;
        return 1.0f / (PI * alpha_x * alpha_y * cos4 * (1.0f + e) * (1.0f + e));
    }
    Float tr_lambda(Vec3 w) const {  // microfacet.rs:148-160
        Float att = std::fabs(tan_theta(w));
        if (std::isinf(att)) return 0.0f;
        Float alpha = std::sqrt(cos2_phi(w) * alpha_x * alpha_x + sin2_phi(w) * alpha_y * alpha_y);
        Float a2t2 = (alpha * att) * (alpha * att);
        return (-1.0f + std::sqrt(1.0f + a2t2)) / 2.0f;
    }
    Float tr_g(Vec3 wo, Vec3 wi) const { return 1.0f / (1.0f + tr_lambda(wo) + tr_lambda(wi)); }  // :21-23
    Float tr_pdf(Vec3 /*wo*/, Vec3 wh) const { return tr_d(wh) * abs_cos_theta(wh); }              // :28-31
    Vec3 tr_sample_wh(Vec3 wo, Float u0, Float u1) const {  // microfacet.rs:162-186
        Float cos_t, phi;
        if (alpha_x == alpha_y) {
            Float tan_theta2 = (alpha_x * alpha_x) * u0 / (1.0f - u0);
            cos_t = 1.0f / std::sqrt(1.0f + tan_theta2);
            phi = 2.0f * PI * u1;
        } else {
            phi = std::atan(alpha_y / alpha_x * std::tan(2.0f * PI * u1 + 0.5f * PI));
            if (u1 > 0.5f) phi += PI;
            Float sp = std::sin(phi), cp = std::cos(phi);
            Float alpha2 = 1.0f / ((cp * cp) / (alpha_x * alpha_x) + (sp * sp) / (alpha_y * alpha_y));
            Float tan_theta2 = alpha2 * u0 / (1.0f - u0);
            cos_t = 1.0f / std::sqrt(1.0f + tan_theta2);
        }
        Float sin_t = std::sqrt(fmax_(0.0f, 1.0f - (cos_t * cos_t)));
        Vec3 wh = spherical_direction(sin_t, cos_t, phi);
        return same_hemisphere(wo, wh) ? wh : -wh;
    }

    Spectrum f(Vec3 wo, Vec3 wi) const {
        if (kind == 0) return r * FRAC_1_PI;   // reflection/mod.rs:159-161
        if (kind == 2) return Spectrum(0.0f);   // :181-183
        if (kind == 4) {   // MicrofacetTransmission::f, :386-404
            if (same_hemisphere(wo, wi)) return Spectrum(0.0f);
            Float cos_theta_o = cos_theta(wo), cos_theta_i = cos_theta(wi);
            if (cos_theta_o == 0.0f || cos_theta_i == 0.0f) return Spectrum(0.0f);
            Float eta = mt_get_eta(wo);
            Vec3 wh = normalize(wo + wi * eta);
            if (wh.z < 0.0f) wh = -wh;
            Float F = fresnel_dielectric(dot(wo, wh), d_eta_i, d_eta_t);
            Float sqrt_denom = dot(wo, wh) + eta * dot(wi, wh);
            Float factor = 1.0f / eta;   // TransportMode::Radiance
            return (Spectrum(1.0f) - Spectrum(F)) * r *
                   std::fabs(tr_d(wh) * tr_g(wo, wi) * (eta * eta) * abs_dot(wi, wh) * abs_dot(wo, wh) * (factor * factor)
                             / (cos_theta_i * cos_theta_o * (sqrt_denom * sqrt_denom)));
        }
        if (kind == 3) {   // OrenNayar::f, reflection/mod.rs:274-296
            Float sin_theta_i = sin_theta(wi), sin_theta_o = sin_theta(wo);
            Float max_cos = 0.0f;
            if (sin_theta_i > 1.0e-4f && sin_theta_o > 1.0e-4f) {
                Float sin_phi_i = sin_phi(wi), cos_phi_i = cos_phi(wi), sin_phi_o = sin_phi(wo), cos_phi_o = cos_phi(wo);
                Float d_cos = cos_phi_i * cos_phi_o + sin_phi_i * sin_phi_o;
                max_cos = fmax_(0.0f, d_cos);
            }
            Float sin_alpha, tan_beta;
            if (abs_cos_theta(wi) > abs_cos_theta(wo)) { sin_alpha = sin_theta_o; tan_beta = sin_theta_i / abs_cos_theta(wi); }
            else { sin_alpha = sin_theta_i; tan_beta = sin_theta_o / abs_cos_theta(wo); }
            return r * FRAC_1_PI * (on_a + (on_b * max_cos * sin_alpha * tan_beta));
        }
        // reflection/mod.rs:318-336
        Float cos_o = abs_cos_theta(wo), cos_i = abs_cos_theta(wi);
        Vec3 wh = wi + wo;
        if (cos_i == 0.0f || cos_o == 0.0f || (wh.x == 0.0f && wh.y == 0.0f && wh.z == 0.0f)) return Spectrum(0.0f);
        wh = normalize(wh);
        Spectrum F = fresnel_eval(dot(wi, faceforward(wh, Vec3(0.0f, 0.0f, 1.0f))));
        return r * tr_d(wh) * tr_g(wo, wi) * F / (4.0f * cos_i * cos_o);
    }
    Float pdf(Vec3 wo, Vec3 wi) const {
        if (kind == 0 || kind == 3) return same_hemisphere(wo, wi) ? abs_cos_theta(wi) * FRAC_1_PI : 0.0f;   // DefaultSampleF :140-146
        if (kind == 2) return 0.0f;   // :194-196
        if (kind == 4) {   // :426-435
            if (same_hemisphere(wo, wi)) return 0.0f;
            Float eta = mt_get_eta(wo);
            Vec3 wh = normalize(wo + wi * eta);
            Float sqrt_denom = dot(wo, wh) + eta * dot(wi, wh);
            Float dwh_dwi = std::fabs(((eta * eta) * dot(wi, wh)) / (sqrt_denom * sqrt_denom));
            return tr_pdf(wo, wh) * dwh_dwi;
        }
        if (!same_hemisphere(wo, wi)) return 0.0f;   // :354-360
        Vec3 wh = normalize(wo + wi);
        return tr_pdf(wo, wh) / (4.0f * dot(wo, wh));
    }
    bool sample_f(Vec3 wo, Float u0, Float u1, ScatterSample* s) const {
        if (kind == 0 || kind == 3) {   // DefaultSampleF :131-138
            Vec3 wi = cosine_sample_hemisphere(u0, u1);
            if (wo.z < 0.0f) wi.z *= -1.0f;
            s->pdf = pdf(wo, wi); s->f = f(wo, wi); s->wi = wi; s->sampled_type = get_type();
            return true;
        }
        if (kind == 2) {   // :185-192; FresnelNoOp::evaluate is Spectrum::uniform(1.0) (fresnel.rs)
            Vec3 wi(-wo.x, -wo.y, wo.z);
            s->pdf = 1.0f; s->f = Spectrum(1.0f) * r / abs_cos_theta(wi); s->wi = wi; s->sampled_type = get_type();
            return true;
        }
        if (kind == 4) {   // :406-424
            if (wo.z == 0.0f) return false;
            Vec3 wh = tr_sample_wh(wo, u0, u1);
            if (dot(wo, wh) < 0.0f) return false;
            Float eta = mt_get_eta(-wo);   // "NOTE: this inverts the eta fraction"
            // refract, reflection/mod.rs:70-78
            Float cos_theta_i = dot(wh, wo);
            Float sin2_theta_i = fmax_(0.0f, 1.0f - cos_theta_i * cos_theta_i);
            Float sin2_theta_t = eta * eta * sin2_theta_i;
            if (sin2_theta_t >= 1.0f) return false;
            Float cos_theta_t = std::sqrt(1.0f - sin2_theta_t);
            Vec3 wi = eta * -wo + (eta * cos_theta_i - cos_theta_t) * wh;
            s->f = f(wo, wi); s->wi = wi; s->pdf = pdf(wo, wi); s->sampled_type = get_type();
            return true;
        }
        Vec3 wh = tr_sample_wh(wo, u0, u1);   // :338-352
        Vec3 wi = reflect(wo, wh);
        if (!same_hemisphere(wo, wi)) return false;
        s->pdf = tr_pdf(wo, wh) / (4.0f * dot(wo, wh));
        s->f = f(wo, wi); s->wi = wi; s->sampled_type = get_type();
        return true;
    }
};

// reflection/bsdf.rs:8-148
struct Bsdf {
    Vec3 ns, ng, ss, ts;
    BxDF bxdfs[8]; int n;
    void init(const SurfaceInteraction& si) {  // :31-46
        ns = si.shading_n; ng = si.hit.n;
        ss = normalize(si.shading_dpdu);
        ts = normalize(cross(ns, ss));
        n = 0;
    }
    void add(const BxDF& b) { bxdfs[n++] = b; }
    int num_components(int flags) const { int c = 0; for (int i = 0; i < n; ++i) if (bxdfs[i].matches(flags)) ++c; return c; }
    Vec3 world_to_local(Vec3 v) const { return Vec3(dot(v, ss), dot(v, ts), dot(v, ns)); }   // :56-58
    Vec3 local_to_world(Vec3 v) const {   // :60-65
        return Vec3(ss.x * v.x + ts.x * v.y + ns.x * v.z, ss.y * v.x + ts.y * v.y + ns.y * v.z, ss.z * v.x + ts.z * v.y + ns.z * v.z);
    }
    Spectrum f(Vec3 wo_w, Vec3 wi_w, int flags) const {   // :67-82
        Vec3 wi = world_to_local(wi_w), wo = world_to_local(wo_w);
        if (wo.z == 0.0f) return Spectrum(0.0f);
        bool reflect_ = dot(wi_w, ng) * dot(wo_w, ng) > 0.0f;
        Spectrum sum(0.0f);
        for (int i = 0; i < n; ++i) {
            if (!bxdfs[i].matches(flags)) continue;
            int ty = bxdfs[i].get_type();
            if ((reflect_ && (ty & BXDF_REFLECTION)) || (!reflect_ && (ty & BXDF_TRANSMISSION))) sum = sum + bxdfs[i].f(wo, wi);
        }
        return sum;
    }
    bool sample_f(Vec3 wo_w, Float u0, Float u1, int flags, ScatterSample* out) const {   // :85-129
        Float matching = (Float)num_components(flags);
        if (matching == 0.0f) return false;
        int comp = (int)fmin_(std::floor(u0 * matching), matching - 1.0f);
        const BxDF* bxdf = nullptr; int cnt = comp;
        for (int i = 0; i < n; ++i) if (bxdfs[i].matches(flags)) { if (cnt-- == 0) { bxdf = &bxdfs[i]; break; } }
        Float ur0 = u0 * matching - (Float)comp, ur1 = u1;
        Vec3 wo = world_to_local(wo_w);
        ScatterSample s;
        if (!bxdf->sample_f(wo, ur0, ur1, &s)) return false;
        if (s.pdf == 0.0f) return false;
        Vec3 wi = s.wi;
        Vec3 wi_w = local_to_world(wi);
        Float pdf = s.pdf; Spectrum fval = s.f;
        if (!(bxdf->get_type() & BXDF_SPECULAR) && matching > 1.0f) {
            for (int i = 0; i < n; ++i) if (bxdfs[i].matches(flags) && &bxdfs[i] != bxdf) pdf += bxdfs[i].pdf(wo, wi);
        }
        if (matching > 1.0f) pdf /= matching;
        if (!(bxdf->get_type() & BXDF_SPECULAR)) {
            bool reflect_ = dot(wi_w, ng) * dot(wo_w, ng) > 0.0f;
            Spectrum sum(0.0f);
            for (int i = 0; i < n; ++i) {
                if (!bxdfs[i].matches(flags)) continue;
                int ty = bxdfs[i].get_type();
                if ((reflect_ && (ty & BXDF_REFLECTION)) || (!reflect_ && (ty & BXDF_TRANSMISSION))) sum = sum + bxdfs[i].f(wo, wi);
            }
            fval = sum;
        }
        out->f = fval; out->wi = wi_w; out->pdf = pdf; out->sampled_type = s.sampled_type;
        return true;
    }
    Float pdf(Vec3 wo_w, Vec3 wi_w, int flags) const {   // :131-144
        Vec3 wo = world_to_local(wo_w), wi = world_to_local(wi_w);
        if (wo.z == 0.0f) return 0.0f;
        Float n_matching = (Float)num_components(flags);
        Float p = 0.0f;
        for (int i = 0; i < n; ++i) if (bxdfs[i].matches(flags)) p += bxdfs[i].pdf(wo, wi);
        return (n_matching > 0.0f) ? p / n_matching : 0.0f;
    }
};

// Kd through its texture: ConstantTexture (texture/mod.rs:34-42), ImageTexture (image.rs), Checkerboard2DTexture with
// AAMethod::None (checkerboard.rs:50-64) or UVTexture (uv.rs:18-23), via UVMapping (mapping.rs:40-52).
inline Spectrum evaluate_kd(const Material& m, const SurfaceInteraction& si) {
    if (m.kd_texture == 0) return m.type == 3 ? m.kr : m.kd;   // the mirror's textured parameter is Kr (mirror.rs:23)
    Float s = m.uv_scale[0] * si.uv[0] + m.uv_delta[0], t = m.uv_scale[1] * si.uv[1] + m.uv_delta[1];
    if (m.kd_texture == 1) return (((int)std::floor(s) + (int)std::floor(t)) % 2 == 0) ? m.tex1 : m.tex2;
    if (m.kd_texture == 3) {   // ImageTexture::evaluate, image.rs:30-33, with UVMapping's dst_dx / dst_dy (mapping.rs:43-44)
        Float dst_dx[2] = {m.uv_scale[0] * si.dudx, m.uv_scale[1] * si.dvdx};
        Float dst_dy[2] = {m.uv_scale[0] * si.dudy, m.uv_scale[1] * si.dvdy};
        return m.image->lookup_trilinear(s, t, dst_dx, dst_dy);
    }
    return Spectrum(s - std::floor(s), t - std::floor(t), 0.0f);
}

// Texture::evaluate of a texture-table entry
inline Spectrum evaluate_texture(const TextureDef& t, const SurfaceInteraction& si) {
    if (t.type == 0) return t.value;
    Float s = t.uv_scale[0] * si.uv[0] + t.uv_delta[0], tt = t.uv_scale[1] * si.uv[1] + t.uv_delta[1];
    if (t.type == 1) return (((int)std::floor(s) + (int)std::floor(tt)) % 2 == 0) ? t.tex1 : t.tex2;
    if (t.type == 3) {
        Float dst_dx[2] = {t.uv_scale[0] * si.dudx, t.uv_scale[1] * si.dvdx};
        Float dst_dy[2] = {t.uv_scale[0] * si.dudy, t.uv_scale[1] * si.dvdy};
        return t.image->lookup_trilinear(s, tt, dst_dx, dst_dy);
    }
    return Spectrum(s - std::floor(s), tt - std::floor(tt), 0.0f);
}
inline bool param_textured(const Material& m, int p) { return m.table && m.ptex[p] != 0; }
inline Spectrum param_spectrum(const Material& m, int p, Spectrum constant, const SurfaceInteraction& si) {
    return param_textured(m, p) ? evaluate_texture((*m.table)[m.ptex[p] - 1], si) : constant;
}
inline Float param_float(const Material& m, int p, Float constant, const SurfaceInteraction& si) {
    return param_textured(m, p) ? evaluate_texture((*m.table)[m.ptex[p] - 1], si).c[0] : constant;
}
// Kd (matte, plastic) / Kr (mirror): the texture table, else the inline slot of ABI v2, else the constant
inline Spectrum param_kd(const Material& m, int p, const SurfaceInteraction& si) {
    return param_textured(m, p) ? evaluate_texture((*m.table)[m.ptex[p] - 1], si) : evaluate_kd(m, si);
}

// material/{matte,metal,plastic,mirror,glass}.rs compute_scattering_functions; false where the reference hits todo!()
inline bool compute_scattering_functions(const Material& m, const SurfaceInteraction& si, Bsdf* bsdf) {
    bsdf->init(si);
    if (m.type == 0) {   // matte.rs:36-52
        Spectrum r = param_kd(m, FTN_PARAM_KD, si).clamp_positive();
        Float sigma = clampf(param_float(m, FTN_PARAM_SIGMA, m.sigma, si), 0.0f, 90.0f);
        if (!r.is_black()) {
            BxDF b{}; b.r = r;
            if (sigma == 0.0f) b.kind = 0;
            else {   // OrenNayar::new(r, Deg(sigma)), reflection/mod.rs:259-267; cgmath Deg -> Rad: deg * (PI / 180)
                Float s = sigma * (Float)(3.14159265358979323846 / 180.0), s2 = s * s;   // the factor is an f64 constant cast to f32 in cgmath
                b.kind = 3; b.on_a = 1.0f - (s2 / (2.0f * (s2 + 0.33f))); b.on_b = 0.45f * s2 / (s2 + 0.09f);
            }
            bsdf->add(b);
        }
    } else if (m.type == 1) {   // metal.rs:38-65
        Float ur = param_float(m, FTN_PARAM_UROUGHNESS, m.u_rough, si), vr = param_float(m, FTN_PARAM_VROUGHNESS, m.v_rough, si);
        if (m.remap) { ur = roughness_to_alpha(ur); vr = roughness_to_alpha(vr); }
        BxDF b{}; b.kind = 1; b.r = Spectrum(1.0f); b.alpha_x = ur; b.alpha_y = vr;
        b.fresnel = 0; b.eta_i = Spectrum(1.0f); b.eta_t = param_spectrum(m, FTN_PARAM_ETA, m.eta, si); b.k = param_spectrum(m, FTN_PARAM_K, m.k, si);
        bsdf->add(b);
    } else if (m.type == 4) {   // glass.rs:52-96; the specular branch is todo!() (path) or a two-branch recursion (direct lighting): rejected at scene creation
        Float eta = param_float(m, FTN_PARAM_INDEX, m.eta.c[0], si);
        Spectrum r = param_spectrum(m, FTN_PARAM_KR, m.kr, si).clamp_positive(), t = param_spectrum(m, FTN_PARAM_KT, m.kt, si).clamp_positive();
        Float ur = param_float(m, FTN_PARAM_UROUGHNESS, m.u_rough, si), vr = param_float(m, FTN_PARAM_VROUGHNESS, m.v_rough, si);
        if (m.remap) { ur = roughness_to_alpha(ur); vr = roughness_to_alpha(vr); }
        if (ur == 0.0f && vr == 0.0f) return false;   // glass.rs:64-67: todo!("FresnelSpecular")
        if (!r.is_black()) {
            BxDF b{}; b.kind = 1; b.r = r; b.alpha_x = ur; b.alpha_y = vr; b.fresnel = 1; b.d_eta_i = 1.0f; b.d_eta_t = eta;
            bsdf->add(b);
        }
        if (!t.is_black()) {
            BxDF b{}; b.kind = 4; b.r = t; b.alpha_x = ur; b.alpha_y = vr; b.d_eta_i = 1.0f; b.d_eta_t = eta;
            bsdf->add(b);
        }
    } else if (m.type == 3) {   // mirror.rs:21-30
        Spectrum r = param_kd(m, FTN_PARAM_KR, si).clamp_positive();
        if (!r.is_black()) { BxDF b{}; b.kind = 2; b.r = r; bsdf->add(b); }
    } else {   // plastic.rs:24-48
        Spectrum kd = param_kd(m, FTN_PARAM_KD, si), ks = param_spectrum(m, FTN_PARAM_KS, m.ks, si);
        if (!kd.is_black()) { BxDF b{}; b.kind = 0; b.r = kd; bsdf->add(b); }
        if (!ks.is_black()) {
            Float rough = param_float(m, FTN_PARAM_UROUGHNESS, m.u_rough, si);
            if (m.remap) rough = roughness_to_alpha(rough);
            BxDF b{}; b.kind = 1; b.r = ks; b.alpha_x = rough; b.alpha_y = rough;
            b.fresnel = 1; b.d_eta_i = 1.5f; b.d_eta_t = 1.0f;
            bsdf->add(b);
        }
    }
    return true;
}

}  // namespace ref
