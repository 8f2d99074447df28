// ORACLE — test infrastructure only (see zmath.hpp). Material side of the reference, uniform-parameter
// subset:
//   bxdf.Result / Sample / Path     src/core/scene/material/bxdf.zig
//   sample Base                     src/core/scene/material/sample_base.zig
//   fresnel.Schlick                 src/core/scene/material/fresnel.zig:5-29
//   ggx.Iso / Aniso, dspbrMicroEc   src/core/scene/material/ggx.zig
//   diffuse.Micro                   src/core/scene/material/diffuse.zig:42-116
//   Substitute Material.sample      src/core/scene/material/substitute/substitute_material.zig:114-221
//   Substitute Sample               src/core/scene/material/substitute/substitute_sample.zig:38-410
//   Light material / Emittance      src/core/scene/material/light/light_material.zig, light/emittance.zig:29-59
//   Glass Material.sample / Sample  src/core/scene/material/glass/glass_material.zig:46-73, glass_sample.zig:32-537
//                                   (thick glass: thickness == 0, abbe == 0)
#pragma once

#include "zsampler.hpp"
#include "zscene.hpp"

namespace zo {

namespace bxdf {

struct Result {
    Vec4f reflection;
    float pdf;
    static Result empty() { return {splat(0.f), 0.f}; }
};

enum class Scattering : uint8_t { Diffuse, Glossy, Specular, None };
enum class Event : uint8_t { Reflection, Transmission, Straight };

struct Path {  // bxdf.zig:42-73
    float      reg_alpha;
    Scattering scattering;
    Event      event;

    static Path diffuseReflection() { return {1.f, Scattering::Diffuse, Event::Reflection}; }
    static Path singularReflection() { return {0.f, Scattering::Specular, Event::Reflection}; }
    static Path singularTransmission() { return {0.f, Scattering::Specular, Event::Transmission}; }
    static Path reflection(float alpha, float specular_threshold) {
        return {alpha, alpha <= specular_threshold ? Scattering::Specular : Scattering::Glossy, Event::Reflection};
    }
    static Path transmission(float alpha, float specular_threshold) {
        return {alpha, alpha <= specular_threshold ? Scattering::Specular : Scattering::Glossy, Event::Transmission};
    }
    bool singular() const { return 0.f == reg_alpha; }
};

struct Sample {  // bxdf.zig:75-82
    Vec4f reflection;
    Vec4f wi;
    float pdf;
    float split_weight;
    float wavelength;
    Path  path;
};

}  // namespace bxdf

struct Renderstate {  // renderstate.zig:10-37 — the fields the in-scope materials read
    Trafo trafo;
    Vec4f p, geo_n, t, b, n, origin, uvw;
    float stochastic_r;
    float ior;
    float reg_weight;
    float reg_alpha;
    uint32_t prop, part;
    bool     primary;
    bool     caustics;
    int8_t   highest_priority;

    // renderstate.zig:58-68
    void regularizeAlpha(const float alpha[2], float specular_threshold, float out[2]) const {
        const float weight = reg_weight;
        if (0.f == weight || (alpha[0] <= specular_threshold && !caustics)) {
            out[0] = alpha[0];
            out[1] = alpha[1];
            return;
        }
        const float k = 1.f - weight * reg_alpha;
        out[0]        = 1.f - ((1.f - alpha[0]) * k);
        out[1]        = 1.f - ((1.f - alpha[1]) * k);
    }
};

struct SampleBase {  // sample_base.zig:14-104
    Frame frame;
    Vec4f geo_n, n, wo;
    float alpha[2];
    bool  translucent = false, can_evaluate = false, avoid_caustics = false, volumetric = false, lower_priority = false;

    bool sameHemisphere(Vec4f v) const { return dot3(geo_n, v) > 0.f; }
    bool avoidCausticsForce(bool force) const { return force || avoid_caustics; }
};

namespace fresnel {
struct Schlick {  // fresnel.zig:9-29
    Vec4f f0;
    Vec4f f(float wo_dot_h) const { return mulAdd(splat(pow5(1.f - wo_dot_h)), splat(1.f) - f0, f0); }
    static float IorToF0(float n0, float n1) {
        const float t = (n0 - n1) / (n0 + n1);
        return t * t;
    }
};
inline float schlick1(float wo_dot_h, float f0) { return std::fmaf(pow5(1.f - wo_dot_h), 1.f - f0, f0); }  // fresnel.zig:5-7
inline float dielectric(float cos_theta_i, float cos_theta_t, float eta_i, float eta_t) {  // fresnel.zig:31-43
    const float t0  = eta_t * cos_theta_i;
    const float t1  = eta_i * cos_theta_t;
    const float r_p = (t0 - t1) / (t0 + t1);
    const float t2  = eta_i * cos_theta_i;
    const float t3  = eta_t * cos_theta_t;
    const float r_o = (t2 - t3) / (t2 + t3);
    return 0.5f * (r_p * r_p + r_o * r_o);
}
}  // namespace fresnel

struct IoR {  // sample_base.zig:106-117
    float eta_t, eta_i;
    IoR   swapped(bool same_side) const { return same_side ? *this : IoR{eta_i, eta_t}; }
};

namespace ggx {

constexpr float MinRoughness = 0.01314f;                     // ggx.zig:14
constexpr float MinAlpha     = MinRoughness * MinRoughness;  // :15

inline float clampRoughness(float roughness) { return max(roughness, MinRoughness); }  // :48-50

struct Micro {
    Vec4f h;
    float n_dot_wi, h_dot_wi;
};

inline float pdfVisible(float d, float g1_wo) { return (0.5f * d) / g1_wo; }  // :437-439

// ggx.zig:34-46
inline Vec4f dspbrMicroEc(const GgxLuts& luts, Vec4f f0, float n_dot_wi, float n_dot_wo, float alpha) {
    const float e_wo  = luts.eM(n_dot_wo, alpha);
    const float e_wi  = luts.eM(n_dot_wi, alpha);
    const float e_avg = luts.eMAvg(alpha);

    const float m = ((1.f - e_wo) * (1.f - e_wi)) / (kPi * (1.f - e_avg));

    const Vec4f f_avg = mulAdd(splat(20.f / 21.f), f0, splat(1.f / 21.f));
    const Vec4f f     = ((f_avg * f_avg) * splat(e_avg)) / mulAdd(-f_avg, splat(1.f - e_avg), splat(1.f));
    return splat(m) * f;
}

// ggx.zig:30-32; E_s_inverse_max_f0 = 4 (ggx_integral.zig:1045)
inline float ilmEpDielectric(const GgxLuts& luts, float n_dot_wo, float alpha, float f0) { return 1.f / luts.eS(n_dot_wo, alpha, f0 * 4.f); }

// Aniso.sample, ggx.zig:393-409 (Dupuy & Benyoub spherical caps)
inline Vec4f sampleVndf(Vec4f wo, const float alpha[2], const float xi[2], const Frame& frame, float& n_dot_h) {
    const Vec4f wo_l = frame.worldToFrame(wo);
    const Vec4f v    = normalize3({{alpha[0] * wo_l[0], alpha[1] * wo_l[1], wo_l[2], 0.f}});

    const float phi       = (2.f * kPi) * xi[0];
    const float z         = std::fmaf(1.f - xi[1], 1.f + v[2], -v[2]);
    const float sin_theta = std::sqrt(saturate(1.f - z * z));
    const float x         = sin_theta * std::cos(phi);
    const float y         = sin_theta * std::sin(phi);

    const Vec4f h = Vec4f{{x, y, z, 0.f}} + v;
    const Vec4f m = normalize3({{alpha[0] * h[0], alpha[1] * h[1], h[2], 0.f}});

    n_dot_h = safe::clamp(m[2]);
    return frame.frameToWorld(m);
}

namespace iso {
inline float distribution(float n_dot_h, float a2) {  // :235-238
    const float d = std::fmaf(n_dot_h * n_dot_h, a2 - 1.f, 1.f);
    return a2 / (kPi * d * d);
}
inline void visibilityAndG1Wo(float n_dot_wi, float n_dot_wo, float alpha2, float out[2]) {  // :240-250
    const float t_wi = std::sqrt(std::fmaf(1.f - alpha2, n_dot_wi * n_dot_wi, alpha2));
    const float t_wo = std::sqrt(std::fmaf(1.f - alpha2, n_dot_wo * n_dot_wo, alpha2));
    out[0]           = 0.5f / (n_dot_wi * t_wo + n_dot_wo * t_wi);
    out[1]           = t_wo + n_dot_wo;
}

// Iso.reflectionF, :73-95
inline bxdf::Result reflection(Vec4f h, Vec4f n, float n_dot_wi, float n_dot_wo, float wo_dot_h, float alpha,
                               const fresnel::Schlick& fr) {
    const float alpha2  = alpha * alpha;
    const float n_dot_h = saturate(dot3(n, h));
    const float d       = distribution(n_dot_h, alpha2);
    float       g[2];
    visibilityAndG1Wo(n_dot_wi, n_dot_wo, alpha2, g);
    const Vec4f f = fr.f(wo_dot_h);
    return {splat(d * g[0]) * f, pdfVisible(d, g[1])};
}

// Iso.reflect, :97-126
inline Micro reflect(Vec4f wo, float n_dot_wo, float alpha, float specular_threshold, const float xi[2],
                     const fresnel::Schlick& fr, const Frame& frame, bxdf::Sample& result) {
    float       n_dot_h;
    const float a[2] = {alpha, alpha};
    const Vec4f h    = sampleVndf(wo, a, xi, frame, n_dot_h);

    const float wo_dot_h = safe::clampDot(wo, h);
    const Vec4f wi       = normalize3(mulAdd(splat(2.f * wo_dot_h), h, -wo));

    const float n_dot_wi = frame.clampNdot(wi);
    const float alpha2   = alpha * alpha;

    const float d = distribution(n_dot_h, alpha2);
    float       g[2];
    visibilityAndG1Wo(n_dot_wi, n_dot_wo, alpha2, g);
    const Vec4f f = fr.f(wo_dot_h);

    result.reflection = splat(d * g[0]) * f;
    result.wi         = wi;
    result.pdf        = pdfVisible(d, g[1]);
    result.path       = bxdf::Path::reflection(alpha, specular_threshold);
    return {h, n_dot_wi, wo_dot_h};
}

inline float gSmithCorrelated(float n_dot_wi, float n_dot_wo, float alpha2) {  // :252-257
    const float a = n_dot_wo * std::sqrt(std::fmaf(1.f - alpha2, n_dot_wi * n_dot_wi, alpha2));
    const float b = n_dot_wi * std::sqrt(std::fmaf(1.f - alpha2, n_dot_wo * n_dot_wo, alpha2));
    return (2.f * n_dot_wi * n_dot_wo) / (a + b);
}
inline float gGgx(float n_dot_v, float alpha2) {  // :447-449
    return (2.f * n_dot_v) / (n_dot_v + std::sqrt(alpha2 + (1.f - alpha2) * (n_dot_v * n_dot_v)));
}
inline float pdfVisibleRefract(float n_dot_wo, float wo_dot_h, float d, float alpha2) {  // :441-445
    const float g1 = gGgx(n_dot_wo, alpha2);
    return g1 * wo_dot_h * d / n_dot_wo;
}

struct ResultF {
    bxdf::Result r;
    float        f;  // fresnel term, lane 0
};

// Iso.reflectionF with a scalar f0, :73-95
inline ResultF reflectionF(Vec4f h, Vec4f n, float n_dot_wi, float n_dot_wo, float wo_dot_h, float alpha, float f0) {
    const float alpha2  = alpha * alpha;
    const float n_dot_h = saturate(dot3(n, h));
    const float d       = distribution(n_dot_h, alpha2);
    float       g[2];
    visibilityAndG1Wo(n_dot_wi, n_dot_wo, alpha2, g);
    const float f = fresnel::schlick1(wo_dot_h, f0);
    return {{splat(d * g[0]) * splat(f), pdfVisible(d, g[1])}, f};
}

// Iso.refractionF, :128-159
inline ResultF refractionF(float n_dot_wi, float n_dot_wo, float wi_dot_h, float wo_dot_h, float n_dot_h, float alpha, IoR ior,
                           float f0) {
    const float alpha2 = alpha * alpha;

    const float abs_wi_dot_h = safe::clampAbs(wi_dot_h);
    const float abs_wo_dot_h = safe::clampAbs(wo_dot_h);

    const float d = distribution(n_dot_h, alpha2);
    const float g = gSmithCorrelated(n_dot_wi, n_dot_wo, alpha2);

    const float cos_x = ior.eta_i > ior.eta_t ? abs_wi_dot_h : abs_wo_dot_h;
    const float f     = 1.f - fresnel::schlick1(cos_x, f0);

    const float sqr_eta_t = ior.eta_t * ior.eta_t;

    const float factor = (abs_wi_dot_h * abs_wo_dot_h) / (n_dot_wi * n_dot_wo);
    const float denom  = pow2(ior.eta_i * wo_dot_h + ior.eta_t * wi_dot_h);

    const float refr = d * g * f;
    const float refl = (factor * sqr_eta_t / denom) * refr;

    const float pdf = pdfVisibleRefract(n_dot_wo, abs_wo_dot_h, d, alpha2);
    return {{splat(refl), pdf * (abs_wi_dot_h * sqr_eta_t / denom)}, f};
}

// Iso.reflectNoFresnel, :161-187
inline float reflectNoFresnel(Vec4f wo, Vec4f h, float n_dot_wo, float n_dot_h, float wo_dot_h, float alpha, float specular_threshold,
                              const Frame& frame, bxdf::Sample& result) {
    const Vec4f wi = normalize3(mulAdd(splat(2.f * wo_dot_h), h, -wo));

    const float n_dot_wi = frame.clampNdot(wi);
    const float alpha2   = alpha * alpha;

    const float d = distribution(n_dot_h, alpha2);
    float       g[2];
    visibilityAndG1Wo(n_dot_wi, n_dot_wo, alpha2, g);

    result.reflection = splat(d * g[0]);
    result.wi         = wi;
    result.pdf        = pdfVisible(d, g[1]);
    result.path       = bxdf::Path::reflection(alpha, specular_threshold);
    return n_dot_wi;
}

// Iso.refractNoFresnel, :189-233
inline float refractNoFresnel(Vec4f wo, Vec4f h, float n_dot_wo, float n_dot_h, float wi_dot_h, float wo_dot_h, float alpha,
                              float specular_threshold, IoR ior, const Frame& frame, bxdf::Sample& result) {
    const float eta = ior.eta_i / ior.eta_t;

    const float abs_wi_dot_h = safe::clampAbs(wi_dot_h);
    const float abs_wo_dot_h = safe::clampAbs(wo_dot_h);

    const Vec4f wi = normalize3(splat(std::fmaf(eta, abs_wo_dot_h, -abs_wi_dot_h)) * h - splat(eta) * wo);

    const float n_dot_wi = frame.clampAbsNdot(wi);

    const float alpha2 = alpha * alpha;

    const float d = distribution(n_dot_h, alpha2);
    const float g = gSmithCorrelated(n_dot_wi, n_dot_wo, alpha2);

    const float refr      = d * g;
    const float factor    = (abs_wi_dot_h * abs_wo_dot_h) / (n_dot_wi * n_dot_wo);
    const float denom     = pow2(ior.eta_i * wo_dot_h + ior.eta_t * wi_dot_h);
    const float sqr_eta_t = ior.eta_t * ior.eta_t;
    const float pdf       = pdfVisibleRefract(n_dot_wo, abs_wo_dot_h, d, alpha2);

    result.reflection = splat((factor * sqr_eta_t / denom) * refr);
    result.wi         = wi;
    result.pdf        = pdf * (abs_wi_dot_h * sqr_eta_t / denom);
    result.path       = bxdf::Path::transmission(alpha, specular_threshold);
    return n_dot_wi;
}
}  // namespace iso

namespace aniso {
inline float distribution(float n_dot_h, float x_dot_h, float y_dot_h, const float a[2]) {  // :411-419
    const float a2x = a[0] * a[0];
    const float a2y = a[1] * a[1];
    const float x   = (x_dot_h * x_dot_h) / a2x;
    const float y   = (y_dot_h * y_dot_h) / a2y;
    const float d   = (x + y) + (n_dot_h * n_dot_h);
    return 1.f / (kPi * (a[0] * a[1]) * (d * d));
}
inline void visibilityAndG1Wo(float t_dot_wi, float t_dot_wo, float b_dot_wi, float b_dot_wo, float n_dot_wi,
                              float n_dot_wo, const float a[2], float out[2]) {  // :421-434
    const float t_wo = length3({{a[0] * t_dot_wo, a[1] * b_dot_wo, n_dot_wo, 0.f}});
    const float t_wi = length3({{a[0] * t_dot_wi, a[1] * b_dot_wi, n_dot_wi, 0.f}});
    out[0]           = 0.5f / (n_dot_wi * t_wo + n_dot_wo * t_wi);
    out[1]           = t_wo + n_dot_wo;
}

// Aniso.reflectionF, :268-305
inline bxdf::Result reflection(Vec4f wi, Vec4f wo, Vec4f h, float n_dot_wi, float n_dot_wo, float wo_dot_h,
                               const float alpha[2], const fresnel::Schlick& fr, const Frame& frame) {
    if (alpha[0] == alpha[1]) return iso::reflection(h, frame.z, n_dot_wi, n_dot_wo, wo_dot_h, alpha[0], fr);

    const float n_dot_h = saturate(dot3(frame.z, h));
    const float x_dot_h = dot3(frame.x, h);
    const float y_dot_h = dot3(frame.y, h);
    const float d       = distribution(n_dot_h, x_dot_h, y_dot_h, alpha);

    float g[2];
    visibilityAndG1Wo(dot3(frame.x, wi), dot3(frame.x, wo), dot3(frame.y, wi), dot3(frame.y, wo), n_dot_wi, n_dot_wo,
                      alpha, g);
    const Vec4f f = fr.f(wo_dot_h);
    return {splat(d * g[0]) * f, pdfVisible(d, g[1])};
}

// Aniso.reflect, :307-353
inline Micro reflect(Vec4f wo, float n_dot_wo, const float alpha[2], float specular_threshold, const float xi[2],
                     const fresnel::Schlick& fr, const Frame& frame, bxdf::Sample& result) {
    if (alpha[0] == alpha[1]) return iso::reflect(wo, n_dot_wo, alpha[0], specular_threshold, xi, fr, frame, result);

    float       n_dot_h;
    const Vec4f h = sampleVndf(wo, alpha, xi, frame, n_dot_h);

    const float x_dot_h  = dot3(frame.x, h);
    const float y_dot_h  = dot3(frame.y, h);
    const float wo_dot_h = safe::clampDot(wo, h);
    const Vec4f wi       = normalize3(mulAdd(splat(2.f * wo_dot_h), h, -wo));
    const float n_dot_wi = frame.clampNdot(wi);

    const float d = distribution(n_dot_h, x_dot_h, y_dot_h, alpha);
    float       g[2];
    visibilityAndG1Wo(dot3(frame.x, wi), dot3(frame.x, wo), dot3(frame.y, wi), dot3(frame.y, wo), n_dot_wi, n_dot_wo,
                      alpha, g);
    const Vec4f f = fr.f(wo_dot_h);

    result.reflection = splat(d * g[0]) * f;
    result.wi         = wi;
    result.pdf        = pdfVisible(d, g[1]);
    result.path       = bxdf::Path::reflection(alpha[1], specular_threshold);
    return {h, n_dot_wi, wo_dot_h};
}
}  // namespace aniso

}  // namespace ggx

namespace diffuse {  // diffuse.zig:42-116
inline float estimateContribution(const GgxLuts& luts, float /*n_dot_wo*/, float alpha, float f0, float albedo) {
    const float e_avg = luts.eAvg(alpha, f0);
    const float a     = e_avg;
    const float b     = 1.f / (kPi * (1.f - e_avg)) * albedo;
    return b / (a + b);
}
inline Vec4f evaluate(const GgxLuts& luts, Vec4f color, float n_dot_wi, float n_dot_wo, float alpha, float f0) {
    const float e_wo  = luts.e(n_dot_wo, alpha, f0);
    const float e_wi  = luts.e(n_dot_wi, alpha, f0);
    const float e_avg = luts.eAvg(alpha, f0);
    return splat(((1.f - e_wo) * (1.f - e_wi)) / (kPi * (1.f - e_avg))) * color;
}
inline bxdf::Result reflection(const GgxLuts& luts, Vec4f color, float f0, float n_dot_wi, float n_dot_wo, float alpha) {
    return {evaluate(luts, color, n_dot_wi, n_dot_wo, alpha, f0), n_dot_wi * kPiInv};
}
inline ggx::Micro reflect(const GgxLuts& luts, Vec4f color, float f0, Vec4f wo, float n_dot_wo, const Frame& frame,
                          float alpha, const float xi[2], bxdf::Sample& result) {
    const Vec4f is = hemisphereCosine(xi);
    const Vec4f wi = normalize3(frame.frameToWorld(is));
    const Vec4f h  = normalize3(wo + wi);

    const float h_dot_wi = safe::clampDot(h, wi);
    const float n_dot_wi = frame.clampNdot(wi);

    result.reflection = evaluate(luts, color, n_dot_wi, n_dot_wo, alpha, f0);
    result.wi         = wi;
    result.pdf        = n_dot_wi * kPiInv;
    result.path       = bxdf::Path::diffuseReflection();
    return {h, n_dot_wi, h_dot_wi};
}
}  // namespace diffuse

// Material sample of the Material union restricted to {Substitute surface, Light}: material_sample.zig.
// Coating, substitute/substitute_coating.zig (the clear coat of a Substitute)
struct Coating {
    Vec4f n, absorption_coef;
    float thickness = 0.f, f0 = 0.f, alpha = 0.f, weight = 0.f;

    struct Result {
        Vec4f reflection, attenuation;
        float f, pdf;
    };

    static Vec4f attenuation3(Vec4f c, float distance) {  // collision_coefficients.zig:64-66
        const Vec4f x = splat(-distance) * c;
        return {{std::exp(x[0]), std::exp(x[1]), std::exp(x[2]), std::exp(x[3])}};
    }
    Vec4f singleAttenuation(float n_dot_wo) const {  // :92-100
        const float d = thickness * (1.f / n_dot_wo);
        return attenuation3(absorption_coef, d);
    }
    Vec4f attenuation(float n_dot_wi, float n_dot_wo) const {  // :102-108
        const float f = weight * fresnel::schlick1(min(n_dot_wi, n_dot_wo), f0);
        const float d = thickness * (1.f / n_dot_wi + 1.f / n_dot_wo);
        return splat(1.f - f) * attenuation3(absorption_coef, d);
    }
    // :35-58
    Result evaluate(const GgxLuts& luts, Vec4f wi, Vec4f wo, Vec4f h, float wo_dot_h, float specular_threshold, bool avoid_caustics) const {
        const float n_dot_wi = safe::clampDot(n, wi);
        const float n_dot_wo = safe::clampAbsDot(n, wo);
        const Vec4f att      = attenuation(n_dot_wi, n_dot_wo);
        if (avoid_caustics && alpha <= specular_threshold) return {splat(0.f), att, 0.f, 0.f};

        const ggx::iso::ResultF gg = ggx::iso::reflectionF(h, n, n_dot_wi, n_dot_wo, wo_dot_h, alpha, f0);
        const float             ep = ggx::ilmEpDielectric(luts, n_dot_wo, alpha, f0);
        return {splat(ep * weight * n_dot_wi) * gg.r.reflection, att, gg.f, gg.r.pdf};
    }
    // :60-82
    Vec4f reflect(const GgxLuts& luts, Vec4f wo, Vec4f h, float n_dot_wo, float n_dot_h, float wo_dot_h, float specular_threshold,
                  bxdf::Sample& result) const {
        Vec4f t, b;
        orthonormalBasis3(n, t, b);
        const float n_dot_wi = ggx::iso::reflectNoFresnel(wo, h, n_dot_wo, n_dot_h, wo_dot_h, alpha, specular_threshold, Frame{t, b, n}, result);
        const float ep       = ggx::ilmEpDielectric(luts, n_dot_wo, alpha, f0);
        result.reflection    = result.reflection * splat(ep * weight * n_dot_wi);
        return attenuation(n_dot_wi, n_dot_wo);
    }
    // :84-90: Micro.n_dot_wi carries the Fresnel term
    ggx::Micro sample(Vec4f wo, const float xi[2], float& n_dot_h) const {
        Vec4f t, b;
        orthonormalBasis3(n, t, b);
        const float a[2]     = {alpha, alpha};
        const Vec4f h        = ggx::sampleVndf(wo, a, xi, Frame{t, b, n}, n_dot_h);
        const float wo_dot_h = safe::clampDot(wo, h);
        return {h, fresnel::schlick1(wo_dot_h, f0), wo_dot_h};
    }
};

struct MaterialSample {
    enum Kind { Light, Substitute, Glass } kind;

    SampleBase super;

    // Substitute, substitute_sample.zig:20-36
    Vec4f albedo, f0;
    float metallic, specular, specular_threshold, opacity;
    Coating coating;

    // Glass, glass_sample.zig:20-30 (abbe == 0, thickness == 0)
    Vec4f absorption_coef;
    float ior, ior_outside, glass_f0;

    const GgxLuts* luts;

    bool canEvaluate() const { return super.can_evaluate; }
    bool isTranslucent() const { return super.translucent; }

    // substitute_sample.zig:236-278
    bxdf::Result baseEvaluate(Vec4f wi, Vec4f wo, Vec4f h, float wo_dot_h, bool force_disable_caustics) const {
        const Frame& frame = super.frame;
        const float* alpha = super.alpha;

        const float n_dot_wi = frame.clampNdot(wi);
        const float n_dot_wo = frame.clampAbsNdot(wo);

        bxdf::Result d  = bxdf::Result::empty();
        float        dw = 0.f;

        if (1.f != metallic) {
            const Vec4f a   = splat(opacity) * albedo;
            const float f0m = hmax3(f0);
            d               = diffuse::reflection(*luts, a, f0m, n_dot_wi, n_dot_wo, alpha[1]);
            const float am  = hmax3(albedo);
            dw              = diffuse::estimateContribution(*luts, n_dot_wo, alpha[1], f0m, am);
        }

        if (super.avoidCausticsForce(force_disable_caustics) && alpha[1] <= specular_threshold) {
            return {splat(n_dot_wi) * d.reflection, dw * d.pdf};
        }

        const float            s = specular;
        const fresnel::Schlick schlick{f0};

        const bxdf::Result gg  = ggx::aniso::reflection(wi, wo, h, n_dot_wi, n_dot_wo, wo_dot_h, alpha, schlick, frame);
        const Vec4f        mms = ggx::dspbrMicroEc(*luts, f0, n_dot_wi, n_dot_wo, alpha[1]);

        const float pdf = dw * d.pdf + (1.f - dw) * gg.pdf;
        return {splat(n_dot_wi) * (d.reflection + splat(s) * (gg.reflection + mms)), pdf};
    }

    // material_sample.zig:56-62 -> substitute_sample.zig:88-145 (surface, opaque, uncoated)
    bxdf::Result evaluate(Vec4f wi, uint32_t max_splits, bool force_disable_caustics) const {
        if (Light == kind) return {splat(0.f), 0.f};
        if (Glass == kind) return glassEvaluate(wi, max_splits, force_disable_caustics);

        const Vec4f wo = super.wo;
        if (!super.sameHemisphere(wo)) return bxdf::Result::empty();

        const Vec4f h        = normalize3(wo + wi);
        const float wo_dot_h = safe::clampDot(wo, h);
        const bxdf::Result base_result = baseEvaluate(wi, wo, h, wo_dot_h, force_disable_caustics);
        if (coating.thickness > 0.f) {  // :138-142
            const Coating::Result c = coating.evaluate(*luts, wi, wo, h, wo_dot_h, specular_threshold, super.avoidCausticsForce(force_disable_caustics));
            const float           pdf = c.f * c.pdf + (1.f - c.f) * base_result.pdf;
            return {c.reflection + c.attenuation * base_result.reflection, pdf};
        }
        return base_result;
    }

    // substitute_sample.zig:338-361
    ggx::Micro diffuseSample(float diffuse_weight, const float xi[2], bxdf::Sample& result) const {
        const Vec4f  wo    = super.wo;
        const Frame& frame = super.frame;
        const float* alpha = super.alpha;

        const float n_dot_wo = frame.clampAbsNdot(wo);

        const Vec4f      a     = splat(opacity) * albedo;
        const float      f0m   = hmax3(f0);
        const ggx::Micro micro = diffuse::reflect(*luts, a, f0m, wo, n_dot_wo, frame, alpha[1], xi, result);

        const fresnel::Schlick schlick{f0};
        const bxdf::Result     gg =
            ggx::aniso::reflection(result.wi, wo, micro.h, micro.n_dot_wi, n_dot_wo, micro.h_dot_wi, alpha, schlick, frame);
        const Vec4f mms = ggx::dspbrMicroEc(*luts, f0, micro.n_dot_wi, frame.clampNdot(wo), alpha[1]);

        const float s     = specular;
        result.reflection = splat(micro.n_dot_wi) * (result.reflection + splat(s) * (gg.reflection + mms));
        result.pdf        = diffuse_weight * result.pdf + (1.f - diffuse_weight) * gg.pdf;
        return micro;
    }

    // substitute_sample.zig:363-410 (no flakes)
    ggx::Micro glossSample(float diffuse_weight, const float xi[2], bxdf::Sample& result) const {
        const Vec4f  wo    = super.wo;
        const Frame& frame = super.frame;
        const float* alpha = super.alpha;
        const float  s     = specular;

        const float            n_dot_wo = frame.clampAbsNdot(wo);
        const fresnel::Schlick schlick{f0};

        const ggx::Micro micro = ggx::aniso::reflect(wo, n_dot_wo, alpha, specular_threshold, xi, schlick, frame, result);
        const Vec4f      mms   = ggx::dspbrMicroEc(*luts, f0, micro.n_dot_wi, frame.clampNdot(wo), alpha[1]);

        bxdf::Result d = bxdf::Result::empty();
        if (diffuse_weight > 0.f) {
            const Vec4f a   = splat(opacity) * albedo;
            const float f0m = hmax3(f0);
            d               = diffuse::reflection(*luts, a, f0m, micro.n_dot_wi, n_dot_wo, alpha[1]);
        }

        result.reflection = splat(micro.n_dot_wi) * (splat(s) * (result.reflection + mms) + d.reflection);
        result.pdf        = (1.f - diffuse_weight) * result.pdf + diffuse_weight * d.pdf;
        return micro;
    }

    // substitute_sample.zig:304-336, 412-433
    void coatingSample(Sampler& sampler, bxdf::Sample& result) const {
        float       n_dot_h;
        const Vec2f s2    = sampler.sample2D();
        const float xi2[2] = {s2[0], s2[1]};
        const ggx::Micro micro = coating.sample(super.wo, xi2, n_dot_h);
        const float      f     = micro.n_dot_wi;

        const Vec4f s3 = sampler.sample3D();
        const float p  = s3[0];
        if (p <= f) {
            // coatingReflect
            const Vec4f wo       = super.wo;
            const float n_dot_wo = safe::clampAbsDot(coating.n, wo);
            const Vec4f coating_attenuation = coating.reflect(*luts, wo, micro.h, n_dot_wo, n_dot_h, micro.h_dot_wi, specular_threshold, result);
            const bxdf::Result base_result  = baseEvaluate(result.wi, wo, micro.h, micro.h_dot_wi, false);
            result.reflection = (result.reflection * splat(f)) + coating_attenuation * base_result.reflection;
            result.pdf        = f * result.pdf + (1.f - f) * base_result.pdf;
        } else {
            float dw = 0.f;
            if (1.f != metallic) {
                const float n_dot_wo = super.frame.clampAbsNdot(super.wo);
                const float f0m      = hmax3(f0);
                const float am       = hmax3(albedo);
                dw                   = diffuse::estimateContribution(*luts, n_dot_wo, super.alpha[1], f0m, am);
            }
            const float xi[2] = {s3[1], s3[2]};
            const float p1    = (p - f) / (1.f - f);
            // coatingBaseSample
            const ggx::Micro base_micro = p1 < dw ? diffuseSample(dw, xi, result) : glossSample(dw, xi, result);
            const Coating::Result c =
                coating.evaluate(*luts, result.wi, super.wo, base_micro.h, base_micro.h_dot_wi, specular_threshold, super.avoid_caustics);
            result.reflection = c.attenuation * result.reflection + c.reflection;
            result.pdf        = (1.f - f) * result.pdf + f * c.pdf;
        }
    }

    // substitute_sample.zig:280-302
    void baseSample(Sampler& sampler, bxdf::Sample& result) const {
        float dw = 0.f;
        if (1.f != metallic) {
            const float n_dot_wo = super.frame.clampAbsNdot(super.wo);
            const float f0m      = hmax3(f0);
            const float am       = hmax3(albedo);
            dw                   = diffuse::estimateContribution(*luts, n_dot_wo, super.alpha[1], f0m, am);
        }

        const Vec4f s3    = sampler.sample3D();
        const float p     = s3[0];
        const float xi[2] = {s3[1], s3[2]};
        if (p < dw) {
            diffuseSample(dw, xi, result);
        } else {
            glossSample(dw, xi, result);
        }
    }

    // material_sample.zig:64-78 -> substitute_sample.zig:147-234 (surface, opaque, uncoated). Returns the count.
    uint32_t sample(Sampler& sampler, uint32_t max_splits, bxdf::Sample buffer[4]) const {
        if (Light == kind) return 0;
        if (Glass == kind) return glassSample(sampler, max_splits, buffer);

        if (!super.sameHemisphere(super.wo)) return 0;

        bxdf::Sample& result = buffer[0];
        result.split_weight  = 1.f;
        result.wavelength    = 0.f;

        if (coating.thickness > 0.f) {
            coatingSample(sampler, result);
        } else {
            baseSample(sampler, result);
        }

        if (0.f == result.pdf) return 0;
        return 1;
    }

    // ---- Glass, glass_sample.zig ----

    // Sample.evaluate, :68-152
    bxdf::Result glassEvaluate(Vec4f wi, uint32_t max_splits, bool force_disable_caustics) const {
        const float alpha = super.alpha[0];
        const bool  rough = alpha > 0.f;

        if (ior == ior_outside || !rough || super.lower_priority ||
            (super.avoidCausticsForce(force_disable_caustics) && alpha <= specular_threshold)) {
            return bxdf::Result::empty();
        }

        const Frame& frame = super.frame;
        const bool   split = max_splits > 1;
        const float  s     = specular;
        const Vec4f  wo    = super.wo;

        if (!super.sameHemisphere(wo)) {
            const IoR   io{ior_outside, ior};  // eta_i = self.ior, eta_t = self.ior_outside
            const Vec4f h = -normalize3(splat(io.eta_t) * wi + splat(io.eta_i) * wo);

            const float wi_dot_h = dot3(wi, h);
            if (wi_dot_h <= 0.f) return bxdf::Result::empty();

            const float wo_dot_h = dot3(wo, h);
            const float eta      = io.eta_i / io.eta_t;
            const float sint2    = (eta * eta) * (1.f - wo_dot_h * wo_dot_h);
            if (sint2 >= 1.f) return bxdf::Result::empty();

            const float n_dot_wi = frame.clampNdot(wi);
            const float n_dot_wo = frame.clampAbsNdot(wo);
            const float n_dot_h  = saturate(frame.nDot(h));

            const ggx::iso::ResultF gg   = ggx::iso::refractionF(n_dot_wi, n_dot_wo, wi_dot_h, wo_dot_h, n_dot_h, alpha, io, glass_f0);
            const float             comp = ggx::ilmEpDielectric(*luts, n_dot_wo, alpha, glass_f0);

            const float split_pdf = split ? 1.f : gg.f;
            return {splat(min(n_dot_wi, n_dot_wo) * comp * s) * gg.r.reflection, split_pdf * gg.r.pdf};
        } else if (super.sameHemisphere(wi)) {
            const float n_dot_wi = frame.clampNdot(wi);
            const float n_dot_wo = frame.clampAbsNdot(wo);

            const Vec4f h        = normalize3(wo + wi);
            const float wo_dot_h = safe::clampDot(wo, h);

            const ggx::iso::ResultF gg   = ggx::iso::reflectionF(h, frame.z, n_dot_wi, n_dot_wo, wo_dot_h, alpha, glass_f0);
            const float             comp = ggx::ilmEpDielectric(*luts, n_dot_wo, alpha, glass_f0);

            const float split_pdf = split ? 1.f : gg.f;
            return {splat(n_dot_wi * comp * s) * gg.r.reflection, split_pdf * gg.r.pdf};
        }
        return bxdf::Result::empty();
    }

    // Sample.sample, :167-200 (thickness == 0, abbe == 0)
    uint32_t glassSample(Sampler& sampler, uint32_t max_splits, bxdf::Sample buffer[4]) const {
        const bool split = max_splits > 1;
        if (super.alpha[0] > 0.f) return glassRoughSample(sampler, split, buffer);
        return glassSpecularSample(sampler, split, buffer);
    }

    static bxdf::Sample glassReflect(Vec4f wo, Vec4f n, float n_dot_wo, float split_weight, float specular) {  // :429-438
        return {splat(specular) * splat(1.f), normalize3(splat(2.f * n_dot_wo) * n - wo), 1.f, split_weight, 0.f,
                bxdf::Path::singularReflection()};
    }
    static bxdf::Sample thickSpecularRefract(Vec4f wo, Vec4f n, float n_dot_wo, float n_dot_t, float eta, float split_weight) {  // :519-537
        return {splat(1.f), normalize3(splat(eta * n_dot_wo - n_dot_t) * n - splat(eta) * wo), 1.f, split_weight, 0.f,
                bxdf::Path::singularTransmission()};
    }

    // specularSample, :202-284 (Thin = false, weight = 1)
    uint32_t glassSpecularSample(Sampler& sampler, bool split, bxdf::Sample buffer[4]) const {
        float eta_i = ior_outside;
        float eta_t = ior;

        const Vec4f wo = super.wo;

        if (eta_i == eta_t || super.lower_priority) {
            buffer[0] = {splat(1.f), -wo, 1.f, 1.f, 0.f, bxdf::Path::singularTransmission()};
            return 1;
        }

        Vec4f n = super.frame.z;
        if (!super.sameHemisphere(wo)) {
            n = -n;
            std::swap(eta_i, eta_t);
        }

        const float n_dot_wo = min(std::fabs(dot3(n, wo)), 1.f);
        const float eta      = eta_i / eta_t;
        const float sint2    = (eta * eta) * (1.f - n_dot_wo * n_dot_wo);

        const float s = specular;

        float n_dot_t, f;
        if (sint2 >= 1.f) {
            n_dot_t = 0.f;
            f       = 1.f;
        } else {
            n_dot_t = std::sqrt(1.f - sint2);
            f       = fresnel::dielectric(n_dot_wo, n_dot_t, eta_i, eta_t);
        }

        if (split) {
            buffer[0] = glassReflect(wo, n, n_dot_wo, f, s);
            if (1.f == f) return 1;
            buffer[1] = thickSpecularRefract(wo, n, n_dot_wo, n_dot_t, eta, 1.f - f);
            return 2;
        }
        const float p = sampler.sample1D();
        if (p <= f) {
            buffer[0] = glassReflect(wo, n, n_dot_wo, 1.f, s);
        } else {
            buffer[0] = thickSpecularRefract(wo, n, n_dot_wo, n_dot_t, eta, 1.f);
        }
        return 1;
    }

    // roughRefract, :440-503 (Thin = false)
    float roughRefract(bool same_side, const Frame& frame, Vec4f wo, Vec4f h, float n_dot_wo, float n_dot_h, float wi_dot_h,
                       float wo_dot_h, float alpha, IoR io, bxdf::Sample& result) const {
        const float r_wo_dot_h = same_side ? -wo_dot_h : wo_dot_h;
        return ggx::iso::refractNoFresnel(wo, h, n_dot_wo, n_dot_h, -wi_dot_h, r_wo_dot_h, alpha, specular_threshold, io, frame, result);
    }

    // roughSample, :286-427 (Thin = false, weight = 1)
    uint32_t glassRoughSample(Sampler& sampler, bool split, bxdf::Sample buffer[4]) const {
        const IoR quo_ior{ior, ior_outside};  // eta_i = ior_outside, eta_t = ior_t

        const Vec4f wo = super.wo;

        if (std::fabs(quo_ior.eta_i - quo_ior.eta_t) <= 2.e-7f || super.lower_priority) {
            buffer[0] = {splat(1.f), -wo, 1.f, 1.f, 0.f, bxdf::Path::singularTransmission()};
            return 1;
        }

        const float alpha = super.alpha[0];

        const bool same_side = super.sameHemisphere(wo);

        const Frame frame = super.frame.swapped(same_side);
        const IoR   io    = quo_ior.swapped(same_side);

        const Vec4f s3    = sampler.sample3D();
        const float xi[2] = {s3[1], s3[2]};

        float       n_dot_h;
        const Vec4f h = ggx::sampleVndf(wo, super.alpha, xi, frame, n_dot_h);

        const float n_dot_wo = frame.clampAbsNdot(wo);
        const float wo_dot_h = safe::clampDot(wo, h);

        const float eta   = io.eta_i / io.eta_t;
        const float sint2 = (eta * eta) * (1.f - wo_dot_h * wo_dot_h);

        const float s = specular;

        float wi_dot_h, f;
        if (sint2 >= 1.f) {
            wi_dot_h = 0.f;
            f        = 1.f;
        } else {
            wi_dot_h          = std::sqrt(1.f - sint2);
            const float cos_x = io.eta_i > io.eta_t ? wi_dot_h : wo_dot_h;
            f                 = fresnel::schlick1(cos_x, glass_f0);
        }

        if (split) {
            const float ep = ggx::ilmEpDielectric(*luts, n_dot_wo, alpha, glass_f0);
            {
                const float n_dot_wi = ggx::iso::reflectNoFresnel(wo, h, n_dot_wo, n_dot_h, wo_dot_h, alpha, specular_threshold, frame, buffer[0]);
                buffer[0].reflection   = buffer[0].reflection * (splat(n_dot_wi * ep * s) * splat(1.f));
                buffer[0].split_weight = f;
                buffer[0].wavelength   = 0.f;
            }
            if (1.f == f) return 1;
            {
                const float n_dot_wi = roughRefract(same_side, frame, wo, h, n_dot_wo, n_dot_h, wi_dot_h, wo_dot_h, alpha, io, buffer[1]);
                if (n_dot_wi < 0.f) return 1;
                buffer[1].reflection   = buffer[1].reflection * (splat(n_dot_wi * ep) * splat(1.f));
                buffer[1].split_weight = 1.f - f;
                buffer[1].wavelength   = 0.f;
            }
            return 2;
        }

        bxdf::Sample& result = buffer[0];

        const float ep = ggx::ilmEpDielectric(*luts, n_dot_wo, alpha, glass_f0);

        const float p = s3[0];
        if (p <= f) {
            const float n_dot_wi = ggx::iso::reflectNoFresnel(wo, h, n_dot_wo, n_dot_h, wo_dot_h, alpha, specular_threshold, frame, result);
            result.reflection    = result.reflection * (splat(f * n_dot_wi * ep * s) * splat(1.f));
            result.pdf *= f;
        } else {
            const float n_dot_wi = roughRefract(same_side, frame, wo, h, n_dot_wo, n_dot_h, wi_dot_h, wo_dot_h, alpha, io, result);
            if (n_dot_wi < 0.f) return 0;
            const float omf   = 1.f - f;
            result.reflection = result.reflection * (splat(omf * n_dot_wi * ep) * splat(1.f));
            result.pdf *= omf;
        }
        result.split_weight = 1.f;
        result.wavelength   = 0.f;
        return 1;
    }
};

// Emittance.radiance, emittance.zig:29-59 (uniform emission map, no profile)
inline Vec4f emittanceRadiance(const ZygpuMaterial& m, Vec4f wi, const Trafo& trafo, float area, bool in_camera) {
    if (-dot3(wi, trafo.r[2]) < m.emission_cos_a) return splat(0.f);

    const float factor    = in_camera ? m.emission_camera_weight : 1.f;
    const Vec4f intensity = load4(m.emission) * Vec4f{{1.f, 1.f, 1.f, 0.f}};  // value * uniform1(1.0) map

    if (0.f != m.emission_normalize) return splat(factor / area) * intensity;
    return splat(factor) * intensity;
}

// Substitute Material.sample, substitute_material.zig:114-221, uniform textures, no coating / flakes / normal map.
inline MaterialSample substituteSample(const ZygpuMaterial& m, Vec4f wo, const Renderstate& rs, float specular_threshold,
                                       const GgxLuts& luts) {
    const Vec4f color     = load4(m.color);
    const float roughness = ggx::clampRoughness(m.roughness);
    const float metallic  = m.metallic;
    const float specular  = m.specular;

    float alpha[2];  // anisotropicAlpha, :299-306
    if (m.anisotropy > 0.f) {
        const float rv = ggx::clampRoughness(roughness * (1.f - m.anisotropy));
        alpha[0]       = roughness * roughness;
        alpha[1]       = rv * rv;
    } else {
        alpha[0] = alpha[1] = roughness * roughness;
    }

    // coating_scale is the uniform 1: weight 1 (substitute_material.zig:128-134)
    const float coating_thickness = 1.f * m.coating_thickness;
    const float coating_weight    = 1.f;
    const float coating_ior       = lerp(rs.ior, m.coating_ior, coating_weight);

    const float ior       = m.ior;
    const float ior_outer = coating_thickness > 0.f ? coating_ior : rs.ior;

    // Surface.init, substitute_sample.zig:38-78
    MaterialSample r;
    r.kind = MaterialSample::Substitute;
    r.luts = &luts;

    const Vec4f c = splat(1.f - metallic) * color;
    float       reg_alpha[2];
    rs.regularizeAlpha(alpha, specular_threshold, reg_alpha);
    const float ior_medium = rs.ior;

    r.super.geo_n          = rs.geo_n;
    r.super.n              = rs.n;
    r.super.wo             = wo;
    r.super.alpha[0]       = reg_alpha[0];
    r.super.alpha[1]       = reg_alpha[1];
    r.super.can_evaluate   = ior != ior_medium;
    r.super.avoid_caustics = !rs.caustics;
    r.super.translucent    = false;
    r.super.volumetric     = false;

    const float f0 = fresnel::Schlick::IorToF0(ior, ior_outer);

    r.albedo             = c;
    r.f0                 = lerp(splat(f0), color, splat(metallic));
    r.metallic           = metallic;
    r.specular           = specular;
    r.specular_threshold = specular_threshold;
    r.opacity            = 1.f;  // 1 - 0.5 * translucency

    r.super.frame = {rs.t, rs.b, rs.n};

    if (coating_thickness > 0.f) {  // :164-181. No coating normal map: n is the base frame's normal without a normal map (the two uniform
        // textures are equal) and the interpolated normal with one (not equal, coating map uniform) - rs.n either way
        r.coating.n               = rs.n;
        const float cr            = ggx::clampRoughness(m.coating_roughness);
        r.coating.absorption_coef = {{m.coating_absorption[0], m.coating_absorption[1], m.coating_absorption[2], 0.f}};
        r.coating.thickness       = coating_thickness;
        r.coating.f0              = fresnel::Schlick::IorToF0(coating_ior, rs.ior);
        const float ca[2]         = {cr * cr, cr * cr};
        float       reg[2];
        rs.regularizeAlpha(ca, specular_threshold, reg);
        r.coating.alpha  = reg[0];
        r.coating.weight = coating_weight;
    }
    return r;
}

// Glass Material.sample, glass_material.zig:46-73 + Sample.init, glass_sample.zig:32-66 (uniform roughness, no normal map)
inline MaterialSample glassSample(const ZygpuMaterial& m, Vec4f wo, const Renderstate& rs, float specular_threshold, const GgxLuts& luts) {
    const float raw_r = m.roughness;
    const float r     = 0.f == raw_r ? 0.f : ggx::clampRoughness(raw_r);

    const float alpha[2] = {r * r, r * r};
    float       reg_alpha[2];
    rs.regularizeAlpha(alpha, specular_threshold, reg_alpha);
    const bool  rough       = reg_alpha[0] > 0.f;
    const float ior_outside = rs.ior;

    MaterialSample s;
    s.kind = MaterialSample::Glass;
    s.luts = &luts;

    s.super.geo_n          = rs.geo_n;
    s.super.n              = rs.n;
    s.super.wo             = wo;
    s.super.alpha[0]       = reg_alpha[0];
    s.super.alpha[1]       = reg_alpha[1];
    s.super.can_evaluate   = rough && m.ior != ior_outside;
    s.super.avoid_caustics = !rs.caustics;
    s.super.lower_priority = int8_t(m.priority) < rs.highest_priority;
    s.super.translucent    = m.thickness > 0.f;
    s.super.frame          = {rs.t, rs.b, rs.n};

    s.absorption_coef    = load4(m.color);
    s.ior                = m.ior;
    s.ior_outside        = ior_outside;
    s.glass_f0           = rough ? fresnel::Schlick::IorToF0(m.ior, ior_outside) : 0.f;
    s.specular           = m.specular;
    s.specular_threshold = specular_threshold;
    return s;
}

// light_material.zig:108-110 -> light_sample.zig:10-12 (Base.initTBN, can_evaluate = false)
inline MaterialSample lightSample(Vec4f wo, const Renderstate& rs) {
    MaterialSample r;
    r.kind                 = MaterialSample::Light;
    r.luts                 = nullptr;
    r.super.frame          = {rs.t, rs.b, rs.n};
    r.super.geo_n          = rs.geo_n;
    r.super.n              = rs.n;
    r.super.wo             = wo;
    r.super.alpha[0]       = 0.f;
    r.super.alpha[1]       = 0.f;
    r.super.can_evaluate   = false;
    r.super.avoid_caustics = !rs.caustics;
    return r;
}

}  // namespace zo
