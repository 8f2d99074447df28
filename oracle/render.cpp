// ORACLE — test infrastructure only (see zmath.hpp). CPU restatement of the reference's forward
// surface-integration pass over the flattened scene of include/zygpu_scene.h:
//   Worker.render                        src/core/rendering/worker.zig:104-168
//   Sensor.cameraSample / addSample      src/core/rendering/sensor/sensor.zig:152-385, 559-628
//   Perspective.generateVertex           src/core/camera/camera_perspective.zig:124-150
//   PathtracerMIS.li & friends           src/core/rendering/integrator/surface/pathtracer_mis.zig:37-341
//   helpers                              src/core/rendering/integrator/helper.zig
//   Vertex / Pool                        src/core/scene/vertex.zig
//   Context / Scene queries              src/core/scene/context.zig:54-73, scene.zig:225-252, 592-634
//   PropBvh                              src/core/scene/prop/prop_tree.zig:56-116, 185-240, 302-356
//   Prop                                 src/core/scene/prop/prop.zig:163-264
//   Rectangle / Cube / Sphere            src/core/scene/shape/{rectangle,cube,sphere}.zig
//   Light / light tree                   src/core/scene/light/light.zig, light_tree.zig:65-227, 346-517
// PARITY UNPINNED: the reference has no tests or fixtures for this path and cannot be built here.
// Scope of this restatement: static scenes, no volumes / media, opaque film, no AOVs, no shadow catchers.
#include "zmaterial.hpp"
#include "ztree.hpp"
#include "zyg_oracle.h"

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>

namespace zo {

namespace {

struct Intersection {  // shape/intersection.zig:46-61
    float    t, u, v;
    uint32_t primitive;
    Trafo    trafo;
};

struct Fragment {  // shape/intersection.zig:63-124 (event is always .Pass without volumes)
    Intersection isec;
    uint32_t     prop, part;
    Vec4f        p, geo_n, t, b, n, uvw;

    bool  hit() const { return ZYGPU_NULL != prop; }
    float offset() const { return uvw[3]; }
    bool  sameHemisphere(Vec4f v) const { return dot3(geo_n, v) > 0.f; }
    Vec4f offsetP(Vec4f v) const {  // :112-116
        const Vec4f nn = sameHemisphere(v) ? geo_n : -geo_n;
        return offsetRay(mulAdd(splat(offset()), nn, p), nn);
    }
    Ray offsetRayTo(Vec4f dir) const { return Ray::init(offsetP(dir), dir, 0.f, RayMaxT); }  // :118-120
};

struct Depth {  // shape/probe.zig:7-22
    uint16_t surface = 0, volume = 0;
    uint32_t total() const { return uint32_t(surface) + volume; }
};

struct State {  // vertex.zig:19-43
    bool primary_ray = true, transparent = true, singular = true, specular = false, translucent = false,
         started_specular = false;

    void update(const bxdf::Path& path) {
        if (bxdf::Scattering::Specular == path.scattering) {
            specular = true;
            singular = path.singular();
            if (primary_ray) started_specular = true;
        } else if (bxdf::Event::Straight != path.event) {
            specular    = false;
            singular    = false;
            primary_ray = false;
        }
    }
};

struct MediumStack {  // prop/medium.zig:30-153; the collision coefficients are the material's (Glass: absorption, no scattering)
    static constexpr uint32_t NumEntries = 4;
    struct Medium {
        uint32_t prop, part;
        float    ior;
        int8_t   priority;
        Vec4f    cc_a;
        bool     matches(uint32_t p, uint32_t pa) const { return prop == p && part == pa; }
    };
    uint32_t index = 0;
    Medium   m_stack[NumEntries];

    bool          empty() const { return 0 == index; }
    const Medium& top() const { return m_stack[index - 1]; }
    Vec4f         topCC() const {
        int8_t   priority = -128;
        uint32_t highest  = 0;
        for (uint32_t i = 0; i < index; ++i) {
            const int8_t lp = m_stack[i].priority;
            if (lp >= priority) {
                priority = lp;
                highest  = i;
            }
        }
        return m_stack[highest].cc_a;
    }
    int8_t highestPriority() const {
        int8_t priority = -128;
        for (uint32_t i = 0; i < index; ++i) priority = std::max(priority, m_stack[i].priority);
        return priority;
    }
    float topIor() const { return index > 0 ? m_stack[index - 1].ior : 1.f; }
    float peekIor(uint32_t prop, uint32_t part) const {
        if (index <= 1) return 1.f;
        const uint32_t back = index - 1;
        return m_stack[back].matches(prop, part) ? m_stack[back - 1].ior : m_stack[back].ior;
    }
    void push(uint32_t prop, uint32_t part, Vec4f cc_a, float ior, int8_t priority) {
        if (index < NumEntries - 1) {
            m_stack[index] = {prop, part, ior, priority, cc_a};
            index += 1;
        }
    }
    void remove(uint32_t prop, uint32_t part) {
        const int32_t back = int32_t(index) - 1;
        for (int32_t i = back; i >= 0; --i) {
            if (m_stack[i].matches(prop, part)) {
                for (int32_t j = i; j < back; ++j) m_stack[j] = m_stack[j + 1];
                index -= 1;
                return;
            }
        }
    }
};

// ray_offset.zig:29-31
inline float offsetF(float t) {
    const float origin      = 1.f / 32.f;
    const float float_scale = 1.f / 65536.f;
    const float int_scale   = 256.f;
    if (t < origin) return t + float_scale;
    int32_t i;
    std::memcpy(&i, &t, 4);
    i = int32_t(uint32_t(i) + uint32_t(int32_t(int_scale)));
    float r;
    std::memcpy(&r, &i, 4);
    return r;
}

struct Vertex {  // vertex.zig:45-85
    Ray   ray;
    Depth probe_depth;

    State    state;
    Depth    depth;
    float    bxdf_pdf               = 0.f;
    float    reg_alpha              = 0.f;
    float    split_weight           = 1.f;
    float    light_split_threshold  = 0.f;
    uint32_t path_count             = 1;
    Vec4f    throughput             = splat(1.f);
    Vec4f    origin;
    Vec4f    geo_n = splat(0.f);

    MediumStack mediums;
};

struct IValue {  // helper.zig:6-20
    Vec4f emission = splat(0.f), direct = splat(0.f), indirect = splat(0.f);
    void  add(Vec4f value, uint32_t depth, uint32_t direct_cutoff, bool is_emission, bool singular) {
        if (is_emission) {
            emission = emission + value;
        } else if (singular || depth < direct_cutoff) {
            direct = direct + value;
        } else {
            indirect = indirect + value;
        }
    }
};

constexpr float LowThreshold = 0.00000001f;  // helper.zig:29

inline float splitThreshold(float split_threshold, Depth depth) {  // helper.zig:33-39
    const uint32_t total_depth = depth.total();
    return min(total_depth < 4 ? split_threshold : LowThreshold, split_threshold);
}
inline float powerHeuristic(float f_pdf, float g_pdf) {  // helper.zig:64-67
    const float f2 = f_pdf * f_pdf;
    return f2 / std::fmaf(g_pdf, g_pdf, f2);
}
inline float predividedPowerHeuristic(float f_pdf, float g_pdf) {  // helper.zig:70-73
    const float f2 = f_pdf * f_pdf;
    return f_pdf / std::fmaf(g_pdf, g_pdf, f2);
}
inline bool russianRoulette(Vec4f& throughput, float r) {  // helper.zig:75-89
    const float mx                       = hmax3(throughput);
    const float continuation_probability = mx / 0.1f;
    if (continuation_probability < 1.f) {
        if (r >= continuation_probability) return true;
        throughput = throughput / splat(continuation_probability);
    }
    return false;
}

struct SampleTo {  // shape/sample.zig:10-32
    Vec4f p, n, wi, uvw;
    float pdf() const { return p[3]; }
};

struct LightPick {
    uint32_t offset;
    float    pdf;
};

// ---- shapes ----------------------------------------------------------------------------------

namespace rectangle {

bool intersect(const Ray& ray, const Trafo& trafo, Intersection& isec) {  // rectangle.zig:30-62
    const Vec4f n     = trafo.r[2];
    const float d     = dot3(n, trafo.position);
    const float hit_t = -(dot3(n, ray.origin) - d) / dot3(n, ray.direction);

    if (hit_t >= ray.min_t && ray.max_t >= hit_t) {
        const Vec4f p = ray.point(hit_t);
        const Vec4f k = p - trafo.position;
        const Vec4f t = -trafo.r[0];

        const float u = dot3(t, k) / (0.5f * trafo.scaleX());
        if (u > 1.f || u < -1.f) return false;

        const Vec4f b = -trafo.r[1];
        const float v = dot3(b, k) / (0.5f * trafo.scaleY());
        if (v > 1.f || v < -1.f) return false;

        isec.u         = u;
        isec.v         = v;
        isec.t         = hit_t;
        isec.primitive = 0;
        isec.trafo     = trafo;
        return true;
    }
    return false;
}

void fragment(const Ray& ray, Fragment& frag) {  // :102-124
    const Vec4f p = ray.point(frag.isec.t);
    const Vec4f n = frag.isec.trafo.r[2];
    const Vec4f t = -frag.isec.trafo.r[0];
    const Vec4f b = -frag.isec.trafo.r[1];

    frag.p     = p;
    frag.t     = t;
    frag.b     = b;
    frag.n     = n;
    frag.geo_n = n;
    if (frag.isec.trafo.scaleZ() < 0.f) {
        const Vec4f k = p - frag.isec.trafo.position;
        const float u = dot3(t, k) * 2.f;
        const float v = dot3(b, k) * 2.f;
        frag.uvw      = {{0.5f * (u + 1.f), 0.5f * (v + 1.f), 0.f, 0.f}};
    } else {
        frag.uvw = {{0.5f * (frag.isec.u + 1.f), 0.5f * (frag.isec.v + 1.f), 0.f, 0.f}};
    }
    frag.part = 0;
}

bool intersectP(const Ray& ray, const Trafo& trafo) {  // :126-152
    Intersection unused;
    return intersect(ray, trafo, unused);
}

// C. Ureña, M. Fajardo, A. King: An Area-Preserving Parametrization for Spherical Rectangles. rectangle.zig:199-303
struct SphQuad {
    Vec4f o, x, y, z;
    float z0, x0, y0, x1, y1, b0, b1, k, S;

    static SphQuad init(Vec4f scale, Vec4f o) {
        const Vec4f s  = {{-0.5f * scale[0], -0.5f * scale[1], 0.f, 0.f}};
        const Vec4f ex = {{scale[0], 0.f, 0.f, 0.f}};
        const Vec4f ey = {{0.f, scale[1], 0.f, 0.f}};

        SphQuad q;
        q.o             = o;
        const float exl = length3(ex);
        const float eyl = length3(ey);
        q.x             = ex / splat(exl);
        q.y             = ey / splat(eyl);
        q.z             = cross3(q.x, q.y);
        const Vec4f d   = s - o;
        q.z0            = dot3(d, q.z);
        if (q.z0 > 0.f) {
            q.z  = -q.z;
            q.z0 = -q.z0;
        }
        q.x0 = dot3(d, q.x);
        q.y0 = dot3(d, q.y);
        q.x1 = q.x0 + exl;
        q.y1 = q.y0 + eyl;

        const Vec4f v00 = {{q.x0, q.y0, q.z0, 0.f}};
        const Vec4f v01 = {{q.x0, q.y1, q.z0, 0.f}};
        const Vec4f v10 = {{q.x1, q.y0, q.z0, 0.f}};
        const Vec4f v11 = {{q.x1, q.y1, q.z0, 0.f}};

        const Vec4f n0 = normalize3(cross3(v00, v10));
        const Vec4f n1 = normalize3(cross3(v10, v11));
        const Vec4f n2 = normalize3(cross3(v11, v01));
        const Vec4f n3 = normalize3(cross3(v01, v00));

        const float g0 = std::acos(-dot3(n0, n1));
        const float g1 = std::acos(-dot3(n1, n2));
        const float g2 = std::acos(-dot3(n2, n3));
        const float g3 = std::acos(-dot3(n3, n0));

        q.b0 = n0[2];
        q.b1 = n2[2];
        q.k  = 2.f * kPi - g2 - g3;
        q.S  = g0 + g1 - q.k;
        return q;
    }

    Vec4f sample(const float uv[2]) const {
        const float au = uv[0] * S + k;
        const float fu = (std::cos(au) * b0 - b1) / std::sin(au);
        float       cu = 1.f / std::sqrt(fu * fu + b0 * b0) * (fu > 0.f ? 1.f : -1.f);
        cu             = cu < -1.f ? -1.f : (cu > 1.f ? 1.f : cu);  // std.math.clamp

        float xu = -(cu * z0) / std::sqrt(1.f - cu * cu);
        xu       = xu < x0 ? x0 : (xu > x1 ? x1 : xu);

        const float d   = std::sqrt(xu * xu + z0 * z0);
        const float h0  = y0 / std::sqrt(d * d + y0 * y0);
        const float h1  = y1 / std::sqrt(d * d + y1 * y1);
        const float hv  = h0 + uv[1] * (h1 - h0);
        const float hv2 = hv * hv;
        uint32_t    eb  = 0x35800000u;
        float       eps;
        std::memcpy(&eps, &eb, 4);
        const float yv = hv2 < 1.f - eps ? ((hv * d) / std::sqrt(1.f - hv2)) : y1;

        return o + splat(xu) * x + splat(yv) * y + splat(z0) * z;
    }

    float pdf(Vec4f scale) const {
        const Vec4f lp                      = o;
        const float sqr_dist                = squaredLength3(lp);
        const float area                    = scale[0] * scale[1];
        const float diff_solid_angle_numer  = area * std::fabs(lp[2]);
        const float diff_solid_angle_denom  = sqr_dist * std::sqrt(sqr_dist);
        return diff_solid_angle_numer > diff_solid_angle_denom * safe::DotMin ? (1.f / S)
                                                                               : (diff_solid_angle_denom / diff_solid_angle_numer);
    }
};

// std.math.clamp asserts lower <= upper; with a degenerate quad the reference would trap in debug builds
// and is unspecified in release builds. The scenes used here never reach that case.

// Rectangle.sampleTo, rectangle.zig:305-397 (UseSphericalSampling = true)
uint32_t sampleTo(Vec4f p, Vec4f n, const Trafo& trafo, bool two_sided, bool total_sphere, uint32_t num_samples,
                  Sampler& sampler, SampleTo* buffer) {
    const float nsf   = float(num_samples);
    const Vec4f scale = trafo.scale();

    const Vec4f   lp    = trafo.worldToFramePoint(p);
    const SphQuad squad = SphQuad::init(scale, lp);

    const float sample_pdf = nsf * squad.pdf(scale);

    uint32_t current_sample = 0;
    for (uint32_t i = 0; i < num_samples; ++i) {
        const Vec2f uv = sampler.sample2D();

        const Vec4f ls  = squad.sample(uv.v);
        const Vec4f ws  = trafo.frameToWorldPoint(ls);
        const Vec4f dir = normalize3(ws - p);

        Vec4f wn = trafo.r[2];
        if (two_sided && dot3(wn, dir) > 0.f) wn = -wn;

        if (-dot3(wn, dir) < safe::DotMin || 0.f == squad.S || (dot3(dir, n) <= 0.f && !total_sphere)) continue;

        buffer[current_sample++] = {{{ws[0], ws[1], ws[2], sample_pdf}}, wn, dir, {{uv[0], uv[1], 0.f, 0.f}}};
    }
    return current_sample;
}

// Rectangle.pdf, rectangle.zig:554-575
float pdf(Vec4f p, const Fragment& frag, uint32_t num_samples) {
    const float   nsf   = float(num_samples);
    const Vec4f   scale = frag.isec.trafo.scale();
    const Vec4f   lp    = frag.isec.trafo.worldToFramePoint(p);
    const SphQuad squad = SphQuad::init(scale, lp);
    return nsf * squad.pdf(scale);
}

}  // namespace rectangle

namespace disk {

bool intersect(const Ray& ray, const Trafo& trafo, Intersection& isec) {  // disk.zig:28-58
    const Vec4f normal = trafo.r[2];
    const float d      = dot3(normal, trafo.position);
    const float denom  = -dot3(normal, ray.direction);
    const float numer  = dot3(normal, ray.origin) - d;
    const float hit_t  = numer / denom;

    if (hit_t >= ray.min_t && ray.max_t >= hit_t) {
        const Vec4f p      = ray.point(hit_t);
        const Vec4f k      = p - trafo.position;
        const float l      = dot3(k, k);
        const float radius = 0.5f * trafo.scaleX();
        if (l <= radius * radius) {
            const Vec4f sk = k / splat(radius);
            isec.u         = -dot3(trafo.r[0], sk);
            isec.v         = -dot3(trafo.r[1], sk);
            isec.t         = hit_t;
            isec.primitive = 0;
            isec.trafo     = trafo;
            return true;
        }
    }
    return false;
}

bool intersectP(const Ray& ray, const Trafo& trafo) {  // :115-134
    Intersection unused;
    return intersect(ray, trafo, unused);
}

void fragment(const Ray& ray, Fragment& frag) {  // :98-113
    const float u = frag.isec.u, v = frag.isec.v;
    frag.p     = ray.point(frag.isec.t);
    frag.t     = -frag.isec.trafo.r[0];
    frag.b     = -frag.isec.trafo.r[1];
    frag.n     = frag.isec.trafo.r[2];
    frag.geo_n = frag.isec.trafo.r[2];
    frag.uvw   = {{0.5f * (u + 1.f), 0.5f * (v + 1.f), 0.f, 0.f}};
    frag.part  = 0;
}

// DiskSamplerData + EquiAngularSampling, disk.zig:181-250
struct DiskSamplerData {
    Vec4f xd, yd;
    static DiskSamplerData init(Vec4f p) {
        const Vec4f td = {{p[1], -p[0], 0.f, 0.f}};
        const Vec4f xd = (0.f == td[0] && 0.f == td[1] && 0.f == td[2]) ? Vec4f{{1.f, 0.f, 0.f, 0.f}} : normalize3(td);
        return {xd, {{-xd[1], xd[0], 0.f, 0.f}}};
    }
};

struct EquiAngularSampling {
    float offset, min_t, max_t, scale, scale_sqr, angle_min, angle_extent;

    static EquiAngularSampling init(Vec4f source, Vec4f origin, Vec4f direction, float min_t, float max_t) {
        const float offset    = dot3(direction, source - origin) / squaredLength3(direction);
        const float scale_sqr = squaredLength3((origin + splat(offset) * direction) - source);
        const float scale     = std::sqrt(scale_sqr);
        const float inv_scale = 0.f == scale ? 0.f : 1.f / scale;
        const float angle_min = std::atan((min_t - offset) * inv_scale);
        const float angle_max = std::atan((max_t - offset) * inv_scale);
        return {offset, min_t, max_t, scale, scale_sqr, angle_min, angle_max - angle_min};
    }
    float sample(float u, float& t) const {
        const float lt = scale * std::tan(angle_min + u * angle_extent);
        const float p  = scale / (angle_extent * (scale_sqr + lt * lt));
        t              = clamp(lt + offset, min_t, max_t);
        return p;
    }
    float pdf(float t) const {
        if (min_t <= t && t < max_t) {
            const float lt = t - offset;
            return scale / (angle_extent * (scale_sqr + lt * lt));
        }
        return 0.f;
    }
    float pdfAndSample(float t, float& u) const {
        const float lt = t - offset;
        u              = saturate((std::atan(lt / scale) - angle_min) / angle_extent);
        return scale / (angle_extent * (scale_sqr + lt * lt));
    }
};

// Disk.sampleTo, disk.zig:252-332 (UseEquiAngularSampling = true)
uint32_t sampleTo(Vec4f p, Vec4f n, const Trafo& trafo, bool two_sided, bool total_sphere, uint32_t num_samples, Sampler& sampler,
                  SampleTo* buffer) {
    const float nsf    = float(num_samples);
    const float radius = 0.5f * trafo.scaleX();

    const Vec4f               lp   = trafo.worldToFramePoint(p);
    const DiskSamplerData     dsd  = DiskSamplerData::init(lp);
    const EquiAngularSampling eas0 = EquiAngularSampling::init(lp, splat(0.f), dsd.yd, -radius, radius);
    if (0.f == eas0.angle_extent) return 0;

    uint32_t current_sample = 0;
    for (uint32_t i = 0; i < num_samples; ++i) {
        const Vec2f r2 = sampler.sample2D();
        float       xy[2];
        diskConcentric(r2.v, xy);

        float u    = xy[0];
        float pdf_ = std::sqrt(1.f - u * u) / (0.25f * kPi);
        u          = (u + 1.f) * 0.5f;

        float y_coord;
        pdf_ *= eas0.sample(u, y_coord);

        const float x_chord = std::sqrt(radius * radius - y_coord * y_coord);
        if (0.f == x_chord) continue;

        const EquiAngularSampling eas1 = EquiAngularSampling::init(lp, splat(y_coord) * dsd.yd, dsd.xd, -x_chord, x_chord);
        if (0.f == eas1.angle_extent) continue;

        float x_coord;
        pdf_ *= eas1.sample(sampler.sample1D(), x_coord);

        const Vec4f l_direction = splat(x_coord) * dsd.xd + splat(y_coord) * dsd.yd - lp;
        const Vec4f axis        = trafo.objectToWorldNormal(l_direction);
        const Vec4f ws          = p + axis;

        Vec4f wn = trafo.r[2];
        if (two_sided && dot3(wn, ws - p) > 0.f) wn = -wn;

        const float sl  = squaredLength3(axis);
        const Vec4f dir = axis / splat(std::sqrt(sl));
        const float c   = -dot3(wn, dir);
        if (c < safe::DotMin || (dot3(dir, n) <= 0.f && !total_sphere)) continue;

        const float v            = (xy[1] + 1.f) * 0.5f;
        buffer[current_sample++] = {{{ws[0], ws[1], ws[2], (nsf * pdf_ * sl) / c}}, wn, dir, {{u, v, 0.f, 0.f}}};
    }
    return current_sample;
}

// Disk.pdf, disk.zig:492-533
float pdf(Vec4f dir, Vec4f p, const Fragment& frag, uint32_t num_samples) {
    const float c      = std::fabs(dot3(frag.isec.trafo.r[2], dir));
    const float nsf    = float(num_samples);
    const float radius = 0.5f * frag.isec.trafo.scaleX();
    const float sl     = squaredDistance3(p, frag.p);

    const Vec4f               lp      = frag.isec.trafo.worldToFramePoint(p);
    const DiskSamplerData     dsd     = DiskSamplerData::init(lp);
    const EquiAngularSampling eas0    = EquiAngularSampling::init(lp, splat(0.f), dsd.yd, -radius, radius);
    const Vec4f               l_point = frag.isec.trafo.worldToFramePoint(frag.p);
    const float               y_coord = dot3(l_point, dsd.yd);

    float       u;
    const float eas_pdf = eas0.pdfAndSample(y_coord, u);
    u                   = u * 2.f - 1.f;
    float pdf_          = std::sqrt(1.f - u * u) / (0.25f * kPi);
    pdf_ *= eas_pdf;

    const float               x_chord = std::sqrt(radius * radius - y_coord * y_coord);
    const EquiAngularSampling eas1    = EquiAngularSampling::init(lp, splat(y_coord) * dsd.yd, dsd.xd, -x_chord, x_chord);
    const float               x_coord = dot3(l_point, dsd.xd);
    pdf_ *= eas1.pdf(x_coord);

    return (nsf * pdf_ * sl) / c;
}

}  // namespace disk

namespace cube {

const AABB kUnit = {{{{-0.5f, -0.5f, -0.5f, -0.5f}}, {{0.5f, 0.5f, 0.5f, 0.5f}}}};

float aabbIntersectP(const AABB& box, const Ray& ray) {  // aabb.zig:62-84
    const Vec4f lower = (box.bounds[0] - ray.origin) * ray.inv_direction;
    const Vec4f upper = (box.bounds[1] - ray.origin) * ray.inv_direction;
    const Vec4f t0    = min4(lower, upper);
    const Vec4f t1    = max4(lower, upper);

    const float imin = max(max(t0[0], t0[1]), t0[2]);
    const float imax = min(min(t1[0], t1[1]), t1[2]);

    const float tboxmin = max(imin, ray.min_t);
    const float tboxmax = min(imax, ray.max_t);

    if (tboxmin <= tboxmax) return imin < ray.min_t ? imax : imin;
    return FLT_MAX;
}

bool intersect(const Ray& ray, const Trafo& trafo, Intersection& isec) {  // cube.zig:24-38
    const Ray   local_ray = trafo.worldToObjectRay(ray);
    const float hit_t     = aabbIntersectP(kUnit, local_ray);
    if (hit_t < ray.max_t) {
        isec.t         = hit_t;
        isec.primitive = 0;
        isec.trafo     = trafo;
        return true;
    }
    return false;
}

void fragment(const Ray& ray, Fragment& frag) {  // cube.zig:40-62
    const float hit_t = frag.isec.t;
    frag.p            = ray.point(hit_t);

    const Ray   local_ray = frag.isec.trafo.worldToObjectRay(ray);
    const Vec4f local_p   = local_ray.point(hit_t);
    Vec4f       distance;
    for (int i = 0; i < 4; ++i) distance[i] = std::fabs(0.5f - std::fabs(local_p[i]));

    const uint32_t i = indexMinComponent3(distance);
    const float    s = std::copysign(1.f, local_p[int(i)]);
    const Vec4f    n = splat(s) * frag.isec.trafo.r[i];

    frag.part  = 0;
    frag.geo_n = n;
    frag.n     = n;
    frag.uvw   = splat(0.f);
    orthonormalBasis3(n, frag.t, frag.b);
}

bool intersectP(const Ray& ray, const Trafo& trafo) {  // cube.zig:64-69
    const Ray local_ray = trafo.worldToObjectRay(ray);
    return kUnit.intersect(local_ray);
}

}  // namespace cube

namespace sphere {

bool intersect(const Ray& ray, const Trafo& trafo, Intersection& isec) {  // sphere.zig:28-62
    const float idl = 1.f / length3(ray.direction);
    const Vec4f nd  = ray.direction * splat(idl);

    const Vec4f v = trafo.position - ray.origin;
    const float b = dot3(nd, v);

    const Vec4f remedy_term  = v - splat(b) * nd;
    const float radius       = 0.5f * trafo.scaleX();
    const float discriminant = radius * radius - dot3(remedy_term, remedy_term);

    if (discriminant > 0.f) {
        const float dist = std::sqrt(discriminant);

        const float t0 = (b - dist) * idl;
        if (t0 >= ray.min_t && ray.max_t >= t0) {
            isec.t         = t0;
            isec.primitive = 0;
            isec.trafo     = trafo;
            return true;
        }
        const float t1 = (b + dist) * idl;
        if (t1 >= ray.min_t && ray.max_t >= t1) {
            isec.t         = t1;
            isec.primitive = 0;
            isec.trafo     = trafo;
            return true;
        }
    }
    return false;
}

void fragment(const Ray& ray, Fragment& frag) {  // sphere.zig:64-92
    const Vec4f p = ray.point(frag.isec.t);
    const Vec4f n = normalize3(p - frag.isec.trafo.position);

    frag.p     = p;
    frag.geo_n = n;
    frag.n     = n;
    frag.part  = 0;

    const Vec4f xyz   = normalize3(frag.isec.trafo.worldToObjectNormal(n));
    const float phi   = -std::atan2(xyz[0], xyz[2]) + kPi;
    const float theta = std::acos(xyz[1]);

    const float sin_phi   = std::sin(phi);
    const float cos_phi   = std::cos(phi);
    const float sin_theta = max(std::sin(theta), 0.00001f);

    const Vec4f t = normalize3(frag.isec.trafo.objectToWorldNormal({{sin_theta * cos_phi, 0.f, sin_theta * sin_phi, 0.f}}));

    frag.t   = t;
    frag.b   = -cross3(t, n);
    frag.uvw = {{phi * (0.5f * kPiInv), theta * kPiInv, 0.f, 0.f}};
}

bool intersectP(const Ray& ray, const Trafo& trafo) {
    Intersection unused;
    return intersect(ray, trafo, unused);
}

inline float conePdfUniform(float one_minus_cos_theta_max) {  // sampling.zig:103-106
    return 1.f / ((2.f * kPi) * max(one_minus_cos_theta_max, 1.0e-20f));
}

// Sphere.sampleTo, sphere.zig:323-393
uint32_t sampleTo(Vec4f p, Vec4f n, const Trafo& trafo, bool total_sphere, uint32_t num_samples, Sampler& sampler, SampleTo* buffer) {
    const Vec4f v = trafo.position - p;
    const float l = length3(v);
    const float r = 0.5f * trafo.scaleX();
    if (l <= (r + 0.0000001f)) return 0;

    const Vec4f z     = splat(1.f / l) * v;
    const Frame frame = Frame::init(z);
    const float nsf   = float(num_samples);

    uint32_t current_sample = 0;
    for (uint32_t i = 0; i < num_samples; ++i) {
        const float sin_theta_max           = r / l;
        const float sin2_theta_max          = sin_theta_max * sin_theta_max;
        const float cos_theta_max           = std::sqrt(1.f - sin2_theta_max);
        float       one_minus_cos_theta_max = 1.f - cos_theta_max;

        const Vec2f s2 = sampler.sample2D();

        float cos_theta  = (cos_theta_max - 1.f) * s2[0] + 1.f;
        float sin2_theta = 1.f - (cos_theta * cos_theta);
        if (sin2_theta_max < 0.00068523f) {
            sin2_theta              = sin2_theta_max * s2[0];
            cos_theta               = std::sqrt(1.f - sin2_theta);
            one_minus_cos_theta_max = 0.5f * sin2_theta_max;
        }

        const float cos_alpha = min(sin2_theta / sin_theta_max + cos_theta * std::sqrt(1.f - min(sin2_theta / sin2_theta_max, 1.f)), 1.f);
        const float sin_alpha = std::sqrt(1.f - cos_alpha * cos_alpha);
        const float phi       = s2[1] * (2.f * kPi);

        // smpl.sphereDirection, sampling.zig:78-83
        const float sin_phi = std::sin(phi), cos_phi = std::cos(phi);
        const Vec4f w  = {{cos_phi * sin_alpha, sin_phi * sin_alpha, cos_alpha, 0.f}};
        const Vec4f wn = frame.frameToWorld(-w);
        const Vec4f lp = trafo.position + splat(r) * wn;
        const Vec4f dir = normalize3(lp - p);
        if (dot3(dir, n) <= 0.f && !total_sphere) continue;

        SampleTo& out = buffer[current_sample++];
        out.p         = {{lp[0], lp[1], lp[2], nsf * conePdfUniform(one_minus_cos_theta_max)}};
        out.n         = wn;
        out.wi        = dir;
        out.uvw       = splat(0.f);
    }
    return current_sample;
}

// Sphere.pdf, sphere.zig:472-487
float pdf(Vec4f p, const Trafo& trafo, uint32_t num_samples) {
    const Vec4f v              = trafo.position - p;
    const float l2             = squaredLength3(v);
    const float r              = 0.5f * trafo.scaleX();
    const float sin2_theta_max = (r * r) / l2;
    const float one_minus_cos_theta_max = sin2_theta_max < 0.00068523f ? 0.5f * sin2_theta_max : 1.f - std::sqrt(max(1.f - sin2_theta_max, 0.f));
    return float(num_samples) * conePdfUniform(one_minus_cos_theta_max);
}

}  // namespace sphere

namespace distant {  // shape/distant.zig:22-146

float solidAngle(float radius) { return (2.f * kPi) * (1.f - std::sqrt(1.f / (radius * radius + 1.f))); }  // :143-145

bool intersect(const Ray& ray, const Trafo& trafo, Intersection& isec) {  // :22-54
    const float radius = trafo.scaleX();
    const Vec4f n      = trafo.r[2];
    const float b      = dot3(n, ray.direction);
    if (b > 0.f || ray.max_t < RayMaxT || radius <= 0.f) return false;

    const float det = (b * b) - dot3(n, n) + (radius * radius);
    if (det >= 0.f) {
        const Vec4f k  = ray.direction - n;
        const Vec4f sk = k / splat(radius);
        isec.u         = dot3(trafo.r[0], sk);
        isec.v         = dot3(trafo.r[1], sk);
        isec.primitive = 0;
        isec.t         = RayMaxT;
        isec.trafo     = trafo;
        return true;
    }
    return false;
}

void fragment(const Ray& ray, Fragment& frag) {  // :56-76
    frag.p       = splat(RayMaxT) * ray.direction;
    const Vec4f n = frag.isec.trafo.r[2];
    frag.geo_n   = n;
    frag.t       = frag.isec.trafo.r[0];
    frag.b       = frag.isec.trafo.r[1];
    frag.n       = n;
    frag.uvw     = {{(frag.isec.u + 1.f) * 0.5f, (frag.isec.v + 1.f) * 0.5f, 0.f, 0.f}};
    frag.part    = 0;
}

uint32_t sampleTo(Vec4f n, const Trafo& trafo, bool total_sphere, Sampler& sampler, SampleTo* buffer) {  // :78-107
    const float radius = trafo.scaleX();
    if (radius <= 0.f) return 0;

    const Vec2f r2 = sampler.sample2D();
    float       xy[2];
    diskConcentric(r2.v, xy);

    const Vec4f ls = {{xy[0], xy[1], 0.f, 0.f}};
    // Mat3x3.transformVector, matrix3x3.zig:113-127 (the rotation rows carry the scale in lane 3, which is not read)
    Vec4f tv = splat(ls[0]) * trafo.r[0];
    tv       = mulAdd(splat(ls[1]), trafo.r[1], tv);
    tv       = mulAdd(splat(ls[2]), trafo.r[2], tv);
    const Vec4f ws  = splat(radius) * tv;
    const Vec4f dir = normalize3(ws - trafo.r[2]);

    if (dot3(dir, n) <= 0.f && !total_sphere) return 0;

    const float solid_angle = solidAngle(radius);
    const Vec4f p           = splat(RayMaxT) * dir;
    buffer[0].p             = {{p[0], p[1], p[2], 1.f / solid_angle}};
    buffer[0].n             = trafo.r[2];
    buffer[0].wi            = dir;
    buffer[0].uvw           = splat(0.f);
    return 1;
}

}  // namespace distant

// ---- emission images: Distribution1D / Distribution2D / ImageImpl / texture lookup ------------
namespace image {

// Distribution1D over a ZygpuImageSampler cdf row. The lookup table is the reference's accelerator for its linear search
// (initLut / map / search, distribution_1d.zig:133-165, 250-258), rebuilt here from the cdf.
struct Distribution1D {
    const float*          cdf = nullptr;
    uint32_t              size = 0;  // entries of cdf
    std::vector<uint32_t> lut;
    float                 lut_range = 0.f;

    void configure(const float* c, uint32_t num_data) {  // configure + initLut
        cdf  = c;
        size = num_data + 1;
        uint32_t lut_size = num_data / 16;
        lut_size          = std::min(std::max(lut_size, 1u), size);
        lut.assign(lut_size + 1, 0u);
        lut_range = float(lut_size);
        lut[0]    = 1;
        uint32_t border = 0;
        for (uint32_t i = 1; i < size; ++i) {
            const uint32_t mapped = map(cdf[i]);
            if (mapped > border) {
                for (uint32_t k = border + 1; k <= mapped && k < lut.size(); ++k) lut[k] = i;
                border = mapped;
            }
        }
    }
    uint32_t map(float s) const { return uint32_t(s * lut_range); }
    uint32_t sample(float r) const {  // :50-54
        const uint32_t bucket = map(r);
        uint32_t       i      = lut[bucket];
        const uint32_t end    = size - 1;
        for (; i < end; ++i) {
            if (cdf[i] >= r) return i - 1;
        }
        return end - 1;
    }
    void sampleDiscrete(float r, uint32_t& offset, float& pdf) const {  // :56-60
        offset = sample(r);
        pdf    = cdf[offset + 1] - cdf[offset];
    }
    void sampleContinuous(float r, float& offset, float& pdf) const {  // :62-75
        const uint32_t o = sample(r);
        const float    c = cdf[o + 1];
        const float    v = c - cdf[o];
        if (0.f == v) {
            offset = 0.f;
            pdf    = 0.f;
            return;
        }
        const float t = (c - r) / v;
        offset        = (float(o) + t) / float(size - 1);
        pdf           = v;
    }
    float pdfI(uint32_t index) const { return cdf[index + 1] - cdf[index]; }
    float pdfF(float u) const {  // :81-86
        const uint32_t len = size;
        const uint32_t o   = std::min(uint32_t(u * float(len - 1)), len - 2);
        return cdf[o + 1] - cdf[o];
    }
};

struct Sampler2D {  // shape_sampler.ImageImpl + Distribution2D, shape_sampler.zig:128-152, distribution_2d.zig:67-87
    const ZygpuImageSampler*    img = nullptr;
    Distribution1D              marginal;
    std::vector<Distribution1D> conditional;

    explicit Sampler2D(const ZygpuImageSampler& is) : img(&is), conditional(is.marginal_cdf ? is.height : 0) {
        if (!is.marginal_cdf) return;  // an image that is only looked up (a colour map)
        marginal.configure(is.marginal_cdf, is.height);
        for (uint32_t y = 0; y < is.height; ++y) conditional[y].configure(is.conditional_cdf + size_t(y) * (is.width + 1), is.width);
    }
    static float address(uint32_t mode, float x) {  // sampler_mode.zig:21-26
        return 0 == mode ? clamp(x, 0.f, 1.f) : x - std::floor(x);
    }
    static int32_t coord(uint32_t mode, int32_t c, int32_t end) {  // :35-40, 76-80
        if (0 == mode) return std::max(std::min(c, end - 1), 0);
        const int32_t m = c % end;
        return m < 0 ? m + end : m;
    }
    void sample(float r0, float r1, float uv[2], float& pdf) const {  // ImageImpl.sample
        float v, vp, u, up;
        marginal.sampleContinuous(r1, v, vp);
        const uint32_t n = uint32_t(conditional.size());
        const uint32_t c = std::min(uint32_t(v * float(n)), n - 1);
        conditional[c].sampleContinuous(r0, u, up);
        uv[0] = u;
        uv[1] = v;
        pdf   = (up * vp) * img->total_weight;
    }
    float pdf(float u, float v) const {  // ImageImpl.pdf
        const float    au = address(img->address_u, u), av = address(img->address_v, v);
        const float    v_pdf = marginal.pdfF(av);
        const uint32_t n     = uint32_t(conditional.size());
        const uint32_t c     = std::min(uint32_t(av * float(n)), n - 1);
        return (conditional[c].pdfF(au) * v_pdf) * img->total_weight;
    }
    // ts.sample2D_3 for an image texture, texture_sampler.zig:63-79, 99-124 (Nearest), 126-170 (LinearStochastic)
    Vec4f texel(float u, float v, float r) const {
        const int32_t d[2] = {int32_t(img->width), int32_t(img->height)};
        const float   st[2] = {img->scale[0] * u, img->scale[1] * v};
        int32_t       xy[2];
        if (0 == img->filter) {
            xy[0] = std::min(int32_t(address(img->address_u, st[0]) * float(d[0])), d[0] - 1);
            xy[1] = std::min(int32_t(address(img->address_v, st[1]) * float(d[1])), d[1] - 1);
        } else {
            const float mst[2] = {address(img->address_u, st[0]) * float(d[0]) - 0.5f, address(img->address_v, st[1]) * float(d[1]) - 0.5f};
            const float fst[2] = {std::floor(mst[0]), std::floor(mst[1])};
            const float w[2]   = {mst[0] - fst[0], mst[1] - fst[1]};
            const float omw[2] = {1.f - w[0], 1.f - w[1]};
            xy[0]              = int32_t(fst[0]);
            xy[1]              = int32_t(fst[1]);
            int32_t index      = 0;
            float   threshold  = omw[0] * omw[1];
            index += r > threshold ? 1 : 0;
            threshold = std::fmaf(w[0], omw[1], threshold);
            index += r > threshold ? 1 : 0;
            threshold = std::fmaf(omw[0], w[1], threshold);
            index += r > threshold ? 1 : 0;
            xy[0] += index & 1;
            xy[1] += (index & 2) >> 1;
            xy[0] = coord(img->address_u, xy[0], d[0]);
            xy[1] = coord(img->address_v, xy[1], d[1]);
        }
        const float* px = img->pixels + 3 * (size_t(xy[1]) * img->width + xy[0]);
        return {{px[0], px[1], px[2], 0.f}};
    }
};

}  // namespace image

namespace rectangle {

// Rectangle.sampleMaterialTo, rectangle.zig (image-mapped area light: texels picked through the material's Distribution2D)
uint32_t sampleMaterialTo(Vec4f p, Vec4f n, const Trafo& trafo, bool two_sided, bool total_sphere, uint32_t num_samples,
                          const image::Sampler2D& shape_sampler, Sampler& sampler, SampleTo* buffer) {
    const float nsf   = float(num_samples);
    const Vec4f scale = trafo.scale();
    const float area  = scale[0] * scale[1];

    uint32_t current_sample = 0;
    for (uint32_t i = 0; i < num_samples; ++i) {
        const Vec2f r2 = sampler.sample2D();
        float       uv[2], rs_pdf;
        shape_sampler.sample(r2.v[0], r2.v[1], uv, rs_pdf);
        if (0.f == rs_pdf) continue;

        const Vec4f ls   = {{-1.f * uv[0] + 0.5f, -1.f * uv[1] + 0.5f, 0.f, 0.f}};
        const Vec4f ws   = trafo.objectToWorldPoint(ls);
        const Vec4f axis = ws - p;

        Vec4f wn = trafo.r[2];
        if (two_sided && dot3(wn, axis) > 0.f) wn = -wn;

        const float sl  = squaredLength3(axis);
        const float t   = std::sqrt(sl);
        const Vec4f dir = axis / splat(t);
        const float c   = -dot3(wn, dir);
        if (c < safe::DotMin || (dot3(dir, n) <= 0.f && !total_sphere)) continue;

        SampleTo& out = buffer[current_sample++];
        out.p         = {{ws[0], ws[1], ws[2], (nsf * rs_pdf * sl) / (c * area)}};
        out.n         = wn;
        out.wi        = dir;
        out.uvw       = {{uv[0], uv[1], 0.f, 0.f}};
    }
    return current_sample;
}

// Rectangle.materialPdf
float materialPdf(Vec4f dir, Vec4f p, const Fragment& frag, uint32_t num_samples, const image::Sampler2D& shape_sampler) {
    const float c            = std::fabs(dot3(frag.isec.trafo.r[2], dir));
    const Vec4f scale        = frag.isec.trafo.scale();
    const float area         = scale[0] * scale[1];
    const float sl           = squaredDistance3(p, frag.p);
    const float material_pdf = shape_sampler.pdf(frag.uvw[0], frag.uvw[1]) * float(num_samples);
    return (material_pdf * sl) / (c * area);
}

}  // namespace rectangle

namespace canopy {  // shape/canopy.zig


constexpr float Eps = -0.0005f;

bool intersect(const Ray& ray, const Trafo& trafo, Intersection& isec) {  // :27-39
    if (ray.max_t < RayMaxT || dot3(ray.direction, trafo.r[2]) < Eps) return false;
    isec.primitive = 0;
    isec.t         = RayMaxT;
    isec.u = isec.v = 0.f;
    isec.trafo     = trafo;
    return true;
}

void hemisphereToDiskEquidistant(Vec4f dir, float disk[2]) {  // :164-177
    const float colatitude = std::acos(dir[2]);
    const float longitude  = std::atan2(-dir[1], dir[0]);
    const float r          = colatitude * (kPiInv * 2.f);
    const float sin_lon    = std::sin(longitude);
    const float cos_lon    = std::cos(longitude);
    disk[0]                = r * cos_lon;
    disk[1]                = r * sin_lon;
}

Vec4f diskToHemisphereEquidistant(const float uv[2]) {  // :179-202
    const float longitude  = std::atan2(-uv[1], uv[0]);
    const float r          = std::sqrt(uv[0] * uv[0] + uv[1] * uv[1]);
    const float colatitude = r * (kPi / 2.f);
    const float sin_col    = std::sin(colatitude);
    const float cos_col    = std::cos(colatitude);
    const float sin_lon    = std::sin(longitude);
    const float cos_lon    = std::cos(longitude);
    return {{sin_col * cos_lon, sin_col * sin_lon, cos_col, 0.f}};
}

void fragment(const Ray& ray, Fragment& frag) {  // :41-62
    const Trafo& trafo = frag.isec.trafo;
    // Mat3x3.transformVectorTransposed, matrix3x3.zig
    const Vec4f d   = ray.direction;
    const Vec4f xyz = normalize3(Vec4f{{dot3(d, trafo.r[0]), dot3(d, trafo.r[1]), dot3(d, trafo.r[2]), 0.f}});
    float       disk[2];
    hemisphereToDiskEquidistant(xyz, disk);
    frag.uvw = {{0.5f * disk[0] + 0.5f, 0.5f * disk[1] + 0.5f, 0.f, 0.f}};

    const Vec4f dir = {{d[0], d[1], d[2], 0.f}};
    frag.p          = splat(RayMaxT) * dir;
    const Vec4f n   = -dir;
    frag.geo_n      = n;
    frag.t          = trafo.r[0];
    frag.b          = trafo.r[1];
    frag.n          = n;
    frag.part       = 0;
}

// Canopy.sampleMaterialTo, :94-131
uint32_t sampleMaterialTo(Vec4f n, const Trafo& trafo, bool total_sphere, const image::Sampler2D& shape_sampler, Sampler& sampler,
                          SampleTo* buffer) {
    const Vec2f r2 = sampler.sample2D();
    float       uv[2], pdf;
    shape_sampler.sample(r2.v[0], r2.v[1], uv, pdf);
    if (0.f == pdf) return 0;

    const float disk[2] = {2.f * uv[0] - 1.f, 2.f * uv[1] - 1.f};
    const float z       = disk[0] * disk[0] + disk[1] * disk[1];
    if (z > 1.f) return 0;

    const Vec4f dir_l = diskToHemisphereEquidistant(disk);
    // Mat3x3.transformVector, matrix3x3.zig:113-127
    Vec4f dir = splat(dir_l[0]) * trafo.r[0];
    dir       = mulAdd(splat(dir_l[1]), trafo.r[1], dir);
    dir       = mulAdd(splat(dir_l[2]), trafo.r[2], dir);
    dir[3]    = 0.f;

    if (dot3(dir, n) <= 0.f && !total_sphere) return 0;

    const Vec4f p = splat(RayMaxT) * dir;
    buffer[0].p   = {{p[0], p[1], p[2], pdf / (2.f * kPi)}};
    buffer[0].n   = -dir;
    buffer[0].wi  = dir;
    buffer[0].uvw = {{uv[0], uv[1], 0.f, 0.f}};
    return 1;
}

}  // namespace canopy

namespace mesh {

// Mesh.fragment, triangle_mesh.zig:310-335 + Data.interpolateData / normal, triangle_data.zig:106-149
void fragment(const ZoMesh& m, Fragment& frag) {
    const uint32_t prim = frag.isec.primitive;
    frag.part           = m.parts[prim];

    const float hit_u = frag.isec.u;
    const float hit_v = frag.isec.v;

    const uint32_t* tri = m.triangles + size_t(prim) * 3;

    auto position = [&](uint32_t i) -> Vec4f { return {{m.positions[size_t(i) * 3], m.positions[size_t(i) * 3 + 1], m.positions[size_t(i) * 3 + 2], 0.f}}; };
    auto shadingNormal = [&](uint32_t i) -> Vec4f {  // enc.decompressNormal, encoding.zig:91-108
        const float o0 = std::fmaf(float(m.normals[size_t(i) * 2]), 1.f / 32768.f, -1.f);
        const float o1 = std::fmaf(float(m.normals[size_t(i) * 2 + 1]), 1.f / 32768.f, -1.f);
        Vec4f       v  = {{o0, o1, -1.f + std::fabs(o0) + std::fabs(o1), 0.f}};
        const float t  = max(v[2], 0.f);
        v[0] += v[0] > 0.f ? -t : t;
        v[1] += v[1] > 0.f ? -t : t;
        return normalize3(v);
    };
    auto interpolate3 = [](Vec4f a, Vec4f b, Vec4f c, float u, float v) -> Vec4f {  // triangle.zig:142-149
        const float w     = 1.f - u - v;
        const Vec4f temp0 = mulAdd(b, splat(u), c * splat(v));
        return mulAdd(a, splat(w), temp0);
    };

    const Vec4f pa = position(tri[0]);
    const Vec4f pb = position(tri[1]);
    const Vec4f pc = position(tri[2]);

    const Vec4f geo_n = normalize3(cross3(pb - pa, pc - pa));
    frag.geo_n        = frag.isec.trafo.objectToWorldNormal(geo_n);

    const Vec4f p = interpolate3(pa, pb, pc, hit_u, hit_v);

    const float uva[2] = {m.uvs[size_t(tri[0]) * 2], m.uvs[size_t(tri[0]) * 2 + 1]};
    const float uvb[2] = {m.uvs[size_t(tri[1]) * 2], m.uvs[size_t(tri[1]) * 2 + 1]};
    const float uvc[2] = {m.uvs[size_t(tri[2]) * 2], m.uvs[size_t(tri[2]) * 2 + 1]};
    const float w      = 1.f - hit_u - hit_v;  // triangle.zig:133-140
    const float uv[2]  = {std::fmaf(uva[0], w, std::fmaf(uvb[0], hit_u, uvc[0] * hit_v)),
                          std::fmaf(uva[1], w, std::fmaf(uvb[1], hit_u, uvc[1] * hit_v))};

    const Vec4f nb = shadingNormal(tri[1]);
    const Vec4f na = shadingNormal(tri[0]);
    const Vec4f nc = shadingNormal(tri[2]);
    const Vec4f ni = normalize3(interpolate3(na, nb, nc, hit_u, hit_v));

    // triangle.positionDifferentials, triangle.zig:102-131
    const float duv02[2]    = {uva[0] - uvc[0], uva[1] - uvc[1]};
    const float duv12[2]    = {uvb[0] - uvc[0], uvb[1] - uvc[1]};
    const float determinant = duv02[0] * duv12[1] - duv02[1] * duv12[0];

    Vec4f       dpdu, dpdv;
    const Vec4f dp02 = pa - pc;
    const Vec4f dp12 = pb - pc;
    if (0.f == std::fabs(determinant)) {
        const Vec4f ng = normalize3(cross3(pc - pa, pb - pa));
        if (std::fabs(ng[0]) > std::fabs(ng[1])) {
            dpdu = Vec4f{{-ng[2], 0.f, ng[0], 0.f}} / splat(std::sqrt(ng[0] * ng[0] + ng[2] * ng[2]));
        } else {
            dpdu = Vec4f{{0.f, ng[2], -ng[1], 0.f}} / splat(std::sqrt(ng[1] * ng[1] + ng[2] * ng[2]));
        }
        dpdv = cross3(ng, dpdu);
    } else {
        const float invdet = 1.f / determinant;
        dpdu               = splat(invdet) * mulAdd(splat(duv12[1]), dp02, splat(-duv02[1]) * dp12);
        dpdv               = splat(invdet) * mulAdd(splat(-duv12[0]), dp02, splat(duv02[0]) * dp12);
    }

    const Vec4f t = normalize3(gramSchmidt(dpdu, ni));
    const Vec4f b = normalize3(gramSchmidt(dpdv, ni));

    frag.p   = frag.isec.trafo.objectToWorldPoint(p);
    frag.t   = frag.isec.trafo.objectToWorldNormal(t);
    frag.b   = frag.isec.trafo.objectToWorldNormal(b);
    frag.n   = frag.isec.trafo.objectToWorldNormal(ni);
    frag.uvw = {{uv[0], uv[1], 0.f, 0.f}};
}

// triangle_mesh.zig:390-394
inline Vec4f orthogonalize(Vec4f a, Vec4f b) { return normalize3(mulAdd(splat(-dot3(a, b)), a, b)); }

// triangle.barycentricCoords, triangle.zig:82-100
inline void barycentricCoords(Vec4f dir, Vec4f a, Vec4f b, Vec4f c, float uv[2]) {
    const Vec4f e1      = b - a;
    const Vec4f e2      = c - a;
    const Vec4f tvec    = -a;
    const Vec4f pvec    = cross3(dir, e2);
    const Vec4f qvec    = cross3(tvec, e1);
    const float e1_d_pv = dot3(e1, pvec);
    const float tv_d_pv = dot3(tvec, pvec);
    const float di_d_qv = dot3(dir, qvec);
    const float inv_det = 1.f / e1_d_pv;
    uv[0]               = tv_d_pv * inv_det;
    uv[1]               = di_d_qv * inv_det;
}

struct SphericalSample {
    Vec4f dir;
    float uv[2];
    float pdf;
};

inline float sphericalArea(Vec4f A, Vec4f B, Vec4f C, float& cos_alpha, float& alpha) {
    const Vec4f BA = orthogonalize(A, B - A);
    const Vec4f CA = orthogonalize(A, C - A);
    const Vec4f AB = orthogonalize(B, A - B);
    const Vec4f CB = orthogonalize(B, C - B);
    const Vec4f BC = orthogonalize(C, B - C);
    const Vec4f AC = orthogonalize(C, A - C);
    cos_alpha         = clamp(dot3(BA, CA), -1.f, 1.f);
    alpha             = std::acos(cos_alpha);
    const float beta  = std::acos(clamp(dot3(AB, CB), -1.f, 1.f));
    const float gamma = std::acos(clamp(dot3(BC, AC), -1.f, 1.f));
    return alpha + beta + gamma - kPi;
}

// Stratified Sampling of Spherical Triangles, James Arvo. triangle_mesh.zig:402-462
inline bool sampleSpherical(Vec4f pos, Vec4f pa, Vec4f pb, Vec4f pc, const float r2[2], SphericalSample& out) {
    const Vec4f pap = pa - pos;
    const Vec4f pbp = pb - pos;
    const Vec4f pcp = pc - pos;

    const Vec4f A = normalize3(pap);
    const Vec4f B = normalize3(pbp);
    const Vec4f C = normalize3(pcp);

    float       cos_alpha, alpha;
    const float sarea = sphericalArea(A, B, C, cos_alpha, alpha);
    if (0.f == sarea) return false;

    const float cos_c = clamp(dot3(A, B), -1.f, 1.f);

    const float area_S      = r2[0] * sarea;
    const float angle_delta = area_S - alpha;
    const float p           = std::sin(angle_delta);
    const float q           = std::cos(angle_delta);

    const float sin_alpha = std::sqrt(1.f - cos_alpha * cos_alpha);
    const float u         = q - cos_alpha;
    const float v         = p + sin_alpha * cos_c;

    const float s   = clamp(((v * q - u * p) * cos_alpha - v) / ((v * p + u * q) * sin_alpha), -1.f, 1.f);
    const Vec4f C_s = splat(s) * A + splat(std::sqrt(1.f - s * s)) * orthogonalize(A, C);

    const float cs_b = dot3(C_s, B);
    const float z    = 1.f - r2[1] * (1.f - cs_b);
    const Vec4f P    = splat(z) * B + splat(std::sqrt(1.f - z * z)) * orthogonalize(B, C_s);

    out.dir = P;
    barycentricCoords(P, pap, pbp, pcp, out.uv);
    out.pdf = 1.f / sarea;
    return true;
}

inline float pdfSpherical(Vec4f pos, Vec4f pa, Vec4f pb, Vec4f pc) {  // :464-487
    float       cos_alpha, alpha;
    const float sarea = sphericalArea(normalize3(pa - pos), normalize3(pb - pos), normalize3(pc - pos), cos_alpha, alpha);
    return 1.f / sarea;
}

inline void triangleUniform(const float uv[2], float out[2]) {  // sampling.zig:39-47 (E. Heitz)
    if (uv[1] > uv[0]) {
        const float x = 0.5f * uv[0];
        out[0]        = x;
        out[1]        = uv[1] - x;
        return;
    }
    const float y = 0.5f * uv[1];
    out[0]        = uv[0] - y;
    out[1]        = y;
}

inline Vec4f interpolate3(Vec4f a, Vec4f b, Vec4f c, float u, float v) {  // triangle.zig:142-149
    const float w     = 1.f - u - v;
    const Vec4f temp0 = mulAdd(b, splat(u), splat(v) * c);
    return mulAdd(a, splat(w), temp0);
}

constexpr float AreaDistanceRatio = 0.001f;  // :489

}  // namespace mesh

// ---- scene -----------------------------------------------------------------------------------

struct Scene {
    const ZygpuScene& s;
    const ZygpuView&  view;
    GgxLuts           luts;
    const ZoMesh*     meshes;  // indexed by ZygpuProp.mesh

    std::vector<image::Sampler2D> image_samplers;  // shape_sampler.ImageImpl per ZygpuImageSampler
    image::Distribution1D         infinite_light_distribution;  // light_tree.zig:273

    Scene(const ZygpuScene& scene, const ZygpuView& v, const ZoMesh* ms) : s(scene), view(v), luts(scene.ggx_luts), meshes(ms) {
        image_samplers.reserve(scene.num_image_samplers);
        for (uint32_t i = 0; i < scene.num_image_samplers; ++i) image_samplers.emplace_back(scene.image_samplers[i]);
        if (scene.light_tree.num_infinite_lights > 0 && scene.light_tree.infinite_cdf) {
            infinite_light_distribution.configure(scene.light_tree.infinite_cdf, scene.light_tree.num_infinite_lights);
        }
    }

    Mesh treeOf(uint32_t mesh) const {
        return {static_cast<const Node*>(meshes[mesh].nodes), meshes[mesh].triangles, meshes[mesh].positions};
    }

    const ZygpuMaterial& propMaterial(uint32_t prop, uint32_t part) const {  // scene.zig:529-532
        return s.materials[s.material_ids[s.props[prop].parts_start + part]];
    }
    uint32_t propMaterialId(uint32_t prop, uint32_t part) const { return s.material_ids[s.props[prop].parts_start + part]; }  // scene.zig propMaterialId
    uint32_t propLightId(uint32_t prop, uint32_t part) const { return s.light_ids[s.props[prop].parts_start + part]; }
    Trafo    propTrafo(uint32_t prop) const { return Trafo::load(s.trafos[prop]); }
    AABB     propAabb(uint32_t prop) const { return {{load4(s.aabbs[prop].min), load4(s.aabbs[prop].max)}}; }

    static bool visible(uint32_t flags, uint32_t depth_surface) {  // prop.zig:38-48, sss = false
        return 0 == depth_surface ? 0 != (flags & ZYG_PROP_VISIBLE_IN_CAMERA) : 0 != (flags & ZYG_PROP_VISIBLE_IN_REFLECTION);
    }

    float shapeArea(uint32_t shape, Vec4f scale) const {  // shape.zig:143-156
        switch (shape) {
            case ZYG_SHAPE_RECTANGLE: return scale[0] * scale[1];
            case ZYG_SHAPE_DISK: return kPi * ((0.5f * scale[0]) * (0.5f * scale[0]));
            case ZYG_SHAPE_SPHERE: return (4.f * kPi) * pow2(0.5f * scale[0]);
            case ZYG_SHAPE_DISTANT: return distant::solidAngle(scale[0]);  // "the solid angle, not the area", shape.zig:146-148
            case ZYG_SHAPE_CANOPY: return 2.f * kPi;
            default: return 0.f;
        }
    }

    bool shapeIntersect(uint32_t shape, uint32_t mesh_id, const Ray& ray, const Trafo& trafo, Intersection& isec) const {  // shape.zig:165-179
        switch (shape) {
            case ZYG_SHAPE_TRIANGLE_MESH: {  // TriangleTree.intersect, triangle_tree.zig:46-109
                ZoHit h;
                if (treeOf(mesh_id).intersect(trafo.worldToObjectRay(ray), h, nullptr, nullptr)) {
                    isec.t         = h.t;
                    isec.u         = h.u;
                    isec.v         = h.v;
                    isec.primitive = h.primitive;
                    isec.trafo     = trafo;
                    return true;
                }
                return false;
            }
            case ZYG_SHAPE_CUBE: return cube::intersect(ray, trafo, isec);
            case ZYG_SHAPE_RECTANGLE: return rectangle::intersect(ray, trafo, isec);
            case ZYG_SHAPE_DISK: return disk::intersect(ray, trafo, isec);
            case ZYG_SHAPE_SPHERE: return sphere::intersect(ray, trafo, isec);
            case ZYG_SHAPE_DISTANT: return distant::intersect(ray, trafo, isec);
            case ZYG_SHAPE_CANOPY: return canopy::intersect(ray, trafo, isec);
            default: return false;
        }
    }
    bool shapeIntersectP(uint32_t shape, uint32_t mesh_id, const Ray& ray, const Trafo& trafo) const {  // shape.zig:221-233
        switch (shape) {
            case ZYG_SHAPE_TRIANGLE_MESH: return treeOf(mesh_id).intersectP(trafo.worldToObjectRay(ray));  // triangle_mesh.zig:337-340
            case ZYG_SHAPE_CUBE: return cube::intersectP(ray, trafo);
            case ZYG_SHAPE_RECTANGLE: return rectangle::intersectP(ray, trafo);
            case ZYG_SHAPE_DISK: return disk::intersectP(ray, trafo);
            case ZYG_SHAPE_SPHERE: return sphere::intersectP(ray, trafo);
            default: return false;
        }
    }
    void shapeFragment(uint32_t shape, uint32_t mesh_id, const Ray& ray, Fragment& frag) const {  // shape.zig:205-219
        switch (shape) {
            case ZYG_SHAPE_TRIANGLE_MESH: mesh::fragment(meshes[mesh_id], frag); break;
            case ZYG_SHAPE_CUBE: cube::fragment(ray, frag); break;
            case ZYG_SHAPE_RECTANGLE: rectangle::fragment(ray, frag); break;
            case ZYG_SHAPE_DISK: disk::fragment(ray, frag); break;
            case ZYG_SHAPE_SPHERE: sphere::fragment(ray, frag); break;
            case ZYG_SHAPE_DISTANT: distant::fragment(ray, frag); break;
            case ZYG_SHAPE_CANOPY: canopy::fragment(ray, frag); break;
            default: break;
        }
    }

    // Prop.intersect, prop.zig:163-197
    bool propIntersect(uint32_t entity, const Ray& ray, uint32_t depth_surface, Intersection& isec) const {
        const ZygpuProp& prop = s.props[entity];
        if (!visible(prop.flags, depth_surface)) return false;
        if (!propAabb(entity).intersect(ray)) return false;
        return shapeIntersect(prop.shape, prop.mesh, ray, propTrafo(entity), isec);
    }

    // Prop.visibility, prop.zig:199-237 (no masks): true = unoccluded
    bool propVisibility(uint32_t entity, const Ray& ray) const {
        const ZygpuProp& prop = s.props[entity];
        if (0 == (prop.flags & ZYG_PROP_VISIBLE_IN_SHADOW)) return true;
        if (!propAabb(entity).intersect(ray)) return true;
        return !shapeIntersectP(prop.shape, prop.mesh, ray, propTrafo(entity));
    }

    static float nodeIntersect(const ZygpuBvhNode& node, const Ray& ray) {  // node.zig:73-87
        const Node* n = reinterpret_cast<const Node*>(&node);
        return n->intersect(ray);
    }

    // PropBvh.intersect, prop_tree.zig:56-116
    bool intersect(Ray& ray, uint32_t depth_surface, Fragment& frag) const {
        const ZygpuPropTree& tree = s.solid_bvh;

        NodeStack stack;
        uint32_t  n = 0 == tree.num_nodes ? NodeStack::End : 0;

        Intersection isec{};
        uint32_t     prop = ZYGPU_NULL;

        while (NodeStack::End != n) {
            const ZygpuBvhNode& node = tree.nodes[n];

            const uint32_t num = node.num_indices;
            if (0 != num) {
                const uint32_t start = node.children_or_start;
                for (uint32_t i = start; i < start + num; ++i) {
                    const uint32_t p = tree.indices[i];
                    if (propIntersect(p, ray, depth_surface, isec)) {
                        ray.max_t = isec.t;
                        prop      = p;
                    }
                }
                n = stack.pop();
                continue;
            }

            uint32_t a = node.children_or_start;
            uint32_t b = a + 1;

            float dista = nodeIntersect(tree.nodes[a], ray);
            float distb = nodeIntersect(tree.nodes[b], ray);
            if (dista > distb) {
                std::swap(a, b);
                std::swap(dista, distb);
            }
            if (FLT_MAX == dista) {
                n = stack.pop();
            } else {
                n = a;
                if (FLT_MAX != distb) stack.push(b);
            }
        }

        const bool hit = ZYGPU_NULL != prop;
        if (hit) {
            frag.isec = isec;
            shapeFragment(s.props[prop].shape, s.props[prop].mesh, ray, frag);
        }
        frag.prop = prop;
        return hit;
    }

    // PropBvh.visibility, prop_tree.zig:185-240; Scene.visibility scene.zig:229-235 (no volume props)
    bool visibility(const Ray& ray) const {
        const ZygpuPropTree& tree = s.solid_bvh;

        NodeStack stack;
        uint32_t  n = 0 == tree.num_nodes ? NodeStack::End : 0;

        while (NodeStack::End != n) {
            const ZygpuBvhNode& node = tree.nodes[n];

            const uint32_t num = node.num_indices;
            if (0 != num) {
                const uint32_t start = node.children_or_start;
                for (uint32_t i = start; i < start + num; ++i) {
                    if (!propVisibility(tree.indices[i], ray)) return false;
                }
                n = stack.pop();
                continue;
            }

            uint32_t a = node.children_or_start;
            uint32_t b = a + 1;

            float dista = nodeIntersect(tree.nodes[a], ray);
            float distb = nodeIntersect(tree.nodes[b], ray);
            if (dista > distb) {
                std::swap(a, b);
                std::swap(dista, distb);
            }
            if (FLT_MAX == dista) {
                n = stack.pop();
            } else {
                n = a;
                if (FLT_MAX != distb) stack.push(b);
            }
        }
        return true;
    }

    // ---- lights ----

    uint32_t lightNumSamples(const ZygpuLight& l, float split_threshold) const {  // shape_sampler.zig:35-41
        if (split_threshold <= LowThreshold) return 1;
        return l.num_samples;
    }

    struct LightProperties {  // light.zig:25-30, scene.zig:664-674
        Vec4f sphere, cone;
        float power;
        bool  two_sided;
    };
    LightProperties lightProperties(uint32_t light_id) const {
        const ZygpuAabb& box = s.light_aabbs[light_id];
        const Vec4f      pos = splat(0.5f) * (load4(box.min) + load4(box.max));
        return {{{pos[0], pos[1], pos[2], box.max[3]}}, load4(s.light_cones + 4 * light_id), box.min[3],
                0 != s.lights[light_id].two_sided};
    }

    static float clampedCosSub(float cos_a, float cos_b, float sin_a, float sin_b) {  // light_tree.zig:217-220
        const float angle = std::fmaf(cos_a, cos_b, sin_a * sin_b);
        return cos_a > cos_b ? 1.f : angle;
    }
    static float clampedSinSub(float cos_a, float cos_b, float sin_a, float sin_b) {  // :222-225
        const float angle = std::fmaf(sin_a, cos_b, -sin_b * cos_a);
        return cos_a > cos_b ? 0.f : angle;
    }

    // light_tree.zig:173-215
    static float importance(Vec4f p, Vec4f n, Vec4f center, Vec4f cone, float radius, float power, bool two_sided,
                            bool total_sphere) {
        const Vec4f axis = p - center;
        const float l    = length3(axis);
        const Vec4f na   = axis / splat(l);
        const Vec4f da   = cone;

        const float sin_cu   = min(radius / l, 1.f);
        const float cos_cone = cone[3];
        const float cos_a    = safe::absDotC(da, na, two_sided);
        const float cos_n    = max(-dot3(n, na), 0.f);

        const Vec4f sa = {{sin_cu, cos_cone, cos_a, cos_n}};
        const Vec4f sb = max4(mulAdd(sa, -sa, splat(1.f)), splat(0.f));
        Vec4f       sr;
        for (int i = 0; i < 4; ++i) sr[i] = std::sqrt(sb[i]);

        const float cos_cu   = sr[0];
        const float sin_cone = sr[1];
        const float sin_a    = sr[2];
        const float sin_n    = sr[3];

        const float ta = clampedCosSub(cos_a, cos_cone, sin_a, sin_cone);
        const float tb = clampedSinSub(cos_a, cos_cone, sin_a, sin_cone);
        const float tc = clampedCosSub(ta, cos_cu, tb, sin_cu);
        const float tn = clampedCosSub(cos_n, cos_cu, sin_n, sin_cu);

        const float ra = total_sphere ? 1.f : tn;
        const float rb = max(tc, 0.f);

        const float clamped_dist = max(l, 0.5f * radius);
        const float rc           = power / (clamped_dist * clamped_dist);

        return max(ra * rb * rc, 0.f);
    }

    // The tree a traversal runs over: the scene's (lights = scene lights) or the PrimitiveTree of a mesh sampler (lights =
    // the emitting triangles of the part).
    struct TreeRef {
        const ZygpuLightNode*   nodes;
        const uint32_t*         node_middles;
        const uint32_t*         light_orders;
        const uint32_t*         light_mapping;
        Vec4f                   bounds_min, bounds_max;
        const ZygpuMeshSampler* sampler;  // null for the scene tree
    };
    TreeRef sceneTree() const {
        const ZygpuLightTree& t = s.light_tree;
        return {t.nodes, t.node_middles, t.light_orders, t.light_mapping, load4(t.bounds.min), load4(t.bounds.max), nullptr};
    }
    static TreeRef primitiveTree(const ZygpuMeshSampler& m) {
        return {m.nodes, m.node_middles, m.light_orders, m.light_mapping, load4(m.bounds.min), load4(m.bounds.max), &m};
    }

    // MeshImpl.lightProperties, shape_sampler.zig:198-226
    LightProperties meshLightProperties(const ZygpuMeshSampler& m, uint32_t light) const {
        const uint32_t global = m.triangle_mapping[light];
        const Mesh     tree   = treeOf(m.mesh);
        const Vec4f    a = tree.position(tree.triangles[3 * size_t(global)]), b = tree.position(tree.triangles[3 * size_t(global) + 1]),
                    c = tree.position(tree.triangles[3 * size_t(global) + 2]);

        const Vec4f center = (a + b + c) / splat(3.f);
        const float sra    = squaredLength3(a - center);
        const float srb    = squaredLength3(b - center);
        const float src    = squaredLength3(c - center);
        const float radius = std::sqrt(max(sra, max(srb, src)));
        const Vec4f nn     = normalize3(cross3(b - a, c - a));
        return {{{center[0], center[1], center[2], radius}}, {{nn[0], nn[1], nn[2], 1.f}}, m.triangle_pdfs[light], 0 != m.two_sided};
    }

    float lightWeight(const TreeRef& tr, Vec4f p, Vec4f n, bool total_sphere, uint32_t light) const {  // light_tree.zig:227-233
        const LightProperties props = tr.sampler ? meshLightProperties(*tr.sampler, light) : lightProperties(light);
        return importance(p, n, props.sphere, props.cone, props.sphere[3], props.power, props.two_sided, total_sphere);
    }

    struct LNode {  // light_tree.Node over the flattened record
        const ZygpuLightNode& r;
        bool                  hasChildren() const { return 0 != (r.meta & 1u); }
        bool                  twoSided() const { return 0 != (r.meta & 2u); }
        uint32_t              childrenOrLight() const { return r.meta >> 2; }
    };

    static Vec4f unorm16ToFloat(const uint16_t v[4]) {  // encoding.zig:49-58
        const float k = 1.f / 65535.f;
        return {{float(v[0]) * k, float(v[1]) * k, float(v[2]) * k, float(v[3]) * k}};
    }
    static Vec4f snorm16ToFloat(const uint16_t v[4]) {  // encoding.zig:71-80
        Vec4f r;
        for (int i = 0; i < 4; ++i) r[i] = std::fmaf(float(v[i]), 1.f / 32768.f, -1.f);
        return r;
    }

    static Vec4f nodeCenter(const TreeRef& tr, const ZygpuLightNode& node) {  // light_tree.zig:39-42
        const Vec4f t = unorm16ToFloat(node.center);
        return lerp(tr.bounds_min, tr.bounds_max, t);
    }
    float nodeWeight(const TreeRef& tr, const ZygpuLightNode& node, Vec4f p, Vec4f n, bool total_sphere) const {  // :57-63
        const Vec4f center = nodeCenter(tr, node);
        const Vec4f cone   = snorm16ToFloat(node.cone);
        return importance(p, n, center, cone, center[3], node.power, 0 != (node.meta & 2u), total_sphere);
    }
    bool nodeSplit(const TreeRef& tr, const ZygpuLightNode& node, Vec4f p, float threshold) const {  // :65-89
        const Vec4f center = nodeCenter(tr, node);
        const float r      = center[3];
        const float d      = min(distance3(p, center), 1.0e6f);
        const float a      = max(d - r, 0.001f);
        const float b      = d + r;

        const float eg  = 1.f / (a * b);
        const float eg2 = eg * eg;
        const float a3  = a * a * a;
        const float b3  = b * b * b;
        const float e2g = (b3 - a3) / (3.f * (b - a) * a3 * b3);
        const float vg  = e2g - eg2;

        const float ve = node.variance;
        const float ee = node.power;
        const float s2 = max(ve * vg + ve * eg2 + ee * ee * vg, 0.f);
        const float ns = 1.f / (1.f + std::sqrt(s2));
        return ns < threshold;
    }

    // Node.randomLight, light_tree.zig:91-145
    LightPick nodeRandomLight(const TreeRef& tr, const ZygpuLightNode& node, Vec4f p, Vec4f n, bool total_sphere, float random) const {
        const uint32_t* light_mapping = tr.light_mapping;
        const uint32_t  num_lights    = node.num_lights;
        const uint32_t  light         = node.meta >> 2;

        if (1 == num_lights) return {light_mapping[light], 1.f};

        uint32_t front = light;
        uint32_t back  = light + num_lights - 1;

        float w_front = lightWeight(tr, p, n, total_sphere, light_mapping[front]);
        float w_back  = lightWeight(tr, p, n, total_sphere, light_mapping[back]);

        float w_sum_front = w_front;
        float w_sum_back  = w_back;
        float w_sum       = 0.f;

        while (front != back) {
            w_sum = w_sum_front + w_sum_back;
            if (w_sum_front <= random * w_sum) {
                front += 1;
                if (front != back) {
                    w_front = lightWeight(tr, p, n, total_sphere, light_mapping[front]);
                    w_sum_front += w_front;
                } else {
                    w_front = w_back;
                }
            } else {
                back -= 1;
                if (front != back) {
                    w_back = lightWeight(tr, p, n, total_sphere, light_mapping[back]);
                    w_sum_back += w_back;
                }
            }
        }
        if (0.f == w_sum) return {0, 0.f};
        return {light_mapping[front], w_front / w_sum};
    }

    // Node.pdf, light_tree.zig:147-170
    float nodePdf(const TreeRef& tr, const ZygpuLightNode& node, Vec4f p, Vec4f n, bool total_sphere, uint32_t id) const {
        const uint32_t num_lights = node.num_lights;
        if (1 == num_lights) return 1.f;

        const uint32_t light = node.meta >> 2;
        const uint32_t end   = light + num_lights;

        float w_id = 0.f;
        float sum  = 0.f;
        for (uint32_t i = light; i < end; ++i) {
            const float lw = lightWeight(tr, p, n, total_sphere, tr.light_mapping[i]);
            sum += lw;
            if (id == i) w_id = lw;
        }
        if (0.f == sum) return 0.f;
        return w_id / sum;
    }

    // Tree.randomLight, light_tree.zig:346-447. Infinite lights go through a Distribution1D the scenes in scope
    // never populate with more than MaxLights - 1 entries, so only the split_infinite branch and the empty
    // distribution are restated.
    uint32_t randomLight(Vec4f p, Vec4f n, bool total_sphere, float random, float split_threshold, LightPick* buffer) const {
        const ZygpuLightTree& tree = s.light_tree;
        const TreeRef         tr   = sceneTree();

        uint32_t current_light = 0;

        float          ip                  = 0.f;
        const uint32_t num_infinite_lights = tree.num_infinite_lights;
        const bool     split               = split_threshold > 0.f;

        if (split && num_infinite_lights < 64 - 1) {
            for (uint32_t i = 0; i < num_infinite_lights; ++i) buffer[current_light++] = {tree.light_mapping[i], 1.f};
        } else {
            ip = tree.infinite_weight;
            if (random < tree.infinite_guard) {
                uint32_t l_offset;
                float    l_pdf;
                infinite_light_distribution.sampleDiscrete(random, l_offset, l_pdf);
                buffer[0] = {tree.light_mapping[l_offset], l_pdf * ip};
                return 1;
            }
        }

        if (0 == tree.num_nodes) return current_light;

        const float    pd              = 1.f - ip;
        const uint32_t max_split_depth = tree.max_split_depth;

        struct Value {
            float    pdf, random;
            uint32_t node, depth;
        };
        Value    stack[12];  // TraversalStackT(MaxSplitDepth)
        uint32_t end = 0;

        Value t{pd, (random - ip) / pd, 0, split ? 0 : max_split_depth};
        stack[end++] = t;

        auto pop = [&]() {
            end -= 1;
            return stack[end];
        };

        while (end > 0) {
            const ZygpuLightNode& node = tree.nodes[t.node];

            if (0 != (node.meta & 1u)) {
                const bool do_split = t.depth < max_split_depth && nodeSplit(tr, node, p, split_threshold);

                const uint32_t c0 = node.meta >> 2;
                const uint32_t c1 = c0 + 1;

                if (do_split) {
                    t.depth += 1;
                    t.node       = c0;
                    stack[end++] = {t.pdf, t.random, c1, t.depth};
                } else {
                    t.depth = max_split_depth;

                    float p0 = nodeWeight(tr, tree.nodes[c0], p, n, total_sphere);
                    float p1 = nodeWeight(tr, tree.nodes[c1], p, n, total_sphere);

                    const float pt = p0 + p1;
                    if (0.f == pt) {
                        t = pop();
                        continue;
                    }

                    p0 /= pt;
                    p1 /= pt;

                    if (t.random < p0) {
                        t.node = c0;
                        t.pdf *= p0;
                        t.random /= p0;
                    } else {
                        t.node = c1;
                        t.pdf *= p1;
                        t.random = min((t.random - p0) / p1, 1.f);
                    }
                }
            } else {
                const LightPick pick = nodeRandomLight(tr, node, p, n, total_sphere, t.random);
                if (pick.pdf > 0.f) buffer[current_light++] = {pick.offset, pick.pdf * t.pdf};
                t = pop();
            }
        }
        return current_light;
    }

    // Tree.pdf, light_tree.zig:449-517
    float lightTreePdf(Vec4f p, Vec4f n, bool total_sphere, float split_threshold, uint32_t id) const {
        const ZygpuLightTree& tree = s.light_tree;
        const TreeRef         tr   = sceneTree();

        const uint32_t lo                  = tree.light_orders[id];
        const uint32_t num_infinite_lights = tree.num_infinite_lights;
        const bool     split               = split_threshold > 0.f;
        const bool     split_infinite      = split && num_infinite_lights < 64 - 1;

        if (lo < tree.infinite_end) {
            if (split_infinite) return 1.f;
            return tree.infinite_weight * infinite_light_distribution.pdfI(lo);
        }
        if (0 == tree.num_nodes) return 0.f;

        const float    ip              = split_infinite ? 0.f : tree.infinite_weight;
        const uint32_t max_split_depth = tree.max_split_depth;

        float    pd    = 1.f - ip;
        uint32_t nid   = 0;
        uint32_t depth = split ? 0 : max_split_depth;
        for (;;) {
            const ZygpuLightNode& node = tree.nodes[nid];
            if (0 != (node.meta & 1u)) {
                const bool     do_split = depth < max_split_depth && nodeSplit(tr, node, p, split_threshold);
                const uint32_t c0       = node.meta >> 2;
                const uint32_t c1       = c0 + 1;
                const uint32_t middle   = tree.node_middles[nid];
                if (do_split) {
                    depth += 1;
                    nid = lo < middle ? c0 : c1;
                } else {
                    depth          = max_split_depth;
                    const float p0 = nodeWeight(tr, tree.nodes[c0], p, n, total_sphere);
                    const float p1 = nodeWeight(tr, tree.nodes[c1], p, n, total_sphere);
                    const float pt = p0 + p1;
                    if (0.f == pt) return 0.f;
                    if (lo < middle) {
                        nid = c0;
                        pd *= p0 / pt;
                    } else {
                        nid = c1;
                        pd *= p1 / pt;
                    }
                }
            } else {
                return pd * nodePdf(tr, node, p, n, total_sphere, lo);
            }
        }
    }

    // Light.sampleTo -> Shape.sampleTo, light.zig:87-106, 163-190; shape.zig:301-338
    // PrimitiveTree.randomLight, light_tree.zig:577-650
    uint32_t primitiveRandomLight(const ZygpuMeshSampler& m, Vec4f p, Vec4f n, bool total_sphere, float random, float split_threshold,
                                  LightPick* buffer) const {
        constexpr uint32_t MaxSplitDepth = 6;
        const TreeRef      tr            = primitiveTree(m);

        uint32_t   current_light = 0;
        const bool split         = split_threshold > 0.f;

        struct Value {
            float    pdf, random;
            uint32_t node, depth;
        };
        Value    stack[MaxSplitDepth + 1];
        uint32_t end = 0;

        Value t{1.f, random, 0, split ? 0 : MaxSplitDepth};
        stack[end++] = t;

        auto pop = [&]() {
            end -= 1;
            return stack[end];
        };

        while (end > 0) {
            const ZygpuLightNode& node = tr.nodes[t.node];
            if (0 != (node.meta & 1u)) {
                const bool     do_split = t.depth < MaxSplitDepth && nodeSplit(tr, node, p, split_threshold);
                const uint32_t c0       = node.meta >> 2;
                const uint32_t c1       = c0 + 1;
                if (do_split) {
                    t.depth += 1;
                    t.node       = c0;
                    stack[end++] = {t.pdf, t.random, c1, t.depth};
                } else {
                    t.depth = MaxSplitDepth;

                    float p0 = nodeWeight(tr, tr.nodes[c0], p, n, total_sphere);
                    float p1 = nodeWeight(tr, tr.nodes[c1], p, n, total_sphere);

                    const float pt = p0 + p1;
                    if (0.f == pt) {
                        t = pop();
                        continue;
                    }
                    p0 /= pt;
                    p1 /= pt;
                    if (t.random < p0) {
                        t.node = c0;
                        t.pdf *= p0;
                        t.random /= p0;
                    } else {
                        t.node = c1;
                        t.pdf *= p1;
                        t.random = min((t.random - p0) / p1, 1.f);
                    }
                }
            } else {
                const LightPick pick = nodeRandomLight(tr, node, p, n, total_sphere, t.random);
                if (pick.pdf > 0.f) buffer[current_light++] = {pick.offset, pick.pdf * t.pdf};
                t = pop();
            }
        }
        return current_light;
    }

    // PrimitiveTree.pdf, light_tree.zig:652-719
    float primitivePdf(const ZygpuMeshSampler& m, Vec4f p, Vec4f n, bool total_sphere, float split_threshold, uint32_t id) const {
        constexpr uint32_t MaxSplitDepth = 6;
        const TreeRef      tr            = primitiveTree(m);

        const uint32_t lo    = tr.light_orders[id];
        const bool     split = split_threshold > 0.f;

        float    pd    = 1.f;
        uint32_t nid   = 0;
        uint32_t depth = split ? 0 : MaxSplitDepth;
        for (;;) {
            const ZygpuLightNode& node = tr.nodes[nid];
            if (0 != (node.meta & 1u)) {
                const bool     do_split = depth < MaxSplitDepth && nodeSplit(tr, node, p, split_threshold);
                const uint32_t c0       = node.meta >> 2;
                const uint32_t c1       = c0 + 1;
                const uint32_t middle   = tr.node_middles[nid];
                if (do_split) {
                    depth += 1;
                    nid = lo < middle ? c0 : c1;
                } else {
                    depth          = MaxSplitDepth;
                    const float p0 = nodeWeight(tr, tr.nodes[c0], p, n, total_sphere);
                    const float p1 = nodeWeight(tr, tr.nodes[c1], p, n, total_sphere);
                    const float pt = p0 + p1;
                    if (0.f == pt) return 0.f;
                    if (lo < middle) {
                        nid = c0;
                        pd *= p0 / pt;
                    } else {
                        nid = c1;
                        pd *= p1 / pt;
                    }
                }
            } else {
                return pd * nodePdf(tr, node, p, n, total_sphere, lo);
            }
        }
    }

    uint32_t lightSampleTo(const ZygpuLight& l, Vec4f p, Vec4f n, const Trafo& trafo, bool total_sphere,
                           float split_threshold, Sampler& sampler, SampleTo* buffer) const {
        const uint32_t num_samples = lightNumSamples(l, split_threshold);
        switch (s.props[l.prop].shape) {
            case ZYG_SHAPE_RECTANGLE:
                if (ZYG_LIGHT_PROP_IMAGE == l.light_class) {
                    return rectangle::sampleMaterialTo(p, n, trafo, 0 != l.two_sided, total_sphere, num_samples, image_samplers[l.sampler],
                                                       sampler, buffer);
                }
                return rectangle::sampleTo(p, n, trafo, 0 != l.two_sided, total_sphere, num_samples, sampler, buffer);
            case ZYG_SHAPE_DISTANT: return distant::sampleTo(n, trafo, total_sphere, sampler, buffer);
            case ZYG_SHAPE_SPHERE: return sphere::sampleTo(p, n, trafo, total_sphere, num_samples, sampler, buffer);
            case ZYG_SHAPE_DISK: return disk::sampleTo(p, n, trafo, 0 != l.two_sided, total_sphere, num_samples, sampler, buffer);
            case ZYG_SHAPE_CANOPY:  // Light.propSampleMaterialTo -> Shape.sampleMaterialTo, light.zig:191-215, shape.zig:348-371
                if (ZYG_LIGHT_PROP_IMAGE != l.light_class) return 0;  // Canopy.sampleTo (uniform sky) is not in scope
                return canopy::sampleMaterialTo(n, trafo, total_sphere, image_samplers[l.sampler], sampler, buffer);
            case ZYG_SHAPE_TRIANGLE_MESH:
                return meshSampleTo(s.mesh_samplers[l.sampler], p, n, trafo, 0 != l.two_sided, total_sphere, split_threshold, sampler, buffer);
            default: return 0;
        }
    }

    // Mesh.sampleTo, triangle_mesh.zig:492-608
    uint32_t meshSampleTo(const ZygpuMeshSampler& m, Vec4f p, Vec4f n, const Trafo& trafo, bool two_sided, bool total_sphere,
                          float split_threshold, Sampler& sampler, SampleTo* buffer) const {
        const Vec4f op = trafo.worldToObjectPoint(p);
        const Vec4f on = trafo.worldToObjectNormal(n);

        const Vec4f scale         = trafo.scale();
        const Vec4f scale_squared = scale * scale;

        const Mesh tree = treeOf(m.mesh);

        LightPick      samples[64];
        const uint32_t num = primitiveRandomLight(m, op, on, total_sphere, sampler.sample1D(), split_threshold, samples);

        uint32_t current_sample = 0;
        for (uint32_t i = 0; i < num; ++i) {
            const LightPick& sp     = samples[i];
            const uint32_t   global = m.triangle_mapping[sp.offset];

            const Vec4f a = tree.position(tree.triangles[3 * size_t(global)]);
            const Vec4f b = tree.position(tree.triangles[3 * size_t(global) + 1]);
            const Vec4f c = tree.position(tree.triangles[3 * size_t(global) + 2]);

            const Vec4f e1 = b - a;
            const Vec4f e2 = c - a;

            const Vec4f cross_axis = cross3(e1, e2);

            const Vec4f ca  = scale_squared * cross_axis;
            const float lca = length3(ca);
            const Vec4f sn  = ca / splat(lca);
            Vec4f       wn  = trafo.objectToWorldNormal(sn);

            const float tri_area = 0.5f * lca;

            const Vec4f center = (a + b + c) / splat(3.f);

            Vec4f dir, v;
            float bary_uv[2];
            float sample_pdf, n_dot_dir;

            if (tri_area / distance3(center, op) > mesh::AreaDistanceRatio) {
                const Vec2f           r2 = sampler.sample2D();
                mesh::SphericalSample sample;
                if (!mesh::sampleSpherical(op, a, b, c, r2.v, sample)) continue;
                if (dot3(sample.dir, on) <= 0.f && !total_sphere) continue;

                bary_uv[0] = sample.uv[0];
                bary_uv[1] = sample.uv[1];

                dir = trafo.objectToWorldNormal(sample.dir);

                const Vec4f sv = mesh::interpolate3(a, b, c, bary_uv[0], bary_uv[1]);
                v              = trafo.objectToWorldPoint(sv);
                sample_pdf     = sp.pdf * sample.pdf;

                if (two_sided && dot3(wn, dir) > 0.f) wn = -wn;
                n_dot_dir = -dot3(wn, dir);
            } else {
                const Vec2f r2 = sampler.sample2D();
                mesh::triangleUniform(r2.v, bary_uv);

                const Vec4f sv = mesh::interpolate3(a, b, c, bary_uv[0], bary_uv[1]);
                v              = trafo.objectToWorldPoint(sv);

                const Vec4f axis = v - p;
                const float sl   = squaredLength3(axis);
                const float d    = std::sqrt(sl);
                dir              = axis / splat(d);

                if (dot3(dir, n) <= 0.f && !total_sphere) continue;
                if (two_sided && dot3(wn, dir) > 0.f) wn = -wn;

                n_dot_dir  = -dot3(wn, dir);
                sample_pdf = (sp.pdf * sl) / (n_dot_dir * tri_area);
            }

            if (n_dot_dir < safe::DotMin) continue;

            SampleTo& out = buffer[current_sample++];
            out.p         = {{v[0], v[1], v[2], sample_pdf}};
            out.n         = wn;
            out.wi        = dir;
            out.uvw       = splat(0.f);  // interpolated uv: only read by emission maps
        }
        return current_sample;
    }

    // Mesh.pdf, triangle_mesh.zig:662-703
    float meshPdf(const ZygpuMeshSampler& m, Vec4f dir, Vec4f p, Vec4f n, const Fragment& frag, bool total_sphere, float split_threshold) const {
        const float n_dot_dir = std::fabs(dot3(frag.geo_n, dir));

        const Vec4f op = frag.isec.trafo.worldToObjectPoint(p);
        const Vec4f on = frag.isec.trafo.worldToObjectNormal(n);

        const uint32_t pm      = m.primitive_mapping[frag.isec.primitive];
        const float    tri_pdf = primitivePdf(m, op, on, total_sphere, split_threshold, pm);

        const Mesh  tree = treeOf(m.mesh);
        const Vec4f a    = tree.position(tree.triangles[3 * size_t(frag.isec.primitive)]);
        const Vec4f b    = tree.position(tree.triangles[3 * size_t(frag.isec.primitive) + 1]);
        const Vec4f c    = tree.position(tree.triangles[3 * size_t(frag.isec.primitive) + 2]);

        const Vec4f cross_axis = cross3(b - a, c - a);
        const Vec4f scale      = frag.isec.trafo.scale();
        const Vec4f ca         = (scale * scale) * cross_axis;
        const float tri_area   = 0.5f * length3(ca);

        const Vec4f center = (a + b + c) / splat(3.f);

        if (tri_area / distance3(center, op) > mesh::AreaDistanceRatio) return tri_pdf * mesh::pdfSpherical(op, a, b, c);
        const float sl = squaredDistance3(p, frag.p);
        return (tri_pdf * sl) / (n_dot_dir * tri_area);
    }

    bool lightFinite(const ZygpuLight& l) const {  // Shape.finite, shape.zig:94-99
        const uint32_t shape = s.props[l.prop].shape;
        return !(ZYG_SHAPE_CANOPY == shape || ZYG_SHAPE_DISTANT == shape || ZYG_SHAPE_DOME == shape);
    }

    // Shape.shadowRay, shape.zig:401-416
    static Ray shadowRay(Vec4f origin, const SampleTo& sample, bool finite = true) {
        if (!finite) return Ray::init(origin, sample.wi, 0.f, RayMaxT);
        const Vec4f light_pos   = offsetRay(sample.p, sample.n);
        const Vec4f shadow_axis = light_pos - origin;
        const float shadow_len  = length3(shadow_axis);
        return Ray::init(origin, shadow_axis / splat(shadow_len), 0.f, shadow_len);
    }

    // Material.evaluateRadiance, material.zig:194-207 for {Light, Substitute (uncoated)}
    Vec4f materialRadiance(const ZygpuMaterial& m, Vec4f wi, const Trafo& trafo, uint32_t prop, bool in_camera, uint32_t part,
                           Vec4f uvw, float stochastic_r) const {
        if (ZYGPU_NULL != m.emission_map) {  // Emittance.radiance with an image emission_map, emittance.zig:29-59
            if (-dot3(wi, trafo.r[2]) < m.emission_cos_a) return splat(0.f);
            const float factor    = in_camera ? m.emission_camera_weight : 1.f;
            const Vec4f intensity = load4(m.emission) * image_samplers[m.emission_map].texel(uvw[0], uvw[1], stochastic_r);
            if (0.f != m.emission_normalize) {
                const Vec4f scale = trafo.scale();
                return splat(factor / shapeArea(s.props[prop].shape, scale)) * intensity;
            }
            return splat(factor) * intensity;
        }
        float area = 1.f;
        if (0.f != m.emission_normalize) {
            const Vec4f scale = trafo.scale();
            area = ZYG_SHAPE_TRIANGLE_MESH == s.props[prop].shape
                       ? s.mesh_part_areas[s.props[prop].parts_start + part] * (scale[0] * scale[1])  // Mesh.area, triangle_mesh.zig:283-286
                       : shapeArea(s.props[prop].shape, scale);
        }
        return emittanceRadiance(m, wi, trafo, area, in_camera);
    }
};

// ---- integrator ------------------------------------------------------------------------------

bool g_wavefront_light_order = false;  // zo_set_wavefront_light_order

// aov.Value, rendering/sensor/aov/aov_value.zig:9-82 (Emission / Direct / Indirect need no slot: they are the IValue)
struct AovValue {
    uint32_t slots = 0;
    Vec4f    values[6];

    bool active() const { return 0 != slots; }
    bool activeClass(uint32_t c) const { return 0 != (slots & (1u << c)); }
    void clear() {
        for (uint32_t c = 0; c < 6; ++c) {
            if (activeClass(c)) values[c] = ZYG_AOV_DEPTH == c ? splat(std::numeric_limits<float>::max()) : splat(0.f);
        }
    }
    void insert3(uint32_t c, Vec4f v) { values[c] = v; }
    void insert1(uint32_t c, float v) { values[c][0] = v; }
};

struct Worker {
    const Scene& scene;
    Generator    rng;
    Sampler      samplers[2];
    AovValue     aov;
    bool         debug = false;  // ZO_DEBUG_PIXEL="x,y": print the vertices of that pixel's paths (diagnostics)

    explicit Worker(const Scene& sc) : scene(sc) {
        samplers[0].is_sobol = ZYG_SAMPLER_SOBOL == sc.view.sampler;
        samplers[0].rng      = &rng;
        samplers[1].is_sobol = false;
        samplers[1].rng      = &rng;
    }

    Sampler& pickSampler(uint32_t bounce) { return bounce < 3 ? samplers[0] : samplers[1]; }  // worker.zig:201-207

    // Scene.lightPdf, scene.zig:624-634
    float lightPdf(const Vertex& vertex, const Fragment& frag) const {
        const uint32_t light_id = scene.propLightId(frag.prop, frag.part);
        if (vertex.state.singular || ZYGPU_NULL == light_id) return 1.f;

        const float select_pdf = scene.lightTreePdf(vertex.origin, vertex.geo_n, vertex.state.translucent,
                                                    vertex.light_split_threshold, light_id);

        // Light.pdf -> Shape.pdf, light.zig:149-157, shape.zig:469-492
        const ZygpuLight& l          = scene.s.lights[light_id];
        float             sample_pdf = 0.f;
        switch (scene.s.props[l.prop].shape) {
            case ZYG_SHAPE_RECTANGLE:
                sample_pdf = ZYG_LIGHT_PROP_IMAGE == l.light_class
                                 ? rectangle::materialPdf(vertex.ray.direction, vertex.origin, frag,
                                                          scene.lightNumSamples(l, vertex.light_split_threshold), scene.image_samplers[l.sampler])
                                 : rectangle::pdf(vertex.origin, frag, scene.lightNumSamples(l, vertex.light_split_threshold));
                break;
            case ZYG_SHAPE_DISK:
                sample_pdf = disk::pdf(vertex.ray.direction, vertex.origin, frag, scene.lightNumSamples(l, vertex.light_split_threshold));
                break;
            case ZYG_SHAPE_DISTANT: sample_pdf = 1.f / distant::solidAngle(frag.isec.trafo.scaleX()); break;  // distant.zig:139-141
            case ZYG_SHAPE_SPHERE:
                sample_pdf = sphere::pdf(vertex.origin, frag.isec.trafo, scene.lightNumSamples(l, vertex.light_split_threshold));
                break;
            case ZYG_SHAPE_CANOPY:  // Light.propMaterialPdf -> Shape.materialPdf, light.zig:371-374, shape.zig:519
                if (ZYG_LIGHT_PROP_IMAGE == l.light_class) sample_pdf = scene.image_samplers[l.sampler].pdf(frag.uvw[0], frag.uvw[1]) / (2.f * kPi);
                break;
            case ZYG_SHAPE_TRIANGLE_MESH:
                sample_pdf = scene.meshPdf(scene.s.mesh_samplers[l.sampler], vertex.ray.direction, vertex.origin, vertex.geo_n, frag,
                                           vertex.state.translucent, vertex.light_split_threshold);
                break;
            default: break;
        }
        return powerHeuristic(vertex.bxdf_pdf, sample_pdf * select_pdf);
    }

    // Vertex.evaluateRadiance, vertex.zig:183-212
    Vec4f evaluateRadiance(const Vertex& vertex, const Fragment& frag, Sampler& sampler) const {
        const Vec4f          wo = -vertex.ray.direction;
        const ZygpuMaterial& m  = scene.propMaterial(frag.prop, frag.part);
        if (0 == (m.flags & ZYG_MATERIAL_EMISSIVE) || (0 == (m.flags & ZYG_MATERIAL_TWO_SIDED) && !frag.sameHemisphere(wo))) {
            return splat(0.f);
        }

        const float stochastic_r = sampler.sample1D();  // rs.stochastic_r

        const bool  in_camera = 0 == vertex.probe_depth.total();
        const Vec4f energy    = scene.materialRadiance(m, wo, frag.isec.trafo, frag.prop, in_camera, frag.part, frag.uvw, stochastic_r);
        const float weight    = lightPdf(vertex, frag);
        return splat(weight) * energy;
    }

    // Context.emission -> PropBvh.emission -> Prop.emission -> Shape.emission, prop_tree.zig:302-356,
    // prop.zig:239-264, rectangle.zig:188-196
    Vec4f emission(const Vertex& vertex, Sampler& sampler) const {
        const ZygpuPropTree& tree = scene.s.unoccluding_bvh;

        NodeStack stack;
        uint32_t  n = 0 == tree.num_nodes ? NodeStack::End : 0;

        Vec4f energy = splat(0.f);

        while (NodeStack::End != n) {
            const ZygpuBvhNode& node = tree.nodes[n];
            const uint32_t      num  = node.num_indices;
            if (0 != num) {
                const uint32_t start = node.children_or_start;
                for (uint32_t i = start; i < start + num; ++i) energy = energy + propEmission(tree.indices[i], vertex, sampler);
                n = stack.pop();
                continue;
            }

            uint32_t a = node.children_or_start;
            uint32_t b = a + 1;

            float dista = Scene::nodeIntersect(tree.nodes[a], vertex.ray);
            float distb = Scene::nodeIntersect(tree.nodes[b], vertex.ray);
            if (dista > distb) {
                std::swap(a, b);
                std::swap(dista, distb);
            }
            if (FLT_MAX == dista) {
                n = stack.pop();
            } else {
                n = a;
                if (FLT_MAX != distb) stack.push(b);
            }
        }
        return energy;
    }

    Vec4f propEmission(uint32_t entity, const Vertex& vertex, Sampler& sampler) const {
        const ZygpuProp& prop = scene.s.props[entity];
        if (!Scene::visible(prop.flags, vertex.probe_depth.surface)) return splat(0.f);
        if (!scene.propAabb(entity).intersect(vertex.ray)) return splat(0.f);

        Fragment frag;
        frag.isec.trafo = scene.propTrafo(entity);
        frag.prop       = entity;

        switch (prop.shape) {
            case ZYG_SHAPE_RECTANGLE:
                if (!rectangle::intersect(vertex.ray, frag.isec.trafo, frag.isec)) return splat(0.f);
                rectangle::fragment(vertex.ray, frag);
                return evaluateRadiance(vertex, frag, sampler);
            case ZYG_SHAPE_DISK:  // Disk.emission, disk.zig:171-179
                if (!disk::intersect(vertex.ray, frag.isec.trafo, frag.isec)) return splat(0.f);
                disk::fragment(vertex.ray, frag);
                return evaluateRadiance(vertex, frag, sampler);
            case ZYG_SHAPE_TRIANGLE_MESH: {  // Mesh.emission -> Tree.emission, triangle_mesh.zig:379-388, triangle_tree.zig:405-477
                const Ray    local_ray = frag.isec.trafo.worldToObjectRay(vertex.ray);
                const Mesh   tree      = scene.treeOf(prop.mesh);
                const ZoMesh& zm       = scene.meshes[prop.mesh];
                Vec4f        energy    = splat(0.f);
                tree.allHits(local_ray, [&](uint32_t i, const Hit& hit) {
                    frag.isec.t         = hit.t;
                    frag.isec.u         = hit.u;
                    frag.isec.v         = hit.v;
                    frag.isec.primitive = i;
                    frag.part           = zm.parts[i];
                    const Vec4f a = tree.position(tree.triangles[3 * size_t(i)]), b = tree.position(tree.triangles[3 * size_t(i) + 1]),
                                c = tree.position(tree.triangles[3 * size_t(i) + 2]);
                    frag.p     = frag.isec.trafo.objectToWorldPoint(mesh::interpolate3(a, b, c, hit.u, hit.v));
                    frag.geo_n = frag.isec.trafo.objectToWorldNormal(normalize3(cross3(b - a, c - a)));
                    frag.uvw   = splat(0.f);  // interpolated uv: only read by emission maps
                    energy     = energy + evaluateRadiance(vertex, frag, sampler);
                });
                return energy;
            }
            case ZYG_SHAPE_SPHERE:  // Sphere.emission, sphere.zig:271-279
                if (!sphere::intersect(vertex.ray, frag.isec.trafo, frag.isec)) return splat(0.f);
                sphere::fragment(vertex.ray, frag);
                return evaluateRadiance(vertex, frag, sampler);
            default: return splat(0.f);  // shape.zig:283-299
        }
    }

    // PathtracerMIS.connectLight, pathtracer_mis.zig:280-341 (no shadow catchers; infinite props evaluated on escape)
    Vec4f connectLight(Vertex& vertex, const Fragment& frag, Sampler& sampler) const {
        if (0 == scene.view.caustics_path && vertex.state.specular && !vertex.state.primary_ray) return splat(0.f);

        vertex.light_split_threshold = splitThreshold(scene.view.split_threshold, vertex.depth);

        Vec4f result = splat(0.f);
        if (frag.hit()) result = evaluateRadiance(vertex, frag, sampler);

        result = result + emission(vertex, sampler);

        if (RayMaxT == vertex.ray.max_t) {  // :313-338
            for (uint32_t i = 0; i < scene.s.num_infinite_props; ++i) {
                const uint32_t prop = scene.s.infinite_props[i];
                Fragment       light_frag;
                if (!scene.propIntersect(prop, vertex.ray, vertex.probe_depth.surface, light_frag.isec)) continue;
                light_frag.prop = prop;
                scene.shapeFragment(scene.s.props[prop].shape, scene.s.props[prop].mesh, vertex.ray, light_frag);
                result = result + evaluateRadiance(vertex, light_frag, sampler);
            }
        }
        return result;
    }

    // hlp.sampleNormal, material_helper.zig:16-79 (tex_coord UV)
    static Vec4f sampleNormal(Vec4f wo, const Renderstate& rs, float nx, float ny) {
        const float nmz = std::sqrt(max(1.f - (nx * nx + ny * ny), 0.01f));
        const Vec4f nm  = {{nx, ny, nmz, 0.f}};
        // rs.tangentToWorld(nm), renderstate.zig:51-58
        const Vec4f w = {{nm[0] * rs.t[0] + nm[1] * rs.b[0] + nm[2] * rs.n[0], nm[0] * rs.t[1] + nm[1] * rs.b[1] + nm[2] * rs.n[1],
                          nm[0] * rs.t[2] + nm[1] * rs.b[2] + nm[2] * rs.n[2], 0.f}};
        const Vec4f n = normalize3(w);

        const Vec4f ng = rs.geo_n;
        const Vec4f r  = reflect3(n, wo);
        const float a  = dot3(ng, r);
        if (a >= 0.f) return n;

        const float cos_threshold = 0.0017453f;  // cos(89.9 degrees)
        if (dot3(ng, wo) < cos_threshold) return ng;

        const float b       = dot3(ng, n);
        const float epsilon = 1e-4f;
        Vec4f       tangent;
        if (b > epsilon) {
            const float distance_to_surface_along_normal = std::fabs(a) / b;
            tangent = normalize3(r + splat(distance_to_surface_along_normal) * n);
        } else {
            tangent = n;
        }
        tangent = tangent + splat(epsilon) * ng;
        return normalize3(wo + tangent);
    }

    // Vertex.sample, vertex.zig:137-181 + Material.sample, material.zig:184-194
    MaterialSample vertexSample(const Vertex& vertex, const Fragment& frag, Sampler& sampler, bool caustics) const {
        const Vec4f          wo = -vertex.ray.direction;
        const ZygpuMaterial& m  = scene.propMaterial(frag.prop, frag.part);

        Renderstate rs;
        rs.trafo = frag.isec.trafo;
        rs.p     = frag.p;
        rs.t     = frag.t;
        rs.b     = frag.b;
        if (0 != (m.flags & ZYG_MATERIAL_TWO_SIDED) && !frag.sameHemisphere(wo)) {
            rs.geo_n = -frag.geo_n;
            rs.n     = -frag.n;
        } else {
            rs.geo_n = frag.geo_n;
            rs.n     = frag.n;
        }
        rs.origin           = vertex.origin;
        rs.uvw              = frag.uvw;
        rs.stochastic_r     = sampler.sample1D();
        rs.ior = frag.sameHemisphere(wo) ? vertex.mediums.topIor() : vertex.mediums.peekIor(frag.prop, frag.part);  // Vertex.iorOutside
        rs.reg_weight       = scene.view.regularize_roughness;
        rs.reg_alpha        = vertex.reg_alpha;
        rs.prop             = frag.prop;
        rs.part             = frag.part;
        rs.primary          = vertex.state.primary_ray;
        rs.caustics         = caustics;
        rs.highest_priority = vertex.mediums.highestPriority();

        switch (m.type) {
            case ZYG_MATERIAL_SUBSTITUTE: {
                if (ZYGPU_NULL == m.color_map && ZYGPU_NULL == m.roughness_map && ZYGPU_NULL == m.metallic_map && ZYGPU_NULL == m.normal_map) {
                    return substituteSample(m, wo, rs, scene.view.specular_threshold, scene.luts);
                }
                ZygpuMaterial textured = m;
                auto texel = [&](uint32_t map) { return scene.image_samplers[map].texel(rs.uvw[0], rs.uvw[1], rs.stochastic_r); };
                if (ZYGPU_NULL != m.color_map) {  // ts.sample2D_3(self.color, rs, ...), substitute_material.zig:120
                    const Vec4f c = texel(m.color_map);
                    textured.color[0] = c[0], textured.color[1] = c[1], textured.color[2] = c[2];
                }
                if (ZYGPU_NULL != m.roughness_map) textured.roughness = texel(m.roughness_map)[0];  // ts.sample2D_1(self.roughness), :122
                if (ZYGPU_NULL != m.metallic_map) textured.metallic = texel(m.metallic_map)[0];     // :123
                MaterialSample result = substituteSample(textured, wo, rs, scene.view.specular_threshold, scene.luts);
                if (ZYGPU_NULL != m.normal_map) {  // :157-159
                    const Vec4f xy = texel(m.normal_map);
                    const Vec4f n  = sampleNormal(wo, rs, xy[0], xy[1]);
                    Vec4f       t, b;
                    orthonormalBasis3(n, t, b);
                    result.super.frame = {t, b, n};
                }
                return result;
            }
            case ZYG_MATERIAL_GLASS: return glassSample(m, wo, rs, scene.view.specular_threshold, scene.luts);
            default: return lightSample(wo, rs);
        }
    }

    // PathtracerMIS.evaluateLight, pathtracer_mis.zig:214-278
    Vec4f evaluateLight(const LightPick& light_pick, const Vertex& vertex, const Fragment& frag, const MaterialSample& mat_sample,
                        uint32_t max_material_splits, Sampler& sampler) const {
        const Vec4f p           = frag.p;
        const Vec4f gn          = mat_sample.super.geo_n;
        const bool  translucent = mat_sample.isTranslucent();

        const ZygpuLight& light = scene.s.lights[light_pick.offset];
        const Trafo       trafo = scene.propTrafo(light.prop);

        Vec4f occluded = splat(0.f);

        SampleTo       samples[64];
        const uint32_t num = scene.lightSampleTo(light, p, gn, trafo, translucent, vertex.light_split_threshold, sampler, samples);

        for (uint32_t i = 0; i < num; ++i) {
            const SampleTo& light_sample = samples[i];

            const Ray shadow = Scene::shadowRay(frag.offsetP(light_sample.wi), light_sample, scene.lightFinite(light));
            if (!scene.visibility(shadow)) continue;

            // Light.evaluateTo, light.zig:119-132
            const float          stochastic_r = sampler.sample1D();
            const ZygpuMaterial& lm       = scene.propMaterial(light.prop, light.part);
            const Vec4f          radiance = scene.materialRadiance(lm, light_sample.wi, trafo, light.prop, false, light.part, light_sample.uvw, stochastic_r);

            const bxdf::Result bxdf_result = mat_sample.evaluate(light_sample.wi, max_material_splits, false);

            const float light_pdf = light_sample.pdf() * light_pick.pdf;
            const float weight    = predividedPowerHeuristic(light_pdf, bxdf_result.pdf);

            occluded = occluded + splat(weight) * radiance * bxdf_result.reflection;
        }
        return occluded;
    }

    // PathtracerMIS.sampleLights, pathtracer_mis.zig:174-212
    Vec4f sampleLights(const Vertex& vertex, const Fragment& frag, const MaterialSample& mat_sample, uint32_t max_material_splits,
                       Sampler& sampler) const {
        Vec4f result = splat(0.f);
        if (!mat_sample.canEvaluate()) return result;

        const Vec4f p           = frag.p;
        const Vec4f n           = mat_sample.super.geo_n;
        const bool  translucent = mat_sample.isTranslucent();

        const float select = sampler.sample1D();

        LightPick      lights[64];
        const uint32_t num = scene.randomLight(p, n, translucent, select, vertex.light_split_threshold, lights);
        if (g_wavefront_light_order) return sampleLightsWavefront(lights, num, vertex, frag, mat_sample, max_material_splits, sampler);
        for (uint32_t i = 0; i < num; ++i) {
            result = result + evaluateLight(lights[i], vertex, frag, mat_sample, max_material_splits, sampler);
        }
        return result;
    }

    // The same estimator with the sampler draws regrouped the way a wavefront has to take them: first the light samples of
    // every pick (Light.sampleTo), then - after all shadow rays - one draw per visible sample (Light.evaluateTo, light.zig:127).
    // The reference interleaves the two per light sample; with one pick and one sample both orders coincide. The device
    // implements this order; tests compare it with this variant per pixel and the two variants with each other statistically.
    Vec4f sampleLightsWavefront(const LightPick* lights, uint32_t num, const Vertex& vertex, const Fragment& frag,
                                const MaterialSample& mat_sample, uint32_t max_material_splits, Sampler& sampler) const {
        const Vec4f p           = frag.p;
        const Vec4f gn          = mat_sample.super.geo_n;
        const bool  translucent = mat_sample.isTranslucent();

        struct Record {
            SampleTo sample;
            uint32_t light;
            float    pick_pdf;
        };
        std::vector<Record> records;
        for (uint32_t i = 0; i < num; ++i) {
            const ZygpuLight& light = scene.s.lights[lights[i].offset];
            const Trafo       trafo = scene.propTrafo(light.prop);
            SampleTo          samples[64];
            const uint32_t    n = scene.lightSampleTo(light, p, gn, trafo, translucent, vertex.light_split_threshold, sampler, samples);
            for (uint32_t k = 0; k < n; ++k) records.push_back({samples[k], lights[i].offset, lights[i].pdf});
        }

        // the device adds the records in order (with one sample per pick this is the reference's summation order too)
        Vec4f result = splat(0.f);
        for (const Record& r : records) {
            const ZygpuLight& light = scene.s.lights[r.light];
            const Trafo       trafo = scene.propTrafo(light.prop);

            const Ray shadow = Scene::shadowRay(frag.offsetP(r.sample.wi), r.sample, scene.lightFinite(light));
            if (!scene.visibility(shadow)) continue;

            const float          stochastic_r = sampler.sample1D();  // Light.evaluateTo
            const ZygpuMaterial& lm       = scene.propMaterial(light.prop, light.part);
            const Vec4f          radiance = scene.materialRadiance(lm, r.sample.wi, trafo, light.prop, false, light.part, r.sample.uvw, stochastic_r);

            const bxdf::Result bxdf_result = mat_sample.evaluate(r.sample.wi, max_material_splits, false);

            const float light_pdf = r.sample.pdf() * r.pick_pdf;
            const float weight    = predividedPowerHeuristic(light_pdf, bxdf_result.pdf);

            result = result + splat(weight) * radiance * bxdf_result.reflection;
        }
        return result;
    }

    // Worker.commonAOV, worker.zig:209-242
    void commonAOV(const Vertex& vertex, const Fragment& frag, const MaterialSample& mat_sample) {
        if (vertex.state.primary_ray && aov.activeClass(ZYG_AOV_ALBEDO) && mat_sample.canEvaluate()) {
            // MaterialSample.aovAlbedo, material_sample.zig:38-45; Substitute: substitute_sample.zig:80-86 (no volumetric materials)
            Vec4f albedo = splat(0.f);
            if (MaterialSample::Substitute == mat_sample.kind) {
                albedo = lerp(mat_sample.albedo, mat_sample.f0, splat(mat_sample.metallic));
            } else if (MaterialSample::Glass == mat_sample.kind) {
                albedo = splat(1.f);
            }
            aov.insert3(ZYG_AOV_ALBEDO, vertex.throughput * albedo);
        }
        if (vertex.probe_depth.surface > 0) return;
        if (aov.activeClass(ZYG_AOV_GEOMETRIC_NORMAL)) aov.insert3(ZYG_AOV_GEOMETRIC_NORMAL, mat_sample.super.geo_n);
        if (aov.activeClass(ZYG_AOV_SHADING_NORMAL)) aov.insert3(ZYG_AOV_SHADING_NORMAL, mat_sample.super.frame.z);
        if (aov.activeClass(ZYG_AOV_ROUGHNESS)) aov.insert1(ZYG_AOV_ROUGHNESS, std::sqrt(mat_sample.super.alpha[0]));
        if (aov.activeClass(ZYG_AOV_DEPTH)) aov.insert1(ZYG_AOV_DEPTH, vertex.ray.max_t);
        if (aov.activeClass(ZYG_AOV_MATERIAL_ID)) aov.insert1(ZYG_AOV_MATERIAL_ID, float(1u + scene.propMaterialId(frag.prop, frag.part)));
    }

    static uint32_t maxSplits(const Vertex& v, uint32_t depth) {  // vertex.zig:306-309
        const uint32_t m = 4 / v.path_count;
        return m - (v.state.primary_ray ? 0 : std::min(depth, m - 1));
    }

    // PathtracerMIS.li, pathtracer_mis.zig:37-172. The uniform-parameter Substitute never returns more than one
    // bxdf sample (substitute_sample.zig:147-234), so the pool degenerates into a loop over one vertex.
    IValue li(Vertex vertex) {
        IValue result;

        std::vector<Vertex> current{vertex}, next;

        // Pool.transparency, vertex.zig:243-268 (no shadow-catcher props): a vertex that ended without a successor adds what a
        // see-through path lost on the way, or its whole weight
        float transparency = 0.f;

        while (!current.empty()) {
            next.clear();
            std::vector<bool> ended;
            for (Vertex& v : current) {
                const size_t before = next.size();
                step(v, result, next);
                ended.push_back(next.size() == before);
            }
            for (size_t i = 0; i < current.size(); ++i) {  // summed when the generation has been consumed, in buffer order
                if (!ended[i]) continue;
                const Vertex& v = current[i];
                transparency += v.state.transparent ? max((1.f - average3(v.throughput)) * v.split_weight, 0.f) : v.split_weight;
            }
            current.swap(next);
        }
        result.direct[3] = 0.f;
        result.direct[3] += transparency;  // result.direct += vertices.transparency, pathtracer_mis.zig:169-170
        return result;
    }

    void step(Vertex& vertex, IValue& result, std::vector<Vertex>& next) {
        const uint32_t max_depth_surface = scene.view.max_depth_surface;
        const uint32_t max_depth_volume  = scene.view.max_depth_volume;

        const uint32_t total_depth = vertex.probe_depth.total();
        Sampler&       sampler     = pickSampler(total_depth);

        // Context.nextEvent, context.zig:54-69 (no volume props)
        Fragment frag;
        if (!vertex.mediums.empty()) {
            // VolumeIntegrator.integrate, volume_integrator.zig:84-130, for a homogeneous, non-scattering medium (the
            // "glass" case of propScatter, :51-66): clip the ray to the medium prop's box, intersect, absorb
            const MediumStack::Medium& medium    = vertex.mediums.top();
            const float                ray_max_t = vertex.ray.max_t;
            const float                limit     = cube::aabbIntersectP(scene.propAabb(medium.prop), vertex.ray);
            vertex.ray.max_t                     = min(offsetF(limit), ray_max_t);
            scene.intersect(vertex.ray, vertex.probe_depth.surface, frag);
            if (frag.hit()) {
                const float d     = vertex.ray.max_t;
                const Vec4f cc_a  = vertex.mediums.topCC();
                const Vec4f x     = splat(-(d - vertex.ray.min_t)) * cc_a;
                const Vec4f tr    = {{std::exp(x[0]), std::exp(x[1]), std::exp(x[2]), std::exp(x[3])}};  // attenuation3
                vertex.throughput = vertex.throughput * tr;
            }
        } else {
            const Vec4f origin = vertex.ray.origin;
            scene.intersect(vertex.ray, vertex.probe_depth.surface, frag);
            const float dif_t = distance3(origin, vertex.ray.origin);
            vertex.ray.origin = origin;
            vertex.ray.max_t += dif_t;
        }

        if (debug) {
            std::printf("[zo] depth %u pc %u sw %g state p%d s%d sg%d | hit prop %u t %.9g media %u | thr %.9g %.9g %.9g | rng %llx\n", total_depth,
                        vertex.path_count, vertex.split_weight, vertex.state.primary_ray, vertex.state.specular, vertex.state.singular,
                        frag.prop, vertex.ray.max_t, vertex.mediums.index, vertex.throughput[0], vertex.throughput[1], vertex.throughput[2],
                        (unsigned long long)rng.state);
        }

        const Vec4f this_light       = connectLight(vertex, frag, sampler);
        const Vec4f split_weight     = splat(vertex.split_weight);
        Vec4f       split_throughput = vertex.throughput * split_weight;

        result.add(split_throughput * this_light, total_depth, 2, 0 == total_depth, vertex.state.singular);

        if (!frag.hit() || vertex.probe_depth.surface >= max_depth_surface || vertex.probe_depth.volume >= max_depth_volume) {
            if (vertex.state.transparent) {  // pathtracer_mis.zig:81-83
                vertex.throughput = vertex.throughput * (splat(1.f) - Vec4f{{min(this_light[0], 1.f), min(this_light[1], 1.f), min(this_light[2], 1.f), min(this_light[3], 1.f)}});
            }
            return;
        }

        if (russianRoulette(vertex.throughput, sampler.sample1D())) return;

        const bool           caustics   = !vertex.state.primary_ray ? 0 != scene.view.caustics_path : true;
        const MaterialSample mat_sample = vertexSample(vertex, frag, sampler, caustics);

        if (aov.active()) commonAOV(vertex, frag, mat_sample);  // pathtracer_mis.zig:95-97

        split_throughput = vertex.throughput * split_weight;

        vertex.light_split_threshold = splitThreshold(scene.view.split_threshold, vertex.probe_depth);
        const uint32_t max_splits    = maxSplits(vertex, total_depth);
        const Vec4f    next_light    = sampleLights(vertex, frag, mat_sample, max_splits, sampler);

        result.add(split_throughput * next_light, total_depth, 1, false, false);

        bxdf::Sample   bxdf_samples[4];
        const uint32_t path_count = mat_sample.sample(sampler, max_splits, bxdf_samples);

        if (0 == path_count) vertex.throughput = splat(0.f);
        if (debug) {
            for (uint32_t i = 0; i < path_count; ++i) {
                std::printf("[zo]   sample %u/%u event %d sw %g pdf %g wi %.9g %.9g %.9g\n", i, path_count, int(bxdf_samples[i].path.event),
                            bxdf_samples[i].split_weight, bxdf_samples[i].pdf, bxdf_samples[i].wi[0], bxdf_samples[i].wi[1], bxdf_samples[i].wi[2]);
            }
        }

        for (uint32_t i = 0; i < path_count; ++i) {
            const bxdf::Sample& sample_result = bxdf_samples[i];

            Vertex next_vertex       = vertex;
            next_vertex.path_count   = vertex.path_count * path_count;
            next_vertex.split_weight = vertex.split_weight * sample_result.split_weight;

            const bxdf::Path path = sample_result.path;
            next_vertex.state.update(path);

            if (bxdf::Event::Straight != path.event) {
                next_vertex.state.translucent = mat_sample.isTranslucent();
                next_vertex.depth             = next_vertex.probe_depth;
                next_vertex.bxdf_pdf          = sample_result.pdf;
                next_vertex.origin            = frag.p;
                next_vertex.geo_n             = mat_sample.super.geo_n;
                next_vertex.reg_alpha         = path.reg_alpha;
            }

            next_vertex.throughput = next_vertex.throughput * (sample_result.reflection / splat(sample_result.pdf));

            next_vertex.ray = frag.offsetRayTo(sample_result.wi);
            next_vertex.probe_depth.surface += 1;  // Probe.Depth.increment, no subsurface

            if (bxdf::Event::Transmission == path.event) {  // Vertex.interfaceChange, vertex.zig:95-110
                if (frag.sameHemisphere(sample_result.wi)) {
                    next_vertex.mediums.remove(frag.prop, frag.part);
                } else {
                    const ZygpuMaterial& material = scene.propMaterial(frag.prop, frag.part);
                    next_vertex.mediums.push(frag.prop, frag.part, load4(material.color), material.ior, int8_t(material.priority));
                }
            }

            next_vertex.state.transparent =
                next_vertex.state.transparent && (bxdf::Event::Transmission == path.event || bxdf::Event::Straight == path.event);

            next.push_back(next_vertex);
        }

        sampler.incrementPadding();
    }
};

// ---- sensor ----------------------------------------------------------------------------------

inline Vec4f clampColor(Vec4f color, float mx) {  // sensor.zig:615-624
    const float mc = hmax3(color);
    if (mc > mx) {
        const float r = mx / mc;
        return splat(r) * color;
    }
    return color;
}

struct Film {
    float*           pixels;  // Pack4f per pixel, weight in w
    const ZygpuView& view;
    float* const*    aov_layers = nullptr;  // aov.Buffer: ZYG_AOV_NUM_CLASSES Pack4f images, null where a class is inactive
    float*           alpha      = nullptr;  // Transparent buffer (buffer_transparent.zig): the fourth lane of `pixels`, sum of weight * alpha

    float eval(float s) const {  // sensor.zig:626-628 + InterpolatedFunction1DN.eval
        const float    x      = std::fabs(s);
        const float    cx     = min(x, view.filter_range_end);
        const float    o      = cx * view.filter_inverse_interval;
        const uint32_t offset = uint32_t(o);
        const float    t      = o - float(offset);
        return lerp(view.filter[offset], view.filter[std::min(offset + 1, 29u)], t);
    }

    void addPixel(uint32_t i, Vec4f color, float weight, bool atomic) const {  // buffer_opaque.zig:39-55
        const Vec4f wc = splat(weight) * color;
        float*      v  = pixels + size_t(i) * 4;
        if (atomic) {
            std::atomic_ref<float>(v[0]).fetch_add(wc[0], std::memory_order_relaxed);
            std::atomic_ref<float>(v[1]).fetch_add(wc[1], std::memory_order_relaxed);
            std::atomic_ref<float>(v[2]).fetch_add(wc[2], std::memory_order_relaxed);
            std::atomic_ref<float>(v[3]).fetch_add(weight, std::memory_order_relaxed);
            if (alpha) std::atomic_ref<float>(alpha[i]).fetch_add(wc[3], std::memory_order_relaxed);  // Transparent.addPixelAtomic
        } else {
            v[0] += wc[0];
            v[1] += wc[1];
            v[2] += wc[2];
            v[3] += weight;
            if (alpha) alpha[i] += wc[3];  // Transparent.addPixel, buffer_transparent.zig:45-54
        }
    }

    void add(int32_t px, int32_t py, float weight, Vec4f color, const int32_t bounds[4], const int32_t isolated[4]) const {  // :559-574
        if (uint32_t(px - bounds[0]) <= uint32_t(bounds[2]) && uint32_t(py - bounds[1]) <= uint32_t(bounds[3])) {
            const uint32_t i    = uint32_t(view.resolution[0] * py + px);
            const bool     iso  = uint32_t(px - isolated[0]) <= uint32_t(isolated[2]) && uint32_t(py - isolated[1]) <= uint32_t(isolated[3]);
            addPixel(i, color, weight, !iso);
        }
    }

    // aov.Buffer.addPixel / addPixelAtomic / lessPixel / overwritePixel, aov_buffer.zig:84-113
    void addAovPixel(uint32_t i, uint32_t c, Vec4f value, float weight, bool atomic) const {
        float* v = aov_layers[c] + size_t(i) * 4;
        if (atomic) {
            std::atomic_ref<float>(v[0]).fetch_add(weight * value[0], std::memory_order_relaxed);
            std::atomic_ref<float>(v[1]).fetch_add(weight * value[1], std::memory_order_relaxed);
            std::atomic_ref<float>(v[2]).fetch_add(weight * value[2], std::memory_order_relaxed);
            std::atomic_ref<float>(v[3]).fetch_add(weight, std::memory_order_relaxed);
        } else {
            const Vec4f wc = splat(weight) * value;
            v[0] += wc[0];
            v[1] += wc[1];
            v[2] += wc[2];
            v[3] += weight;
        }
    }
    // Sensor.addAov / lessAov / overwriteAov, sensor.zig:576-613
    void addAov(int32_t px, int32_t py, uint32_t c, float weight, Vec4f value, const int32_t bounds[4], const int32_t isolated[4]) const {
        if (uint32_t(px - bounds[0]) <= uint32_t(bounds[2]) && uint32_t(py - bounds[1]) <= uint32_t(bounds[3])) {
            const uint32_t i   = uint32_t(view.resolution[0] * py + px);
            const bool     iso = uint32_t(px - isolated[0]) <= uint32_t(isolated[2]) && uint32_t(py - isolated[1]) <= uint32_t(isolated[3]);
            addAovPixel(i, c, value, weight, !iso);
        }
    }
    void lessAov(int32_t px, int32_t py, uint32_t c, float value, const int32_t bounds[4]) const {
        if (uint32_t(px - bounds[0]) <= uint32_t(bounds[2]) && uint32_t(py - bounds[1]) <= uint32_t(bounds[3])) {
            float* v = aov_layers[c] + size_t(view.resolution[0] * py + px) * 4;
            if (value < v[0]) v[0] = value;
        }
    }
    void overwriteAov(int32_t px, int32_t py, uint32_t c, float weight, float value, const int32_t bounds[4]) const {
        if (uint32_t(px - bounds[0]) <= uint32_t(bounds[2]) && uint32_t(py - bounds[1]) <= uint32_t(bounds[3])) {
            float* v = aov_layers[c] + size_t(view.resolution[0] * py + px) * 4;
            if (weight > v[3]) {
                v[0] = value;
                v[3] = weight;
            }
        }
    }

    // the AOV half of Sensor.addSample, sensor.zig:197-219, 244-275, 328-377
    void addAovSample(int32_t x, int32_t y, const float* wx, const float* wy, Vec4f emission, Vec4f direct, Vec4f indirect,
                      const AovValue& aov, const int32_t bounds[4], const int32_t isolated[4]) const {
        const int32_t r = view.filter_radius_int;
        for (uint32_t c = 0; c < ZYG_AOV_NUM_CLASSES; ++c) {
            if (!aov.activeClass(c) || !aov_layers || !aov_layers[c]) continue;
            const Vec4f avalue = ZYG_AOV_EMISSION == c ? emission : (ZYG_AOV_DIRECT == c ? direct : (ZYG_AOV_INDIRECT == c ? indirect : aov.values[c]));
            if (0 == r) {
                const uint32_t id = uint32_t(view.resolution[0] * y + x);
                float*         v  = aov_layers[c] + size_t(id) * 4;
                if (ZYG_AOV_DEPTH == c) {
                    if (avalue[0] < v[0]) v[0] = avalue[0];
                } else if (ZYG_AOV_MATERIAL_ID == c) {
                    if (1.f > v[3]) {
                        v[0] = avalue[0];
                        v[3] = 1.f;
                    }
                } else {
                    addAovPixel(id, c, avalue, 1.f, false);
                }
            } else if (ZYG_AOV_DEPTH == c) {
                lessAov(x, y, c, avalue[0], bounds);
            } else if (ZYG_AOV_MATERIAL_ID == c) {
                overwriteAov(x, y, c, wx[r] * wy[r], avalue[0], bounds);
            } else {
                for (int32_t j = 0; j <= 2 * r; ++j) {
                    for (int32_t i = 0; i <= 2 * r; ++i) addAov(x - r + i, y - r + j, c, wx[i] * wy[j], avalue, bounds, isolated);
                }
            }
        }
    }

    // Sensor.addSample, sensor.zig:168-385
    void addSample(int32_t x, int32_t y, const float pixel_uv[2], const IValue& value, const AovValue& aov, const int32_t bounds[4],
                   const int32_t isolated[4]) const {
        const Vec4f emission = clampColor(value.emission, view.clamp_emission);
        const Vec4f direct   = clampColor(value.direct, view.clamp_direct);
        const Vec4f indirect = clampColor(value.indirect, view.clamp_indirect);
        const Vec4f summed   = emission + direct + indirect;
        const Vec4f composed = {{summed[0], summed[1], summed[2], value.direct[3]}};

        const float ox = pixel_uv[0] - 0.5f;
        const float oy = pixel_uv[1] - 0.5f;

        const int32_t r = view.filter_radius_int;
        if (0 == r) {
            addPixel(uint32_t(view.resolution[0] * y + x), composed, 1.f, false);
            if (aov.active()) addAovSample(x, y, nullptr, nullptr, emission, direct, indirect, aov, bounds, isolated);
            return;
        }
        // r = 1: weights eval(o + 1), eval(o), eval(o - 1); r = 2: eval(o + 2) ... eval(o - 2); rows outer, columns inner.
        float wx[5], wy[5];
        for (int32_t i = 0; i <= 2 * r; ++i) {
            wx[i] = eval(ox + float(r - i));
            wy[i] = eval(oy + float(r - i));
        }
        for (int32_t j = 0; j <= 2 * r; ++j) {
            for (int32_t i = 0; i <= 2 * r; ++i) add(x - r + i, y - r + j, wx[i] * wy[j], composed, bounds, isolated);
        }
        if (aov.active()) addAovSample(x, y, wx, wy, emission, direct, indirect, aov, bounds, isolated);
    }
};

// Perspective.generateVertex, camera_perspective.zig:124-150 (static camera: the shutter time only costs a draw)
Vertex generateVertex(const ZygpuView& view, int32_t px, int32_t py, const float pixel_uv[2], const float lens_uv[2]) {
    const float c0 = float(px) + pixel_uv[0];
    const float c1 = float(py) + pixel_uv[1];

    Vec4f direction = load4(view.left_top) + load4(view.d_x) * splat(c0) + load4(view.d_y) * splat(c1);
    Vec4f origin;

    if (view.aperture_radius > 0.f) {
        float lens[2];
        diskConcentric(lens_uv, lens);  // Aperture.sample, aperture.zig:46-53 (no shape texture)
        origin            = {{lens[0] * view.aperture_radius, lens[1] * view.aperture_radius, 0.f, 0.f}};
        const Vec4f t     = splat(view.focus_distance / direction[2]);
        const Vec4f focus = t * direction;
        direction         = focus - origin;
    } else {
        origin = load4(view.eye_offset);
    }

    const Trafo trafo       = Trafo::load(view.camera_trafo);
    const Vec4f origin_w    = trafo.objectToWorldPoint(origin);
    const Vec4f direction_w = trafo.objectToWorldVector(normalize3(direction));

    Vertex v;
    v.ray    = Ray::init(origin_w, direction_w, 0.f, RayMaxT);
    v.origin = origin_w;  // Vertex.init, vertex.zig:67-85
    return v;
}

// Worker.render, worker.zig:104-168, for one tile
void renderTile(Worker& worker, const Film& film, const int32_t tile[4], uint32_t iteration, uint32_t num_samples,
                uint32_t num_expected_samples) {
    const ZygpuView& view = worker.scene.view;

    const int32_t crop[4]     = {view.crop[0], view.crop[1], view.crop[2] - (view.crop[0] + 1), view.crop[3] - (view.crop[1] + 1)};
    const int32_t fr          = view.filter_radius_int;
    const int32_t isolated[4] = {tile[0] + fr, tile[1] + fr, (tile[2] - fr) - (tile[0] + fr), (tile[3] - fr) - (tile[1] + fr)};

    const int32_t  r0 = view.resolution[0] + 2 * fr;
    const int32_t  r1 = view.resolution[1] + 2 * fr;
    const uint32_t a  = uint32_t(r0) * uint32_t(r1);
    const uint64_t o  = uint64_t(iteration) * a;
    const uint32_t so = iteration / num_expected_samples;

    for (int32_t y = tile[1]; y <= tile[3]; ++y) {
        const uint32_t pixel_n = uint32_t((y + fr) * r0);
        for (int32_t x = tile[0]; x <= tile[2]; ++x) {
            const uint32_t pixel_id = pixel_n + uint32_t(x + fr);

            worker.rng.start(0, uint64_t(pixel_id) + o);
            {
                const char* dbg = std::getenv("ZO_DEBUG_PIXEL");
                int                dx = -1, dy = -1;
                if (dbg) std::sscanf(dbg, "%d,%d", &dx, &dy);
                worker.debug = dbg && x == dx && y == dy;
                if (worker.debug) std::printf("[zo] pixel %d %d iteration %u\n", x, y, iteration);
            }

            const uint64_t sample_index = uint64_t(pixel_id) * uint64_t(num_expected_samples) + uint64_t(iteration);
            const uint32_t tsi          = uint32_t(sample_index);
            const uint32_t seed         = uint32_t(sample_index >> 32) + so;

            worker.samplers[0].startPixel(tsi, seed);

            for (uint32_t s = 0; s < num_samples; ++s) {
                // Sensor.cameraSample, sensor.zig:152-166
                const Vec4f s4 = worker.samplers[0].sample4D();
                (void)worker.samplers[0].sample1D();  // time
                worker.samplers[0].incrementPadding();

                const float pixel_uv[2] = {s4[0], s4[1]};
                const float lens_uv[2]  = {s4[2], s4[3]};

                worker.aov.clear();  // worker.zig:155

                const Vertex vertex = generateVertex(view, x, y, pixel_uv, lens_uv);
                const IValue ivalue = worker.li(vertex);

                film.addSample(x, y, pixel_uv, ivalue, worker.aov, crop, isolated);

                worker.samplers[0].incrementSample();
            }
        }
    }
}

}  // namespace

}  // namespace zo

extern "C" {

// Driver.renderFrameIterationForward over the whole crop (driver.zig:338-348): samples [iteration, iteration +
// num_samples) of every pixel are added to `film` (Pack4f per pixel of the full resolution, weight in w; not cleared).
// per_sample_iterations != 0 renders the range as num_samples calls of (iteration + k, 1) — the progressive API's
// schedule (capi.zig:602-609), which reseeds the PCG stream per sample (worker.zig:143) and is what the device does.
void zo_render(const ZygpuScene* scene, const ZygpuView* view, const ZoMesh* meshes, uint32_t iteration, uint32_t num_samples,
               int per_sample_iterations, float* film_pixels, uint32_t threads) {
    zo_render_layers(scene, view, meshes, iteration, num_samples, per_sample_iterations, film_pixels, nullptr, nullptr, threads);
}
void zo_render_aov(const ZygpuScene* scene, const ZygpuView* view, const ZoMesh* meshes, uint32_t iteration, uint32_t num_samples,
                   int per_sample_iterations, float* film_pixels, float* const* aov_layers, uint32_t threads) {
    zo_render_layers(scene, view, meshes, iteration, num_samples, per_sample_iterations, film_pixels, aov_layers, nullptr, threads);
}

// zo_render that also fills the AOV layers of view->aov_slots: aov_layers[c] = Pack4f image of class c, cleared by the caller to the
// class default (aov.Buffer.clear: Depth floatMax in xyz, 0 elsewhere, weight 0); entries of inactive classes may be null.
// `alpha` (one float per pixel, not cleared, may be null): the alpha lane of the Transparent sensor buffer, sum of weight * alpha.
void zo_render_layers(const ZygpuScene* scene, const ZygpuView* view, const ZoMesh* meshes, uint32_t iteration, uint32_t num_samples,
                      int per_sample_iterations, float* film_pixels, float* const* aov_layers, float* alpha, uint32_t threads) {
    using namespace zo;

    const Scene sc(*scene, *view, meshes);
    const Film  film{film_pixels, *view, aov_layers, alpha};
    const uint32_t aov_slots = aov_layers ? view->aov_slots : 0u;

    const int32_t fr   = view->filter_radius_int;
    const int32_t td   = 32;  // Worker.TileDimensions
    const int32_t ntx  = (view->crop[2] - view->crop[0] + td - 1) / td;
    const int32_t nty  = (view->crop[3] - view->crop[1] + td - 1) / td;
    const int32_t nt   = ntx * nty;

    if (0 == threads) threads = std::max(1u, std::thread::hardware_concurrency());

    const uint32_t passes        = per_sample_iterations ? num_samples : 1;
    const uint32_t pass_samples  = per_sample_iterations ? 1 : num_samples;

    for (uint32_t pass = 0; pass < passes; ++pass) {
        std::atomic<int32_t> current{0};

        auto work = [&]() {
            Worker worker(sc);
            worker.aov.slots = aov_slots;
            for (;;) {
                const int32_t c = current.fetch_add(1, std::memory_order_relaxed);
                if (c >= nt) return;
                // TileQueue.pop, tile_queue.zig:47-87 (row-major instead of the generalised Hilbert order: the order
                // only decides which thread renders a tile)
                int32_t start[2] = {(c % ntx) * td + view->crop[0], (c / ntx) * td + view->crop[1]};
                int32_t end[2]   = {std::min(start[0] + td, view->crop[2]), std::min(start[1] + td, view->crop[3])};
                if (fr > 0) {
                    if (view->crop[1] == start[1]) start[1] -= fr;
                    if (view->crop[3] == end[1]) end[1] += fr;
                    if (view->crop[0] == start[0]) start[0] -= fr;
                    if (view->crop[2] == end[0]) end[0] += fr;
                }
                const int32_t tile[4] = {start[0], start[1], end[0] - 1, end[1] - 1};
                renderTile(worker, film, tile, iteration + pass, pass_samples, view->spp_total);
            }
        };

        if (threads <= 1) {
            work();
        } else {
            std::vector<std::thread> pool;
            for (uint32_t t = 0; t < threads; ++t) pool.emplace_back(work);
            for (auto& t : pool) t.join();
        }
    }
}

// Tree.randomLight / Tree.pdf (light_tree.zig:346-517) over the scene's light tree, for the host-side tests of the builder:
// picks[2 * i] = light id, picks[2 * i + 1] = pdf. Returns the number of picks.
uint32_t zo_light_tree_random(const ZygpuScene* scene, const ZygpuView* view, const float p[3], const float n[3], int total_sphere,
                              float random, float split_threshold, float* picks) {
    const zo::Scene sc(*scene, *view, nullptr);
    zo::LightPick   buffer[64];
    const uint32_t  num = sc.randomLight({{p[0], p[1], p[2], 0.f}}, {{n[0], n[1], n[2], 0.f}}, 0 != total_sphere, random, split_threshold, buffer);
    for (uint32_t i = 0; i < num; ++i) {
        picks[2 * i]     = float(buffer[i].offset);
        picks[2 * i + 1] = buffer[i].pdf;
    }
    return num;
}
float zo_light_tree_pdf(const ZygpuScene* scene, const ZygpuView* view, const float p[3], const float n[3], int total_sphere,
                        float split_threshold, uint32_t light) {
    const zo::Scene sc(*scene, *view, nullptr);
    return sc.lightTreePdf({{p[0], p[1], p[2], 0.f}}, {{n[0], n[1], n[2], 0.f}}, 0 != total_sphere, split_threshold, light);
}

// ImageImpl.sample / pdf / ts.sample2D_3 of image sampler `index` of the compiled scene (shape_sampler.zig:128-152)
void zo_image_sample(const ZygpuScene* scene, uint32_t index, uint32_t n, const float* r2, float* uv_pdf) {
    const zo::image::Sampler2D sampler(scene->image_samplers[index]);
    for (uint32_t i = 0; i < n; ++i) {
        float uv[2], pdf;
        sampler.sample(r2[2 * i], r2[2 * i + 1], uv, pdf);
        uv_pdf[3 * i] = uv[0], uv_pdf[3 * i + 1] = uv[1], uv_pdf[3 * i + 2] = pdf;
    }
}
void zo_image_pdf(const ZygpuScene* scene, uint32_t index, uint32_t n, const float* uv, float* pdf) {
    const zo::image::Sampler2D sampler(scene->image_samplers[index]);
    for (uint32_t i = 0; i < n; ++i) pdf[i] = sampler.pdf(uv[2 * i], uv[2 * i + 1]);
}
void zo_image_texel(const ZygpuScene* scene, uint32_t index, uint32_t n, const float* uvr, float* rgb) {
    const zo::image::Sampler2D sampler(scene->image_samplers[index]);
    for (uint32_t i = 0; i < n; ++i) {
        const zo::Vec4f c = sampler.texel(uvr[3 * i], uvr[3 * i + 1], uvr[3 * i + 2]);
        rgb[3 * i] = c[0], rgb[3 * i + 1] = c[1], rgb[3 * i + 2] = c[2];
    }
}

void zo_set_wavefront_light_order(int on) { zo::g_wavefront_light_order = 0 != on; }

// Transparent.resolveTonemap, buffer_transparent.zig:82-93
void zo_resolve_transparent(const ZygpuView* view, const float* film_pixels, const float* alpha, uint32_t num_pixels, float* rgba) {
    zo_resolve(view, film_pixels, num_pixels, rgba);
    for (uint32_t i = 0; i < num_pixels; ++i) rgba[size_t(i) * 4 + 3] = std::fabs(alpha[i] / film_pixels[size_t(i) * 4 + 3]);
}

// Opaque.resolveTonemap with the Linear tonemapper, buffer_opaque.zig:73-79, tonemapper.zig:36-39, aces.zig:19-27
void zo_resolve(const ZygpuView* view, const float* film_pixels, uint32_t num_pixels, float* rgba) {
    using namespace zo;
    for (uint32_t i = 0; i < num_pixels; ++i) {
        const float* p = film_pixels + size_t(i) * 4;
        Vec4f        color = Vec4f{{p[0], p[1], p[2], 0.f}} / splat(p[3]);
        for (int k = 0; k < 4; ++k) color[k] = std::fabs(color[k]);
        const Vec4f scaled = splat(view->exposure_factor) * color;
        const Vec4f srgb   = Vec4f{{1.70505155f, -0.13025714f, -0.02400328f, 0.f}} * splat(scaled[0]) +
                           Vec4f{{-0.62179068f, 1.14080289f, -0.12896877f, 0.f}} * splat(scaled[1]) +
                           Vec4f{{-0.08325840f, -0.01054853f, 1.15297171f, 0.f}} * splat(scaled[2]);
        rgba[size_t(i) * 4 + 0] = srgb[0];
        rgba[size_t(i) * 4 + 1] = srgb[1];
        rgba[size_t(i) * 4 + 2] = srgb[2];
        rgba[size_t(i) * 4 + 3] = 1.f;
    }
}

// Denoise.init + process + filter + estimateNoise, src/it/denoise.zig
void zo_denoise(const ZygpuView* view, const float* film, const float* normal_layer, const float* albedo_layer, float sigma, float* rgba) {
    using namespace zo;
    const int32_t w = view->resolution[0], h = view->resolution[1];

    const int32_t      radius = int32_t(std::ceil(3.f * sigma));  // :34-72
    std::vector<float> weights;
    float              wsum   = 0.f;
    const float        sigma2 = sigma * sigma;
    for (int32_t y = -radius; y <= radius; ++y) {
        for (int32_t x = -radius; x <= radius; ++x) {
            const float p = (float(x) * float(x) + float(y) * float(y)) / (2.f * sigma2);
            weights.push_back(std::exp(-p));
            wsum += weights.back();
        }
    }
    for (float& g : weights) g /= wsum;

    auto colorAt = [&](int32_t x, int32_t y) {  // Opaque.resolveTonemap (Linear) before the matrix to sRGB primaries
        const float* p = film + (size_t(y) * size_t(w) + size_t(x)) * 4;
        return splat(view->exposure_factor) * Vec4f{{std::fabs(p[0] / p[3]), std::fabs(p[1] / p[3]), std::fabs(p[2] / p[3]), 0.f}};
    };
    auto normalAt = [&](int32_t x, int32_t y) {
        const float* p = normal_layer + (size_t(y) * size_t(w) + size_t(x)) * 4;
        return Vec4f{{p[0] / p[3], p[1] / p[3], p[2] / p[3], 0.f}};
    };
    auto albedoAt = [&](int32_t x, int32_t y) {
        const float* p = albedo_layer + (size_t(y) * size_t(w) + size_t(x)) * 4;
        return Vec4f{{std::fabs(p[0]) / p[3], std::fabs(p[1]) / p[3], std::fabs(p[2]) / p[3], 0.f}};
    };
    auto luma = [](Vec4f c) { return std::pow(hmax3(c), 1.f / 2.2f); };

    for (int32_t py = 0; py < h; ++py) {
        for (int32_t px = 0; px < w; ++px) {
            const Vec4f ref_color = colorAt(px, py), ref_n = normalAt(px, py), ref_albedo = albedoAt(px, py);

            float sum = 0.f, l[9];  // estimateNoise, :375-451
            for (int32_t y = -1, k = 0; y <= 1; ++y) {
                for (int32_t x = -1; x <= 1; ++x, ++k) {
                    l[k] = luma(colorAt(std::clamp(px + x, 0, w - 1), std::clamp(py + y, 0, h - 1)));
                    sum += l[k];
                }
            }
            const float norm = 1.f / 9.f;
            const float mean = sum * norm;
            float       dif_sum = 0.f;
            for (int k = 0; k < 9; ++k) {
                const float dif = l[k] - mean;
                dif_sum += dif * dif;
            }
            const float std_dev        = std::sqrt(norm * dif_sum);
            const float coef           = mean > 0.f ? std_dev / mean : 0.f;
            const float noise_estimate = min(coef * 20.f * min(mean, 1.f), 1.f);

            Vec4f    result = splat(0.f);  // filter, :175-246 (dd is overwritten with 1: the depth input has no effect)
            uint32_t tap    = 0;
            for (int32_t y = -radius; y <= radius; ++y) {
                for (int32_t x = -radius; x <= radius; ++x) {
                    const int32_t sx = std::clamp(px + x, 0, w - 1), sy = std::clamp(py + y, 0, h - 1);
                    const float   weight      = weights[tap++];
                    const float   dot_n       = saturate(dot3(ref_n, normalAt(sx, sy)));
                    const float   dist_albedo = min(distance3(ref_albedo, albedoAt(sx, sy)), 1.f);
                    const float   strength    = 1.f * (dot_n * dot_n) * (1.f - dist_albedo) * noise_estimate;
                    const Vec4f   color       = lerp(ref_color, colorAt(sx, sy), splat(strength));
                    result                    = result + splat(weight) * color;
                }
            }
            const Vec4f srgb = Vec4f{{1.70505155f, -0.13025714f, -0.02400328f, 0.f}} * splat(result[0]) +
                               Vec4f{{-0.62179068f, 1.14080289f, -0.12896877f, 0.f}} * splat(result[1]) +
                               Vec4f{{-0.08325840f, -0.01054853f, 1.15297171f, 0.f}} * splat(result[2]);
            float* o = rgba + (size_t(py) * size_t(w) + size_t(px)) * 4;
            o[0] = srgb[0], o[1] = srgb[1], o[2] = srgb[2], o[3] = 1.f;
        }
    }
}

// aov.Buffer.resolve, aov_buffer.zig:51-82
void zo_resolve_aov(uint32_t aov_class, const float* layer, uint32_t num_pixels, float* rgba) {
    using namespace zo;
    for (uint32_t i = 0; i < num_pixels; ++i) {
        const float* p = layer + size_t(i) * 4;
        float*       o = rgba + size_t(i) * 4;
        if (ZYG_AOV_ALBEDO == aov_class || aov_class >= ZYG_AOV_EMISSION) {
            const Vec4f color = Vec4f{{std::fabs(p[0]), std::fabs(p[1]), std::fabs(p[2]), 0.f}} / splat(p[3]);
            const Vec4f srgb  = Vec4f{{1.70505155f, -0.13025714f, -0.02400328f, 0.f}} * splat(color[0]) +
                               Vec4f{{-0.62179068f, 1.14080289f, -0.12896877f, 0.f}} * splat(color[1]) +
                               Vec4f{{-0.08325840f, -0.01054853f, 1.15297171f, 0.f}} * splat(color[2]);
            o[0] = srgb[0], o[1] = srgb[1], o[2] = srgb[2], o[3] = 1.f;
        } else if (ZYG_AOV_GEOMETRIC_NORMAL == aov_class || ZYG_AOV_SHADING_NORMAL == aov_class) {
            o[0] = p[0] / p[3], o[1] = p[1] / p[3], o[2] = p[2] / p[3], o[3] = 1.f;
        } else if (ZYG_AOV_ROUGHNESS == aov_class) {
            o[0] = p[0] / p[3], o[1] = 0.f, o[2] = 0.f, o[3] = 1.f;
        } else {
            o[0] = p[0], o[1] = 0.f, o[2] = 0.f, o[3] = 1.f;
        }
    }
}

// integrate_micro_directional_albedo of the reference's LUT generator (ggx_integrate.zig:27-57) through this
// oracle's ggx::iso::reflect: lets tests pin the GGX restatement against the E_m table the reference ships.
float zo_ggx_micro_directional_albedo(float alpha, float n_dot_wo, uint32_t num_samples) {
    using namespace zo;
    if (0.f == alpha) return 1.f;
    const float            calpha = max(alpha, ggx::MinAlpha);
    const fresnel::Schlick schlick{splat(1.f)};
    const Frame            frame{{{1.f, 0.f, 0.f, 0.f}}, {{0.f, 1.f, 0.f, 0.f}}, {{0.f, 0.f, 1.f, 0.f}}};
    const Vec4f            wo = {{std::sqrt(1.f - n_dot_wo * n_dot_wo), 0.f, n_dot_wo, 0.f}};

    float accum = 0.f;
    for (uint32_t i = 0; i < num_samples; ++i) {
        // math.hammersley(i, num_samples, 0), sample_distribution.zig:3-18
        uint32_t bits = i;
        bits          = (bits << 16) | (bits >> 16);
        bits          = ((bits & 0x55555555u) << 1) | ((bits & 0xAAAAAAAAu) >> 1);
        bits          = ((bits & 0x33333333u) << 2) | ((bits & 0xCCCCCCCCu) >> 2);
        bits          = ((bits & 0x0F0F0F0Fu) << 4) | ((bits & 0xF0F0F0F0u) >> 4);
        bits          = ((bits & 0x00FF00FFu) << 8) | ((bits & 0xFF00FF00u) >> 8);
        const float xi[2] = {float(i) / float(num_samples), float(bits) * 2.3283064365386963e-10f};

        bxdf::Sample     result;
        const ggx::Micro micro = ggx::iso::reflect(wo, n_dot_wo, calpha, 0.f, xi, schlick, frame, result);
        accum += ((micro.n_dot_wi * result.reflection[0]) / result.pdf) / float(num_samples);
    }
    return accum;
}

// integrate_directional_albedo of the reference's LUT generator (ggx_integrate.zig:89-116): ggx.Iso.reflect with Schlick(f0)
// plus the multi-scatter term dspbrMicroEc over the E_m / E_m_avg tables. Recomputing the E table through it pins Schlick,
// dspbrMicroEc and the bilinear / linear table evaluation against numbers the reference holds.
float zo_ggx_directional_albedo(const float* luts_base, float alpha, float f0, float n_dot_wo, uint32_t num_samples) {
    using namespace zo;
    const GgxLuts          luts(luts_base);
    const float            calpha = max(alpha, ggx::MinAlpha);
    const fresnel::Schlick schlick{splat(f0)};
    const Frame            frame{{{1.f, 0.f, 0.f, 0.f}}, {{0.f, 1.f, 0.f, 0.f}}, {{0.f, 0.f, 1.f, 0.f}}};
    const Vec4f            wo = {{std::sqrt(1.f - n_dot_wo * n_dot_wo), 0.f, n_dot_wo, 0.f}};

    float accum = 0.f;
    for (uint32_t i = 0; i < num_samples; ++i) {
        uint32_t bits = i;
        bits          = (bits << 16) | (bits >> 16);
        bits          = ((bits & 0x55555555u) << 1) | ((bits & 0xAAAAAAAAu) >> 1);
        bits          = ((bits & 0x33333333u) << 2) | ((bits & 0xCCCCCCCCu) >> 2);
        bits          = ((bits & 0x0F0F0F0Fu) << 4) | ((bits & 0xF0F0F0F0u) >> 4);
        bits          = ((bits & 0x00FF00FFu) << 8) | ((bits & 0xFF00FF00u) >> 8);
        const float xi[2] = {float(i) / float(num_samples), float(bits) * 2.3283064365386963e-10f};

        bxdf::Sample     result;
        const ggx::Micro micro = ggx::iso::reflect(wo, n_dot_wo, calpha, 0.f, xi, schlick, frame, result);
        const float      mms   = ggx::dspbrMicroEc(luts, splat(f0), micro.n_dot_wi, n_dot_wo, calpha)[0];
        accum += ((micro.n_dot_wi * (result.reflection[0] + mms)) / result.pdf) / float(num_samples);
    }
    return min(accum, 1.f);
}

// integrate_average_albedo (ggx_integrate.zig:118-132): the cosine-weighted mean of the E table (trilinear evaluation).
float zo_ggx_average_albedo(const float* luts_base, float alpha, float f0, uint32_t num_samples) {
    using namespace zo;
    const GgxLuts luts(luts_base);
    float         accum = 0.f;
    for (uint32_t i = 0; i < num_samples; ++i) {
        uint32_t bits = i;
        bits          = (bits << 16) | (bits >> 16);
        bits          = ((bits & 0x55555555u) << 1) | ((bits & 0xAAAAAAAAu) >> 1);
        bits          = ((bits & 0x33333333u) << 2) | ((bits & 0xCCCCCCCCu) >> 2);
        bits          = ((bits & 0x0F0F0F0Fu) << 4) | ((bits & 0xF0F0F0F0u) >> 4);
        bits          = ((bits & 0x00FF00FFu) << 8) | ((bits & 0xFF00FF00u) >> 8);
        const float xi[2] = {float(i) / float(num_samples), float(bits) * 2.3283064365386963e-10f};
        // smpl.hemisphereCosine(xi)[2], sampling.zig
        float xy[2];
        diskConcentric(xi, xy);
        const float z = std::sqrt(max(0.f, 1.f - xy[0] * xy[0] - xy[1] * xy[1]));
        accum += luts.e(z, alpha, f0) / float(num_samples);
    }
    return accum;
}

// integrate_micro_average_albedo (ggx_integrate.zig:59-73): the cosine-weighted mean of the E_m table (bilinear evaluation).
float zo_ggx_micro_average_albedo(const float* luts_base, float alpha, uint32_t num_samples) {
    using namespace zo;
    const GgxLuts luts(luts_base);
    float         accum = 0.f;
    for (uint32_t i = 0; i < num_samples; ++i) {
        uint32_t bits = i;
        bits          = (bits << 16) | (bits >> 16);
        bits          = ((bits & 0x55555555u) << 1) | ((bits & 0xAAAAAAAAu) >> 1);
        bits          = ((bits & 0x33333333u) << 2) | ((bits & 0xCCCCCCCCu) >> 2);
        bits          = ((bits & 0x0F0F0F0Fu) << 4) | ((bits & 0xF0F0F0F0u) >> 4);
        bits          = ((bits & 0x00FF00FFu) << 8) | ((bits & 0xFF00FF00u) >> 8);
        const float xi[2] = {float(i) / float(num_samples), float(bits) * 2.3283064365386963e-10f};
        float       xy[2];
        diskConcentric(xi, xy);
        const float z = std::sqrt(max(0.f, 1.f - xy[0] * xy[0] - xy[1] * xy[1]));
        accum += luts.eM(z, alpha) / float(num_samples);
    }
    return accum;
}

// integrate_f_s_ss of the reference's LUT generator (ggx_integrate.zig:134-205) through this oracle's VNDF sampling,
// reflectNoFresnel / refractNoFresnel and schlick1: lets tests pin the rough-dielectric lobes of Glass against the E_s
// table the reference ships (ggx_integral.zig:1045-1046 ff.).
float zo_ggx_f_s_ss(float alpha, float f0, float ior_t, float n_dot_wo, uint32_t num_samples) {
    using namespace zo;
    if (alpha < ggx::MinAlpha || ior_t <= 1.f) return 1.f;
    const Frame frame{{{1.f, 0.f, 0.f, 0.f}}, {{0.f, 1.f, 0.f, 0.f}}, {{0.f, 0.f, 1.f, 0.f}}};
    const float cn_dot_wo = safe::clamp(n_dot_wo);
    const Vec4f wo        = {{std::sqrt(1.f - cn_dot_wo * cn_dot_wo), 0.f, cn_dot_wo, 0.f}};
    const IoR   ior{ior_t, 1.f};

    float accum = 0.f;
    for (uint32_t i = 0; i < num_samples; ++i) {
        uint32_t bits = i;  // math.hammersley(i, num_samples, 0), sample_distribution.zig:3-18
        bits          = (bits << 16) | (bits >> 16);
        bits          = ((bits & 0x55555555u) << 1) | ((bits & 0xAAAAAAAAu) >> 1);
        bits          = ((bits & 0x33333333u) << 2) | ((bits & 0xCCCCCCCCu) >> 2);
        bits          = ((bits & 0x0F0F0F0Fu) << 4) | ((bits & 0xF0F0F0F0u) >> 4);
        bits          = ((bits & 0x00FF00FFu) << 8) | ((bits & 0xFF00FF00u) >> 8);
        const float xi[2] = {float(i) / float(num_samples), float(bits) * 2.3283064365386963e-10f};

        float       n_dot_h;
        const float a2[2] = {alpha, alpha};
        const Vec4f h     = ggx::sampleVndf(wo, a2, xi, frame, n_dot_h);

        const float wo_dot_h = safe::clampDot(wo, h);
        const float eta      = ior.eta_i / ior.eta_t;
        const float sint2    = (eta * eta) * (1.f - wo_dot_h * wo_dot_h);
        const float wi_dot_h = std::sqrt(1.f - sint2);
        const float cos_x    = ior.eta_i > ior.eta_t ? wi_dot_h : wo_dot_h;
        const float f        = fresnel::schlick1(cos_x, f0);

        bxdf::Sample result;
        {
            const float n_dot_wi = ggx::iso::reflectNoFresnel(wo, h, cn_dot_wo, n_dot_h, wo_dot_h, alpha, 0.f, frame, result);
            accum += (min(n_dot_wi, n_dot_wo) * f * result.reflection[0]) / result.pdf;
        }
        {
            const float n_dot_wi = ggx::iso::refractNoFresnel(wo, h, cn_dot_wo, n_dot_h, -wi_dot_h, -wo_dot_h, alpha, 0.f, ior, frame, result);
            accum += (n_dot_wi * (1.f - f) * result.reflection[0]) / result.pdf;
        }
    }
    return accum / float(num_samples);
}

// Sobol.sample1D stream: startPixel(sample, seed) then n draws with incrementPadding every `pad_every` draws (0 = never).
void zo_sobol_stream(uint32_t sample, uint32_t seed, uint32_t n, uint32_t pad_every, float* out) {
    zo::Sobol s;
    s.startPixel(sample, seed);
    for (uint32_t i = 0; i < n; ++i) {
        out[i] = s.sample1D();
        if (pad_every && 0 == (i + 1) % pad_every) s.incrementPadding();
    }
}

void zo_sobol_directions(uint32_t* out) { std::memcpy(out, zo::sobolDirections().d, sizeof(uint32_t) * 160); }

}  // extern "C"
