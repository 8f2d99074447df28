#include "render_common.cuh"

namespace zygpu {

namespace {

// Shape.fragment, shape.zig:205-219
__device__ __forceinline__ void shapeFragment(const SceneDevice& sc, uint32_t prop, const RayT& ray, const HitD& isec, FragD& frag) {
    frag.prop      = prop;
    frag.primitive = isec.primitive;
    frag.trafo     = loadTrafo(sc.trafos, prop);
    switch (sc.props[prop].shape) {
        case ZYG_SHAPE_CUBE: cubeFragment(ray, isec, frag); break;
        case ZYG_SHAPE_RECTANGLE: rectangleFragment(ray, isec, frag); break;
        case ZYG_SHAPE_DISK: diskFragment(ray, isec, frag); break;
        case ZYG_SHAPE_SPHERE: sphereFragment(ray, isec, frag); break;
        // Distant and Canopy are infinite props: only met on escape, where shade_a calls their fragment functions itself
        case ZYG_SHAPE_TRIANGLE_MESH: meshFragment(sc.mesh_shading[sc.props[prop].mesh], isec, frag); break;
        default: break;
    }
}

// ---- lights ----------------------------------------------------------------------------------

struct VertexD {  // the parts of Vertex (vertex.zig:45-64) the light code reads
    RayT     ray;
    V3       origin, geo_n;
    float    bxdf_pdf, light_split_threshold;
    uint32_t state, probe_depth;
};

__device__ __forceinline__ uint32_t lightNumSamples(const ZygpuLight& l, float split_threshold) {  // shape_sampler.zig:35-41
    return split_threshold <= kLowThreshold ? 1u : l.num_samples;
}

// C. Ureña, M. Fajardo, A. King: An Area-Preserving Parametrization for Spherical Rectangles. rectangle.zig:199-303
struct SphQuadD {
    V3    o, x, y, z;
    float z0, x0, y0, x1, y1, b0, b1, k, S;

    __device__ void init(V3 scale, V3 origin) {
        const V3 s  = {-0.5f * scale.x, -0.5f * scale.y, 0.f};
        const V3 ex = {scale.x, 0.f, 0.f};
        const V3 ey = {0.f, scale.y, 0.f};

        o               = origin;
        const float exl = length3(ex);
        const float eyl = length3(ey);
        x               = divs3(ex, exl);
        y               = divs3(ey, eyl);
        z               = cross3(x, y);
        const V3 d      = sub3(s, o);
        z0              = dot3(d, z);
        if (z0 > 0.f) {
            z  = neg3(z);
            z0 = -z0;
        }
        x0 = dot3(d, x);
        y0 = dot3(d, y);
        x1 = x0 + exl;
        y1 = y0 + eyl;

        const V3 v00 = {x0, y0, z0}, v01 = {x0, y1, z0}, v10 = {x1, y0, z0}, v11 = {x1, y1, z0};

        const V3 n0 = normalize3(cross3(v00, v10));
        const V3 n1 = normalize3(cross3(v10, v11));
        const V3 n2 = normalize3(cross3(v11, v01));
        const V3 n3 = normalize3(cross3(v01, v00));

        const float g0 = acosf(-dot3(n0, n1));
        const float g1 = acosf(-dot3(n1, n2));
        const float g2 = acosf(-dot3(n2, n3));
        const float g3 = acosf(-dot3(n3, n0));

        b0 = n0.z;
        b1 = n2.z;
        k  = 2.f * kPi - g2 - g3;
        S  = g0 + g1 - k;
    }

    __device__ V3 sample(float u0, float u1) const {
        const float au = u0 * S + k;
        float       sa, ca;
        sincosf(au, &sa, &ca);
        const float fu = __fdiv_rn(ca * b0 - b1, sa);
        float       cu = __fdiv_rn(1.f, __fsqrt_rn(fu * fu + b0 * b0)) * (fu > 0.f ? 1.f : -1.f);
        cu             = cu < -1.f ? -1.f : (cu > 1.f ? 1.f : cu);

        float xu = __fdiv_rn(-(cu * z0), __fsqrt_rn(1.f - cu * cu));
        xu       = xu < x0 ? x0 : (xu > x1 ? x1 : xu);

        const float d   = __fsqrt_rn(xu * xu + z0 * z0);
        const float h0  = __fdiv_rn(y0, __fsqrt_rn(d * d + y0 * y0));
        const float h1  = __fdiv_rn(y1, __fsqrt_rn(d * d + y1 * y1));
        const float hv  = h0 + u1 * (h1 - h0);
        const float hv2 = hv * hv;
        const float eps = __uint_as_float(0x35800000u);
        const float yv  = hv2 < 1.f - eps ? __fdiv_rn(hv * d, __fsqrt_rn(1.f - hv2)) : y1;

        return add3(add3(add3(o, scale3(xu, x)), scale3(yv, y)), scale3(z0, z));
    }

    __device__ float pdf(V3 scale) const {
        const float sqr_dist = squaredLength3(o);
        const float area     = scale.x * scale.y;
        const float numer    = area * fabsf(o.z);
        const float denom    = sqr_dist * __fsqrt_rn(sqr_dist);
        return numer > denom * kDotMin ? __fdiv_rn(1.f, S) : __fdiv_rn(denom, numer);
    }
};

struct LightPropsD {  // light.zig:25-30, scene.zig:664-674
    V3    center;
    float radius;
    V3    cone_axis;
    float cone_cos;
    float power;
    bool  two_sided;
};

__device__ __forceinline__ LightPropsD lightProperties(const SceneDevice& sc, uint32_t light_id) {
    const float4 mi   = __ldg(sc.light_aabbs + 2 * size_t(light_id));
    const float4 ma   = __ldg(sc.light_aabbs + 2 * size_t(light_id) + 1);
    const float4 cone = __ldg(sc.light_cones + light_id);
    return {{0.5f * (mi.x + ma.x), 0.5f * (mi.y + ma.y), 0.5f * (mi.z + ma.z)}, ma.w, {cone.x, cone.y, cone.z}, cone.w, mi.w,
            0 != sc.lights[light_id].two_sided};
}

__device__ __forceinline__ float clampedCosSub(float cos_a, float cos_b, float sin_a, float sin_b) {  // light_tree.zig:217-220
    const float angle = __fmaf_rn(cos_a, cos_b, sin_a * sin_b);
    return cos_a > cos_b ? 1.f : angle;
}
__device__ __forceinline__ float clampedSinSub(float cos_a, float cos_b, float sin_a, float sin_b) {  // :222-225
    const float angle = __fmaf_rn(sin_a, cos_b, -sin_b * cos_a);
    return cos_a > cos_b ? 0.f : angle;
}

// light_tree.zig:173-215. Out of line on purpose: every argument is a value (nothing forces kernel parameters into local memory), and the
// tree code calls it from a dozen places - inlined, its IEEE divisions and square roots alone were a fifth of the shade kernels' code,
// which are bound by instruction fetch (ncu: `no_instruction` next to `long_scoreboard`).
#ifndef ZYGPU_IMPORTANCE_INLINE
#define ZYGPU_IMPORTANCE_INLINE __noinline__
#endif
__device__ ZYGPU_IMPORTANCE_INLINE float lightImportance(V3 p, V3 n, V3 center, V3 cone_axis, float cos_cone, float radius, float power, bool two_sided,
                                 bool total_sphere) {
    const V3    axis = sub3(p, center);
    const float l    = length3(axis);
    const V3    na   = divs3(axis, l);

    const float sin_cu = zmin(__fdiv_rn(radius, l), 1.f);
    const float dca    = dot3(cone_axis, na);
    const float cos_a  = two_sided ? fabsf(dca) : dca;
    const float cos_n  = zmax(-dot3(n, na), 0.f);

    const float cos_cu   = __fsqrt_rn(zmax(__fmaf_rn(sin_cu, -sin_cu, 1.f), 0.f));
    const float sin_cone = __fsqrt_rn(zmax(__fmaf_rn(cos_cone, -cos_cone, 1.f), 0.f));
    const float sin_a    = __fsqrt_rn(zmax(__fmaf_rn(cos_a, -cos_a, 1.f), 0.f));
    const float sin_n    = __fsqrt_rn(zmax(__fmaf_rn(cos_n, -cos_n, 1.f), 0.f));

    const float ta = clampedCosSub(cos_a, cos_cone, sin_a, sin_cone);
    const float tb = clampedSinSub(cos_a, cos_cone, sin_a, sin_cone);
    const float tc = clampedCosSub(ta, cos_cu, tb, sin_cu);
    const float tn = clampedCosSub(cos_n, cos_cu, sin_n, sin_cu);

    const float ra = total_sphere ? 1.f : tn;
    const float rb = zmax(tc, 0.f);

    const float clamped_dist = zmax(l, 0.5f * radius);
    const float rc           = __fdiv_rn(power, clamped_dist * clamped_dist);

    return zmax(ra * rb * rc, 0.f);
}

// The tree a traversal runs over: the scene's (lights = scene lights) or the PrimitiveTree of a mesh sampler (lights = the
// emitting triangles of the part, light_tree.zig:520-719).
struct TreeD {
    const ZygpuLightNode*    nodes;
    const uint32_t*          middles;
    const uint32_t*          orders;
    const uint32_t*          mapping;
    float4                   bounds_min, bounds_max;
    const MeshSamplerDevice* sampler;  // null for the scene tree
};

__device__ __forceinline__ TreeD sceneTree(const SceneDevice& sc) {
    return {sc.lt_nodes, sc.lt_middles, sc.lt_orders, sc.lt_mapping, sc.lt_bounds_min, sc.lt_bounds_max, nullptr};
}
__device__ __forceinline__ TreeD primitiveTree(const MeshSamplerDevice& m) {
    return {m.nodes, m.node_middles, m.light_orders, m.light_mapping, m.bounds_min, m.bounds_max, &m};
}

__device__ __forceinline__ V3 meshPosition(const MeshDevice& mesh, uint32_t index) {
    return {__ldg(mesh.positions + 3 * size_t(index)), __ldg(mesh.positions + 3 * size_t(index) + 1), __ldg(mesh.positions + 3 * size_t(index) + 2)};
}
__device__ __forceinline__ void meshTriangle(const MeshDevice& mesh, uint32_t t, V3& a, V3& b, V3& c) {
    a = meshPosition(mesh, __ldg(mesh.triangles + 3 * size_t(t)));
    b = meshPosition(mesh, __ldg(mesh.triangles + 3 * size_t(t) + 1));
    c = meshPosition(mesh, __ldg(mesh.triangles + 3 * size_t(t) + 2));
}

// MeshImpl.lightProperties, shape_sampler.zig:198-226. The reference derives centre, bounding radius and normal of the triangle every
// time a leaf of the primitive tree weighs it; here one kernel does that at upload with the same arithmetic and the walks read 32 bytes.
__global__ void meshLightPropsKernel(MeshDevice mesh, MeshSamplerDevice m, float4* __restrict__ props) {
    const uint32_t light = blockIdx.x * blockDim.x + threadIdx.x;
    if (light >= m.num_triangles) return;
    V3 a, b, c;
    meshTriangle(mesh, __ldg(m.triangle_mapping + light), a, b, c);
    const V3    center = divs3(add3(add3(a, b), c), 3.f);
    const float sra    = squaredLength3(sub3(a, center));
    const float srb    = squaredLength3(sub3(b, center));
    const float src    = squaredLength3(sub3(c, center));
    const float radius = __fsqrt_rn(zmax(sra, zmax(srb, src)));
    const V3    nn     = normalize3(cross3(sub3(b, a), sub3(c, a)));
    props[2 * size_t(light)]     = make_float4(center.x, center.y, center.z, radius);
    props[2 * size_t(light) + 1] = make_float4(nn.x, nn.y, nn.z, __ldg(m.triangle_pdfs + light));
}

__device__ __forceinline__ LightPropsD meshLightProperties(const SceneDevice&, const MeshSamplerDevice& m, uint32_t light) {
    const float4 a = __ldg(m.triangle_props + 2 * size_t(light));
    const float4 b = __ldg(m.triangle_props + 2 * size_t(light) + 1);
    return {{a.x, a.y, a.z}, a.w, {b.x, b.y, b.z}, 1.f, b.w, 0 != m.two_sided};
}

__device__ __forceinline__ float lightWeight(const SceneDevice& sc, const TreeD& tr, V3 p, V3 n, bool total_sphere, uint32_t light) {  // light_tree.zig:227-233
    const LightPropsD lp = tr.sampler ? meshLightProperties(sc, *tr.sampler, light) : lightProperties(sc, light);
    return lightImportance(p, n, lp.center, lp.cone_axis, lp.cone_cos, lp.radius, lp.power, lp.two_sided, total_sphere);
}

struct LightNodeD {
    V3       center;
    float    radius;
    V3       cone_axis;
    float    cone_cos;
    float    power, variance;
    uint32_t meta, num_lights;
};

__device__ __forceinline__ LightNodeD loadLightNode(const TreeD& sc, uint32_t i) {  // light_tree.zig:25-42
    const uint4* p  = reinterpret_cast<const uint4*>(sc.nodes + i);
    const uint4  a  = __ldg(p);
    const uint4  b  = __ldg(p + 1);
    const float  ku = 1.f / 65535.f;
    const float  tx = float(a.x & 0xffffu) * ku, ty = float(a.x >> 16) * ku, tz = float(a.y & 0xffffu) * ku, tw = float(a.y >> 16) * ku;
    LightNodeD   n;
    n.center    = {zlerp(sc.bounds_min.x, sc.bounds_max.x, tx), zlerp(sc.bounds_min.y, sc.bounds_max.y, ty),
                   zlerp(sc.bounds_min.z, sc.bounds_max.z, tz)};
    n.radius    = zlerp(sc.bounds_min.w, sc.bounds_max.w, tw);
    n.cone_axis = {__fmaf_rn(float(a.z & 0xffffu), 1.f / 32768.f, -1.f), __fmaf_rn(float(a.z >> 16), 1.f / 32768.f, -1.f),
                   __fmaf_rn(float(a.w & 0xffffu), 1.f / 32768.f, -1.f)};
    n.cone_cos  = __fmaf_rn(float(a.w >> 16), 1.f / 32768.f, -1.f);
    n.power     = __uint_as_float(b.x);
    n.variance  = __uint_as_float(b.y);
    n.meta      = b.z;
    n.num_lights = b.w;
    return n;
}

__device__ __forceinline__ bool lightNodeIsLeaf(const TreeD& tr, uint32_t i) {  // the `has_children` bit of Node.meta
    return 0 == (__ldg(reinterpret_cast<const uint32_t*>(tr.nodes + i) + 6) & 1u);
}

__device__ __forceinline__ float lightNodeWeight(const LightNodeD& node, V3 p, V3 n, bool total_sphere) {  // light_tree.zig:57-63
    return lightImportance(p, n, node.center, node.cone_axis, node.cone_cos, node.radius, node.power, 0 != (node.meta & 2u), total_sphere);
}

__device__ __forceinline__ bool lightNodeSplit(const LightNodeD& node, V3 p, float threshold) {  // light_tree.zig:65-89
    const float r = node.radius;
    const float d = zmin(length3(sub3(p, node.center)), 1.0e6f);
    const float a = zmax(d - r, 0.001f);
    const float b = d + r;

    const float eg  = __fdiv_rn(1.f, a * b);
    const float eg2 = eg * eg;
    const float a3  = a * a * a;
    const float b3  = b * b * b;
    const float e2g = __fdiv_rn(b3 - a3, 3.f * (b - a) * a3 * b3);
    const float vg  = e2g - eg2;

    const float ve = node.variance;
    const float ee = node.power;
    const float s2 = zmax(ve * vg + ve * eg2 + ee * ee * vg, 0.f);
    const float ns = __fdiv_rn(1.f, 1.f + __fsqrt_rn(s2));
    return ns < threshold;
}

struct LightPickD {
    uint32_t offset;
    float    pdf;
};

// Node.randomLight, light_tree.zig:91-145
__device__ __forceinline__ LightPickD lightNodeRandomLight(const SceneDevice& sc, const TreeD& tr, const LightNodeD& node, V3 p, V3 n, bool total_sphere,
                                           float random) {
    const uint32_t num_lights = node.num_lights;
    const uint32_t light      = node.meta >> 2;
    if (1 == num_lights) return {__ldg(tr.mapping + light), 1.f};

    uint32_t front = light;
    uint32_t back  = light + num_lights - 1;

    // The reference weighs the first and the last light, then the next one from whichever end the running sums point at. Here every
    // weight comes from one call site (stage 0 and 1 are the two initial ones): the lanes of a warp stay on one copy of the code
    // whichever end they advance, and the kernels hold one copy of Light properties + importance instead of four.
    float w_front = 0.f, w_back = 0.f;
    float w_sum_front = 0.f, w_sum_back = 0.f;
    float w_sum = 0.f;
    for (uint32_t stage = 0;; stage = stage < 2 ? stage + 1 : 2) {
        bool     advance_front;
        uint32_t index;
        if (0 == stage) {
            advance_front = true;
            index         = front;
        } else if (1 == stage) {
            advance_front = false;
            index         = back;
        } else {
            if (front == back) break;
            w_sum         = w_sum_front + w_sum_back;
            advance_front = w_sum_front <= random * w_sum;
            if (advance_front) {
                front += 1;
            } else {
                back -= 1;
            }
            if (front == back) {
                if (advance_front) w_front = w_back;
                break;
            }
            index = advance_front ? front : back;
        }
        const float w = lightWeight(sc, tr, p, n, total_sphere, __ldg(tr.mapping + index));
        if (advance_front) {
            w_front = w;
            w_sum_front += w;  // the first one: 0 + w
        } else {
            w_back = w;
            w_sum_back += w;
        }
    }
    if (0.f == w_sum) return {0, 0.f};
    return {__ldg(tr.mapping + front), __fdiv_rn(w_front, w_sum)};
}

// Node.pdf, light_tree.zig:147-170
__device__ __forceinline__ float lightNodePdf(const SceneDevice& sc, const TreeD& tr, const LightNodeD& node, V3 p, V3 n, bool total_sphere, uint32_t id) {
    const uint32_t num_lights = node.num_lights;
    if (1 == num_lights) return 1.f;
    const uint32_t light = node.meta >> 2;
    const uint32_t end   = light + num_lights;
    float          w_id = 0.f, sum = 0.f;
    for (uint32_t i = light; i < end; ++i) {
        const float lw = lightWeight(sc, tr, p, n, total_sphere, __ldg(tr.mapping + i));
        sum += lw;
        if (id == i) w_id = lw;
    }
    if (0.f == sum) return 0.f;
    return __fdiv_rn(w_id, sum);
}

constexpr uint32_t kMaxLightPicks = 64;  // Tree.MaxLights

// Tree.randomLight, light_tree.zig:346-447. `emit` is called for every pick in the reference's order.
template <typename Emit>
__device__ __forceinline__ void lightTreeRandomLight(const SceneDevice& sc, V3 p, V3 n, bool total_sphere, float random, float split_threshold, Emit&& emit) {
    float       ip    = 0.f;
    const bool  split = split_threshold > 0.f;
    const TreeD tr    = sceneTree(sc);

    if (split && sc.lt_num_infinite < kMaxLightPicks - 1) {
        for (uint32_t i = 0; i < sc.lt_num_infinite; ++i) emit(LightPickD{__ldg(sc.lt_mapping + i), 1.f});
    } else {
        ip = sc.lt_infinite_weight;
        if (random < sc.lt_infinite_guard) {
            // infinite_light_distribution.sampleDiscrete(random), distribution_1d.zig:56-60
            const uint32_t l = dist1dSample(sc.lt_infinite_cdf, sc.lt_num_infinite + 1, random);
            emit(LightPickD{__ldg(sc.lt_mapping + l), (__ldg(sc.lt_infinite_cdf + l + 1) - __ldg(sc.lt_infinite_cdf + l)) * ip});
            return;
        }
    }
    if (0 == sc.lt_num_nodes) return;

    const float    pd              = 1.f - ip;
    const uint32_t max_split_depth = sc.lt_max_split_depth;

    struct Value {
        float    pdf, random;
        uint32_t node, depth;
    };
    Value    stack[12];
    uint32_t end = 0;

    Value t{pd, __fdiv_rn(random - ip, pd), 0, split ? 0 : max_split_depth};
    stack[end++] = t;

    while (end > 0) {
        const LightNodeD node = loadLightNode(tr, t.node);
        if (0 != (node.meta & 1u)) {
            const bool     do_split = t.depth < max_split_depth && lightNodeSplit(node, p, split_threshold);
            const uint32_t c0       = node.meta >> 2;
            const uint32_t c1       = c0 + 1;
            if (do_split) {
                t.depth += 1;
                t.node       = c0;
                stack[end++] = {t.pdf, t.random, c1, t.depth};
            } else {
                t.depth = max_split_depth;

                float p0 = lightNodeWeight(loadLightNode(tr, c0), p, n, total_sphere);
                float p1 = lightNodeWeight(loadLightNode(tr, c1), p, n, total_sphere);

                const float pt = p0 + p1;
                if (0.f == pt) {
                    t = stack[--end];
                    continue;
                }
                p0 = __fdiv_rn(p0, pt);
                p1 = __fdiv_rn(p1, pt);
                if (t.random < p0) {
                    t.node = c0;
                    t.pdf *= p0;
                    t.random = __fdiv_rn(t.random, p0);
                } else {
                    t.node = c1;
                    t.pdf *= p1;
                    t.random = zmin(__fdiv_rn(t.random - p0, p1), 1.f);
                }
            }
        } else {
            const LightPickD pick = lightNodeRandomLight(sc, tr, node, p, n, total_sphere, t.random);
            if (pick.pdf > 0.f) emit(LightPickD{pick.offset, pick.pdf * t.pdf});
            t = stack[--end];
        }
    }
}

// PrimitiveTree.randomLight, light_tree.zig:577-650, as a walk that can be left and resumed between two iterations of the reference's
// loop: the light kernel gives every lane one iteration per step of its own loop, so a lane whose pick splits into 64 triangles does not
// hold 31 others that found one. `emit` receives (part triangle, pdf) in the reference's order.
struct PrimitiveWalkD {
    static constexpr uint32_t kMaxSplitDepth = 6;
    struct Value {
        float    pdf, random;
        uint32_t node, depth;
    };
    Value    stack[kMaxSplitDepth + 1];
    Value    t;
    uint32_t end;

    __device__ __forceinline__ void start(float random, float split_threshold) {
        t        = {1.f, random, 0, split_threshold > 0.f ? 0 : kMaxSplitDepth};
        stack[0] = t;
        end      = 1;
    }

    // one iteration; false once the stack is empty
    template <typename Emit>
    __device__ __forceinline__ bool step(const SceneDevice& sc, const TreeD& tr, V3 p, V3 n, bool total_sphere, float split_threshold, Emit&& emit) {
        const LightNodeD node = loadLightNode(tr, t.node);
        if (0 != (node.meta & 1u)) {
            const bool     do_split = t.depth < kMaxSplitDepth && lightNodeSplit(node, p, split_threshold);
            const uint32_t c0       = node.meta >> 2;
            const uint32_t c1       = c0 + 1;
            if (do_split) {
                t.depth += 1;
                t.node       = c0;
                stack[end++] = {t.pdf, t.random, c1, t.depth};
            } else {
                t.depth = kMaxSplitDepth;

                float p0 = lightNodeWeight(loadLightNode(tr, c0), p, n, total_sphere);
                float p1 = lightNodeWeight(loadLightNode(tr, c1), p, n, total_sphere);

                const float pt = p0 + p1;
                if (0.f == pt) {
                    t = stack[--end];
                    return end > 0;
                }
                p0 = __fdiv_rn(p0, pt);
                p1 = __fdiv_rn(p1, pt);
                if (t.random < p0) {
                    t.node = c0;
                    t.pdf *= p0;
                    t.random = __fdiv_rn(t.random, p0);
                } else {
                    t.node = c1;
                    t.pdf *= p1;
                    t.random = zmin(__fdiv_rn(t.random - p0, p1), 1.f);
                }
            }
        } else {
            const LightPickD pick = lightNodeRandomLight(sc, tr, node, p, n, total_sphere, t.random);
            if (pick.pdf > 0.f) emit(LightPickD{pick.offset, pick.pdf * t.pdf});
            t = stack[--end];
        }
        return end > 0;
    }
};

template <typename Emit>
__device__ __forceinline__ void primitiveTreeRandomLight(const SceneDevice& sc, const MeshSamplerDevice& m, V3 p, V3 n, bool total_sphere, float random,
                                         float split_threshold, Emit&& emit) {
    const TreeD    tr = primitiveTree(m);
    PrimitiveWalkD walk;
    walk.start(random, split_threshold);
    while (walk.step(sc, tr, p, n, total_sphere, split_threshold, emit)) {
    }
}

// PrimitiveTree.pdf, light_tree.zig:652-719
__device__ __forceinline__ float primitiveTreePdf(const SceneDevice& sc, const MeshSamplerDevice& m, V3 p, V3 n, bool total_sphere, float split_threshold,
                                  uint32_t id) {
    constexpr uint32_t kMaxSplitDepth = 6;
    const TreeD        tr             = primitiveTree(m);
    const uint32_t     lo             = __ldg(tr.orders + id);
    const bool         split          = split_threshold > 0.f;

    float    pd    = 1.f;
    uint32_t nid   = 0;
    uint32_t depth = split ? 0 : kMaxSplitDepth;
    for (;;) {
        const LightNodeD node = loadLightNode(tr, nid);
        if (0 != (node.meta & 1u)) {
            const bool     do_split = depth < kMaxSplitDepth && lightNodeSplit(node, p, split_threshold);
            const uint32_t c0       = node.meta >> 2;
            const uint32_t c1       = c0 + 1;
            const uint32_t middle   = __ldg(tr.middles + nid);
            if (do_split) {
                depth += 1;
                nid = lo < middle ? c0 : c1;
            } else {
                depth          = kMaxSplitDepth;
                const float p0 = lightNodeWeight(loadLightNode(tr, c0), p, n, total_sphere);
                const float p1 = lightNodeWeight(loadLightNode(tr, c1), p, n, total_sphere);
                const float pt = p0 + p1;
                if (0.f == pt) return 0.f;
                if (lo < middle) {
                    nid = c0;
                    pd *= __fdiv_rn(p0, pt);
                } else {
                    nid = c1;
                    pd *= __fdiv_rn(p1, pt);
                }
            }
        } else {
            return pd * lightNodePdf(sc, tr, node, p, n, total_sphere, lo);
        }
    }
}

// Tree.pdf, light_tree.zig:449-517
__device__ __forceinline__ float lightTreePdf(const SceneDevice& sc, V3 p, V3 n, bool total_sphere, float split_threshold, uint32_t id) {
    const TreeD    tr             = sceneTree(sc);
    const uint32_t lo             = __ldg(sc.lt_orders + id);
    const bool     split          = split_threshold > 0.f;
    const bool     split_infinite = split && sc.lt_num_infinite < kMaxLightPicks - 1;

    if (lo < sc.lt_infinite_end) {  // infinite_weight * infinite_light_distribution.pdfI(lo)
        return split_infinite ? 1.f : sc.lt_infinite_weight * (__ldg(sc.lt_infinite_cdf + lo + 1) - __ldg(sc.lt_infinite_cdf + lo));
    }
    if (0 == sc.lt_num_nodes) return 0.f;

    const float    ip              = split_infinite ? 0.f : sc.lt_infinite_weight;
    const uint32_t max_split_depth = sc.lt_max_split_depth;

    float    pd    = 1.f - ip;
    uint32_t nid   = 0;
    uint32_t depth = split ? 0 : max_split_depth;
    for (;;) {
        const LightNodeD node = loadLightNode(tr, nid);
        if (0 != (node.meta & 1u)) {
            const bool     do_split = depth < max_split_depth && lightNodeSplit(node, p, split_threshold);
            const uint32_t c0       = node.meta >> 2;
            const uint32_t c1       = c0 + 1;
            const uint32_t middle   = __ldg(sc.lt_middles + nid);
            if (do_split) {
                depth += 1;
                nid = lo < middle ? c0 : c1;
            } else {
                depth          = max_split_depth;
                const float p0 = lightNodeWeight(loadLightNode(tr, c0), p, n, total_sphere);
                const float p1 = lightNodeWeight(loadLightNode(tr, c1), p, n, total_sphere);
                const float pt = p0 + p1;
                if (0.f == pt) return 0.f;
                if (lo < middle) {
                    nid = c0;
                    pd *= __fdiv_rn(p0, pt);
                } else {
                    nid = c1;
                    pd *= __fdiv_rn(p1, pt);
                }
            }
        } else {
            return pd * lightNodePdf(sc, tr, node, p, n, total_sphere, lo);
        }
    }
}

// Mesh.pdf, triangle_mesh.zig:662-703. Inlined: only the MeshLights instances of the kernels contain it, and an out-of-line copy
// taking the scene by reference makes every thread copy the kernel's parameter structs to local memory.
__device__ __forceinline__ float meshLightPdf(const SceneDevice& sc, const MeshSamplerDevice& m, const VertexD& vertex, const FragD& frag) {
    const float n_dot_dir = fabsf(dot3(frag.geo_n, vertex.ray.d));

    const V3 op = frag.trafo.worldToObjectPoint(vertex.origin);
    const V3 on = frag.trafo.worldToObjectNormal(vertex.geo_n);

    const uint32_t pm      = __ldg(m.primitive_mapping + frag.primitive);
    const float    tri_pdf = primitiveTreePdf(sc, m, op, on, 0 != (vertex.state & kTranslucent), vertex.light_split_threshold, pm);

    V3 a, b, c;
    meshTriangle(sc.meshes[m.mesh], frag.primitive, a, b, c);
    const V3    ca       = mul3(mul3(frag.trafo.scale, frag.trafo.scale), cross3(sub3(b, a), sub3(c, a)));
    const float tri_area = 0.5f * length3(ca);
    const V3    center   = divs3(add3(add3(a, b), c), 3.f);

    if (__fdiv_rn(tri_area, length3(sub3(center, op))) > kAreaDistanceRatio) return tri_pdf * pdfSpherical(op, a, b, c);
    const float sl = squaredLength3(sub3(vertex.origin, frag.p));
    return __fdiv_rn(tri_pdf * sl, n_dot_dir * tri_area);
}

// One triangle of Mesh.sampleTo, triangle_mesh.zig:492-608: `sp` = (triangle of the part, its pdf) as PrimitiveTree.randomLight emits it
__device__ __forceinline__ uint32_t meshLightTriangleSample(const SceneDevice& sc, const PathState& st, uint32_t slot, const ZygpuLight& light,
                                                         LightPickD pick, const TrafoD& trafo, const FragD& frag, V3 n, V3 op, V3 on,
                                                         bool translucent, LightPickD sp, SamplerD& sampler, uint32_t num_records) {
    const MeshSamplerDevice& m = sc.mesh_samplers[light.sampler];
    const V3                 p = frag.p;
    const V3     scale_squared = mul3(trafo.scale, trafo.scale);
    V3 a, b, c;
    meshTriangle(sc.meshes[m.mesh], __ldg(m.triangle_mapping + sp.offset), a, b, c);

    const V3    ca  = mul3(scale_squared, cross3(sub3(b, a), sub3(c, a)));
    const float lca = length3(ca);
    V3          wn  = trafo.objectToWorldNormal(divs3(ca, lca));

    const float tri_area = 0.5f * lca;
    const V3    center   = divs3(add3(add3(a, b), c), 3.f);

    float u0, u1;
    sampler.sample2D(u0, u1);

    V3    dir, v;
    float sample_pdf, n_dot_dir;
    if (__fdiv_rn(tri_area, length3(sub3(center, op))) > kAreaDistanceRatio) {
        V3    sdir;
        float bu, bv, spdf;
        if (!sampleSpherical(op, a, b, c, u0, u1, sdir, bu, bv, spdf)) return num_records;
        if (dot3(sdir, on) <= 0.f && !translucent) return num_records;
        dir        = trafo.objectToWorldNormal(sdir);
        v          = trafo.objectToWorldPoint(interpolate3(a, b, c, bu, bv));
        sample_pdf = sp.pdf * spdf;
        if (0 != light.two_sided && dot3(wn, dir) > 0.f) wn = neg3(wn);
        n_dot_dir = -dot3(wn, dir);
    } else {
        float bu, bv;
        triangleUniform(u0, u1, bu, bv);
        v = trafo.objectToWorldPoint(interpolate3(a, b, c, bu, bv));

        const V3    axis = sub3(v, p);
        const float sl   = squaredLength3(axis);
        const float d    = __fsqrt_rn(sl);
        dir              = divs3(axis, d);
        if (dot3(dir, n) <= 0.f && !translucent) return num_records;
        if (0 != light.two_sided && dot3(wn, dir) > 0.f) wn = neg3(wn);
        n_dot_dir  = -dot3(wn, dir);
        sample_pdf = __fdiv_rn(sp.pdf * sl, n_dot_dir * tri_area);
    }
    if (n_dot_dir < kDotMin) return num_records;

    if (num_records < st.shadow_stride) {
        const size_t rec       = size_t(slot) * st.shadow_stride + num_records;
        const V3     origin    = frag.offsetP(dir);
        const V3     light_pos = offsetRay(v, wn);
        st.sh_o[rec]  = make_float4(origin.x, origin.y, origin.z, sample_pdf * pick.pdf);
        st.sh_p[rec]  = make_float4(light_pos.x, light_pos.y, light_pos.z, __uint_as_float(pick.offset));
        st.sh_wi[rec] = make_float4(dir.x, dir.y, dir.z, 0.f);
        num_records += 1;
    } else {
        st.counters[3] = 1;
    }
    return num_records;
}

// Mesh.sampleTo, triangle_mesh.zig:492-608: appends the shadow records of one picked mesh light, returns the new record count.
// Inlined for the same reason.
__device__ __forceinline__ uint32_t meshLightSampleTo(const SceneDevice& sc, const PathState& st, uint32_t slot, const ZygpuLight& light,
                                                   LightPickD pick, const TrafoD& trafo, const FragD& frag, V3 n, bool translucent,
                                                   float split_threshold, SamplerD& sampler, uint32_t num_records) {
    const MeshSamplerDevice& m  = sc.mesh_samplers[light.sampler];
    const V3                 op = trafo.worldToObjectPoint(frag.p);
    const V3                 on = trafo.worldToObjectNormal(n);
    const float              r1 = sampler.sample1D();
    primitiveTreeRandomLight(sc, m, op, on, translucent, r1, split_threshold, [&](LightPickD sp) {
        num_records = meshLightTriangleSample(sc, st, slot, light, pick, trafo, frag, n, op, on, translucent, sp, sampler, num_records);
    });
    return num_records;
}

// Scene.lightPdf, scene.zig:624-634 (+ Light.pdf -> Shape.pdf, light.zig:149-157, shape.zig:469-492, rectangle.zig:554-575)
// `MeshLights`: the scene has triangle-mesh lights (their sampling code is compiled out otherwise)
template <bool MeshLights>
__device__ __forceinline__ float sceneLightPdf(const SceneDevice& sc, const VertexD& vertex, const FragD& frag) {
    const uint32_t light_id = __ldg(sc.light_ids + sc.props[frag.prop].parts_start + frag.part);
    if (0 != (vertex.state & kSingular) || ZYGPU_NULL == light_id) return 1.f;

    const float select_pdf =
        lightTreePdf(sc, vertex.origin, vertex.geo_n, 0 != (vertex.state & kTranslucent), vertex.light_split_threshold, light_id);

    const ZygpuLight l          = sc.lights[light_id];
    float            sample_pdf = 0.f;
    if (ZYG_SHAPE_RECTANGLE == sc.props[l.prop].shape && ZYG_LIGHT_PROP_IMAGE == l.light_class) {  // Rectangle.materialPdf
        const float c            = fabsf(dot3(frag.trafo.r2, vertex.ray.d));
        const float area         = frag.trafo.scale.x * frag.trafo.scale.y;
        const float sl           = squaredLength3(sub3(vertex.origin, frag.p));
        const float material_pdf = imagePdf(sc.image_samplers[l.sampler], frag.u, frag.v) * float(lightNumSamples(l, vertex.light_split_threshold));
        sample_pdf               = __fdiv_rn(material_pdf * sl, c * area);
    } else if (ZYG_SHAPE_RECTANGLE == sc.props[l.prop].shape) {
        const float nsf = float(lightNumSamples(l, vertex.light_split_threshold));
        SphQuadD    squad;
        squad.init(frag.trafo.scale, frag.trafo.worldToFramePoint(vertex.origin));
        sample_pdf = nsf * squad.pdf(frag.trafo.scale);
    } else if (ZYG_SHAPE_DISTANT == sc.props[l.prop].shape) {  // Distant.pdf, distant.zig:139-141
        sample_pdf = __fdiv_rn(1.f, distantSolidAngle(frag.trafo.scale.x));
    } else if (ZYG_SHAPE_SPHERE == sc.props[l.prop].shape) {  // Sphere.pdf, sphere.zig:472-487
        sample_pdf = float(lightNumSamples(l, vertex.light_split_threshold)) * sphereLightPdf(frag.trafo, vertex.origin);
    } else if (ZYG_SHAPE_DISK == sc.props[l.prop].shape) {  // Disk.pdf, disk.zig:492-533
        sample_pdf = diskLightPdfLocal(frag.trafo.worldToFramePoint(vertex.origin), frag.trafo.worldToFramePoint(frag.p), 0.5f * frag.trafo.scale.x,
                                       fabsf(dot3(frag.trafo.r2, vertex.ray.d)), squaredLength3(sub3(vertex.origin, frag.p)),
                                       float(lightNumSamples(l, vertex.light_split_threshold)));
    } else if (ZYG_SHAPE_CANOPY == sc.props[l.prop].shape) {  // Light.propMaterialPdf -> Shape.materialPdf, shape.zig:519
        if (ZYG_LIGHT_PROP_IMAGE == l.light_class) sample_pdf = __fdiv_rn(imagePdf(sc.image_samplers[l.sampler], frag.u, frag.v), 2.f * kPi);
    } else if (MeshLights && ZYG_SHAPE_TRIANGLE_MESH == sc.props[l.prop].shape && ZYGPU_NULL != l.sampler) {
        sample_pdf = meshLightPdf(sc, sc.mesh_samplers[l.sampler], vertex, frag);
    }
    return powerHeuristic(vertex.bxdf_pdf, sample_pdf * select_pdf);
}

// Vertex.evaluateRadiance, vertex.zig:183-212
template <bool MeshLights>
__device__ __forceinline__ V3 evaluateRadiance(const SceneDevice& sc, const VertexD& vertex, const FragD& frag, SamplerD& sampler) {
    const V3            wo = neg3(vertex.ray.d);
    const ZygpuMaterial m  = sc.materials[__ldg(sc.material_ids + sc.props[frag.prop].parts_start + frag.part)];
    if (0 == (m.flags & ZYG_MATERIAL_EMISSIVE) || (0 == (m.flags & ZYG_MATERIAL_TWO_SIDED) && !frag.sameHemisphere(wo))) {
        return splat3(0.f);
    }
    const float stochastic_r = sampler.sample1D();  // rs.stochastic_r

    const bool  in_camera = 0 == vertex.probe_depth;
    const float area      = 0.f != m.emission_normalize ? shapeArea(sc.props[frag.prop].shape, frag.trafo.scale) : 1.f;
    const V3    energy    = ZYGPU_NULL != m.emission_map
                                ? emittanceRadianceMapped(m, wo, frag.trafo, area, in_camera,
                                                          imageTexel(sc.image_samplers[m.emission_map], frag.u, frag.v, stochastic_r))
                                : emittanceRadiance(m, wo, frag.trafo, area, in_camera);
    const float weight    = sceneLightPdf<MeshLights>(sc, vertex, frag);
    return scale3(weight, energy);
}

// Mesh.emission -> Tree.emission, triangle_mesh.zig:379-388, triangle_tree.zig:405-477: every triangle of an un-occluding mesh
// emitter the segment crosses, visited in the reference's order (binary tree, near child first) because each hit draws from
// the sampler.
template <bool MeshLights>
__device__ __forceinline__ V3 meshEmission(const SceneDevice& sc, uint32_t entity, const ZygpuProp& prop, const VertexD& vertex, SamplerD& sampler) {
    FragD frag;
    frag.prop  = entity;
    frag.trafo = loadTrafo(sc.trafos, entity);
    frag.t = frag.b = frag.n = splat3(0.f);
    frag.u = frag.v = 0.f;

    const MeshDevice&  mesh  = sc.meshes[prop.mesh];
    const MeshShading& shade = sc.mesh_shading[prop.mesh];
    const RayT         ray   = worldToObjectRay(frag.trafo, vertex.ray);

    uint32_t stack[64];
    uint32_t end = 0;
    uint32_t n   = 0;
    V3       energy = splat3(0.f);

    while (kEnd != n) {
        const float4 nmin = __ldg(mesh.binary_nodes + 2 * size_t(n));
        const float4 nmax = __ldg(mesh.binary_nodes + 2 * size_t(n) + 1);
        const uint32_t num = __float_as_uint(nmax.w);
        if (0 != num) {
            const uint32_t start = __float_as_uint(nmin.w);
            for (uint32_t i = start; i < start + num; ++i) {
                V3 a, b, c;
                meshTriangle(mesh, i, a, b, c);
                float ht, hu, hv;
                if (!intersectTriangle(ray, a, sub3(b, a), sub3(c, a), ht, hu, hv)) continue;
                frag.primitive = i;
                frag.part      = __ldg(shade.parts + i);
                frag.p         = frag.trafo.objectToWorldPoint(interpolate3(a, b, c, hu, hv));
                frag.geo_n     = frag.trafo.objectToWorldNormal(normalize3(cross3(sub3(b, a), sub3(c, a))));
                energy         = add3(energy, evaluateRadiance<MeshLights>(sc, vertex, frag, sampler));
            }
            n = 0 == end ? kEnd : stack[--end];
            continue;
        }
        uint32_t a = __float_as_uint(nmin.w);
        uint32_t b = a + 1;
        float dista = intersectNode(__ldg(mesh.binary_nodes + 2 * size_t(a)), __ldg(mesh.binary_nodes + 2 * size_t(a) + 1), ray);
        float distb = intersectNode(__ldg(mesh.binary_nodes + 2 * size_t(b)), __ldg(mesh.binary_nodes + 2 * size_t(b) + 1), ray);
        if (dista > distb) {
            const uint32_t tn = a;
            a                 = b;
            b                 = tn;
            const float td    = dista;
            dista             = distb;
            distb             = td;
        }
        if (FLT_MAX == dista) {
            n = 0 == end ? kEnd : stack[--end];
        } else {
            n = a;
            if (FLT_MAX != distb && end < 64) stack[end++] = b;
        }
    }
    return energy;
}

// Prop.emission + Shape.emission, prop.zig:239-264, shape.zig:283-299, rectangle.zig:188-196
template <bool MeshLights>
__device__ __forceinline__ V3 propEmission(const SceneDevice& sc, uint32_t entity, const VertexD& vertex, SamplerD& sampler) {
    const ZygpuProp prop = sc.props[entity];
    if (!propVisible(prop.flags, vertex.probe_depth)) return splat3(0.f);
    if (!aabbIntersect(sc.aabbs, entity, vertex.ray)) return splat3(0.f);
    if (MeshLights && ZYG_SHAPE_TRIANGLE_MESH == prop.shape) return meshEmission<MeshLights>(sc, entity, prop, vertex, sampler);
    if (ZYG_SHAPE_RECTANGLE != prop.shape && ZYG_SHAPE_SPHERE != prop.shape && ZYG_SHAPE_DISK != prop.shape) return splat3(0.f);

    FragD frag;
    frag.prop  = entity;
    frag.trafo = loadTrafo(sc.trafos, entity);
    HitD isec;
    if (ZYG_SHAPE_SPHERE == prop.shape) {  // Sphere.emission, sphere.zig:271-279
        if (!sphereIntersect(vertex.ray, frag.trafo, isec)) return splat3(0.f);
        sphereFragment(vertex.ray, isec, frag);
    } else if (ZYG_SHAPE_DISK == prop.shape) {  // Disk.emission, disk.zig:171-179
        if (!diskIntersect(vertex.ray, frag.trafo, isec)) return splat3(0.f);
        diskFragment(vertex.ray, isec, frag);
    } else {
        if (!rectangleIntersect(vertex.ray, frag.trafo, isec)) return splat3(0.f);
        rectangleFragment(vertex.ray, isec, frag);
    }
    return evaluateRadiance<MeshLights>(sc, vertex, frag, sampler);
}

// Context.emission -> PropBvh.emission, prop_tree.zig:302-356: every un-occluding emitter crossed before the hit
template <bool MeshLights>
__device__ __forceinline__ V3 unoccludingEmission(const SceneDevice& sc, const VertexD& vertex, SamplerD& sampler) {
    uint32_t stack[kPropStack];
    uint32_t end = 0;
    uint32_t n   = 0 == sc.num_unocc_nodes ? kEnd : 0;

    V3 energy = splat3(0.f);

    while (kEnd != n) {
        const float4 nmin = __ldg(sc.unocc_nodes + 2 * size_t(n));
        const float4 nmax = __ldg(sc.unocc_nodes + 2 * size_t(n) + 1);

        const uint32_t num = __float_as_uint(nmax.w);
        if (0 != num) {
            const uint32_t start = __float_as_uint(nmin.w);
            for (uint32_t i = start; i < start + num; ++i) {
                energy = add3(energy, propEmission<MeshLights>(sc, __ldg(sc.unocc_indices + i), vertex, sampler));
            }
            n = 0 == end ? kEnd : stack[--end];
            continue;
        }

        uint32_t a = __float_as_uint(nmin.w);
        uint32_t b = a + 1;

        float dista = intersectNode(__ldg(sc.unocc_nodes + 2 * size_t(a)), __ldg(sc.unocc_nodes + 2 * size_t(a) + 1), vertex.ray);
        float distb = intersectNode(__ldg(sc.unocc_nodes + 2 * size_t(b)), __ldg(sc.unocc_nodes + 2 * size_t(b) + 1), vertex.ray);
        if (dista > distb) {
            const uint32_t tn = a;
            a                 = b;
            b                 = tn;
            const float td    = dista;
            dista             = distb;
            distb             = td;
        }
        if (FLT_MAX == dista) {
            n = 0 == end ? kEnd : stack[--end];
        } else {
            n = a;
            if (FLT_MAX != distb) stack[end++] = b;
        }
    }
    return energy;
}

// ---- accumulation (IValue.add, helper.zig:11-19) ---------------------------------------------

__device__ __forceinline__ void ivalueAdd(const PathState& st, uint32_t slot, V3 value, uint32_t depth, uint32_t direct_cutoff,
                                          bool is_emission, bool singular) {
    float4* target = is_emission ? st.acc_e : ((singular || depth < direct_cutoff) ? st.acc_d : st.acc_i);
    float4  a      = target[slot];
    a.x += value.x;
    a.y += value.y;
    a.z += value.z;
    target[slot] = a;
}

// Pool.consume's bookkeeping for a vertex that ended without a successor (vertex.zig:243-268, no shadow catchers): a path that only went
// straight on / through interfaces adds what it lost on the way, any other path its whole weight
__device__ __forceinline__ void alphaAdd(const PathState& st, uint32_t slot, bool transparent, V3 throughput, float split_weight) {
    const float avg = __fdiv_rn((throughput.x + throughput.y) + throughput.z, 3.f);
    st.acc_i[slot].w += transparent ? zmax((1.f - avg) * split_weight, 0.f) : split_weight;
}

// ---- stage kernels ---------------------------------------------------------------------------

// Worker.render per sample: Sensor.cameraSample (sensor.zig:152-166) + Perspective.generateVertex
// (camera_perspective.zig:124-150) + Vertex.init (vertex.zig:67-85)
__global__ void __launch_bounds__(kBlock) generateKernel(ZygpuView view, PathState st, PassParams pass) {
    __shared__ __align__(16) uint32_t sobol_tables[kSobolTableWords];
    loadSobolTables(sobol_tables);
    for (uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x; slot < pass.num_paths; slot += gridDim.x * blockDim.x) {
        const SlotId id = slotId(slot, pass);

        SamplerD sampler;
        sampler.sobol.tables = sobol_tables;
        sampler.use_sobol    = ZYG_SAMPLER_SOBOL == view.sampler;
        seedSamplers(id, pass, view.spp_total, sampler.sobol, sampler.rng);

        const int32_t fr = view.filter_radius_int;
        const int32_t px = int32_t(id.pixel_id % pass.padded_w) - fr;
        const int32_t py = int32_t(id.pixel_id / pass.padded_w) - fr;

        float s4[4];
        if (sampler.use_sobol) {  // sample4D
            if (sampler.sobol.dimension >= 2) sampler.sobol.incrementSeed();
            const uint32_t d       = sampler.sobol.dimension;
            sampler.sobol.dimension = d + 4;
            s4[0] = sampler.sobol.buffer[d];
            s4[1] = sampler.sobol.buffer[d + 1];
            s4[2] = sampler.sobol.buffer[d + 2];
            s4[3] = sampler.sobol.buffer[d + 3];
        } else {
            for (int i = 0; i < 4; ++i) s4[i] = sampler.rng.randomFloat();
        }
        (void)sampler.sample1D();  // shutter time: static scenes
        sampler.incrementPadding();

        const float c0 = float(px) + s4[0];
        const float c1 = float(py) + s4[1];

        const V3 left_top = {view.left_top[0], view.left_top[1], view.left_top[2]};
        const V3 d_x      = {view.d_x[0], view.d_x[1], view.d_x[2]};
        const V3 d_y      = {view.d_y[0], view.d_y[1], view.d_y[2]};

        V3 direction = add3(add3(left_top, scale3(c0, d_x)), scale3(c1, d_y));
        V3 origin;
        if (view.aperture_radius > 0.f) {
            float lx, ly;
            diskConcentric(s4[2], s4[3], lx, ly);  // Aperture.sample, aperture.zig:46-53
            origin         = {lx * view.aperture_radius, ly * view.aperture_radius, 0.f};
            const float t  = __fdiv_rn(view.focus_distance, direction.z);
            const V3 focus = scale3(t, direction);
            direction      = sub3(focus, origin);
        } else {
            origin = {view.eye_offset[0], view.eye_offset[1], view.eye_offset[2]};
        }

        const TrafoD trafo = {{view.camera_trafo.r[0][0], view.camera_trafo.r[0][1], view.camera_trafo.r[0][2]},
                              {view.camera_trafo.r[1][0], view.camera_trafo.r[1][1], view.camera_trafo.r[1][2]},
                              {view.camera_trafo.r[2][0], view.camera_trafo.r[2][1], view.camera_trafo.r[2][2]},
                              {view.camera_trafo.r[0][3], view.camera_trafo.r[1][3], view.camera_trafo.r[2][3]},
                              {view.camera_trafo.position[0], view.camera_trafo.position[1], view.camera_trafo.position[2]}};

        const V3 origin_w    = trafo.objectToWorldPoint(origin);
        const V3 direction_w = trafo.objectToWorldVector(normalize3(direction));

        const uint32_t state = kPrimaryRay | kTransparent | kSingular;
        st.ray_o[slot]  = make_float4(origin_w.x, origin_w.y, origin_w.z, __uint_as_float(packFlags(state, 0, 0)));
        st.ray_d[slot]  = make_float4(direction_w.x, direction_w.y, direction_w.z, kRayMaxT);
        st.thr[slot]    = make_float4(1.f, 1.f, 1.f, 0.f);
        st.prev_p[slot] = make_float4(origin_w.x, origin_w.y, origin_w.z, 0.f);
        st.prev_n[slot] = make_float4(0.f, 0.f, 0.f, 1.f);  // split_weight = 1
        st.acc_e[slot]  = make_float4(0.f, 0.f, 0.f, s4[0]);
        st.acc_d[slot]  = make_float4(0.f, 0.f, 0.f, s4[1]);
        st.acc_i[slot]  = make_float4(0.f, 0.f, 0.f, 0.f);
        if (nullptr != st.aov_misc) {  // aov.Value.clear, worker.zig:155: Depth starts at floatMax, everything else at 0
            st.aov_albedo[slot] = make_float4(0.f, 0.f, 0.f, 0.f);
            st.aov_gn[slot]     = make_float4(0.f, 0.f, 0.f, 0.f);
            st.aov_sn[slot]     = make_float4(0.f, 0.f, 0.f, 0.f);
            st.aov_misc[slot]   = make_float4(0.f, FLT_MAX, 0.f, 0.f);
        }
        storeSampler(st, slot, sampler, kPoolFirst);  // the camera vertex sits in lane 0 (vertex id == slot)
        st.queue_a[slot] = slot;
        if (st.lanes > 1) st.queue_t[slot] = slot;
    }
    if (0 == blockIdx.x && 0 == threadIdx.x) {
        st.counters[0] = pass.num_paths;
        st.counters[1] = 0;
        st.counters[4] = 0;
        st.counters[7]  = pass.num_paths;
        st.counters[9]  = 0;
        st.counters[10] = 0;
        st.counters[11] = 0;
    }
}

__device__ __forceinline__ const ZygpuMaterial& propMaterial(const SceneDevice& sc, uint32_t prop, uint32_t part) {
    return sc.materials[__ldg(sc.material_ids + sc.props[prop].parts_start + part)];
}

// Vertex.iorOutside (vertex.zig:87-93) and Stack.highestPriority (medium.zig:74-82) for Vertex.sample
__device__ __forceinline__ void mediaForSample(const SceneDevice& sc, const MediaD& media, const FragD& frag, V3 wo, float& ior_outside,
                                               int& highest_priority) {
    ior_outside      = 1.f;
    highest_priority = -128;
    if (0 == media.count) return;
    for (uint32_t i = 0; i < media.count; ++i) highest_priority = max(highest_priority, propMaterial(sc, media.prop[i], media.part[i]).priority);
    const uint32_t back = media.count - 1;
    if (frag.sameHemisphere(wo)) {  // Stack.topIor
        ior_outside = propMaterial(sc, media.prop[back], media.part[back]).ior;
    } else if (media.count > 1) {  // Stack.peekIor
        const uint32_t i = (media.prop[back] == frag.prop && media.part[back] == frag.part) ? back - 1 : back;
        ior_outside      = propMaterial(sc, media.prop[i], media.part[i]).ior;
    }
}

// Pool.maxSplits, vertex.zig:306-309
__device__ __forceinline__ uint32_t maxSplits(uint32_t path_count_log2, bool primary_ray, uint32_t depth) {
    const uint32_t m = 4u >> path_count_log2;
    return m - (primary_ray ? 0u : min(depth, m - 1u));
}

struct LoadedVertex {
    VertexD  v;
    V3       throughput;
    float    reg_alpha;
    float    split_weight;
    uint32_t vertex_depth;
    uint32_t path_count_log2, num_media;
    HitD     isec;
    uint32_t prop;
};

__device__ __forceinline__ LoadedVertex loadVertex(const PathState& st, uint32_t slot /* vertex id */) {
    const float4 o  = st.ray_o[slot];
    const float4 d  = st.ray_d[slot];
    const float4 t  = st.thr[slot];
    const float4 pp = st.prev_p[slot];
    const float4 pn = st.prev_n[slot];
    const float4 h  = st.hit[slot];

    LoadedVertex    r;
    const uint32_t  flags = __float_as_uint(o.w);
    r.v.ray               = makeRay({o.x, o.y, o.z}, {d.x, d.y, d.z}, 0.f, d.w);
    r.v.origin            = {pp.x, pp.y, pp.z};
    r.v.geo_n             = {pn.x, pn.y, pn.z};
    r.v.bxdf_pdf          = t.w;
    r.v.light_split_threshold = 0.f;  // set by the stages before it is read
    r.split_weight        = pn.w;
    r.v.state             = flags & 0xffu;
    r.v.probe_depth       = (flags >> 8) & 0xffu;
    r.vertex_depth        = (flags >> 16) & 0xffu;
    r.path_count_log2     = (flags >> 24) & 3u;
    r.num_media           = (flags >> 26) & 3u;
    r.throughput          = {t.x, t.y, t.z};
    r.reg_alpha           = pp.w;
    r.isec                = {d.w, h.x, h.y, __float_as_uint(h.z)};
    r.prop                = __float_as_uint(h.w);
    return r;
}

// PathtracerMIS.li up to the shadow rays: connectLight (pathtracer_mis.zig:280-341), termination (:76-86), Russian
// roulette (:88, helper.zig:75-89), Vertex.sample (:93), sampleLights / evaluateLight up to the visibility test
// (:174-250).
// Features the scene needs of shade_a; what it does not need is compiled out (each costs registers in the hottest kernel).
// kFeatureTextured: a material reads image maps per vertex (colour / roughness / metallic / normal): instances without it hold none of
// that code (as a run-time branch it cost the map-free scenes 4 % through registers and code size)
enum : uint32_t { kFeatureSplit = 1, kFeatureMeshLights = 2, kFeatureInfiniteLights = 4, kFeatureDeferredLights = 8, kFeatureTextured = 16 };


// hlp.sampleNormal, material_helper.zig:16-79 for a UV-mapped normal map: the tangent-space normal of the map in the shading frame, then the
// adaption that keeps the reflection of `wo` above the geometric surface.
__device__ __forceinline__ V3 sampleNormal(V3 wo, V3 t, V3 b, V3 n, V3 geo_n, float nx, float ny) {
    const float nz = __fsqrt_rn(zmax(1.f - (nx * nx + ny * ny), 0.01f));
    // rs.tangentToWorld(nm), renderstate.zig:51-58
    const V3 w  = {(nx * t.x + ny * b.x) + nz * n.x, (nx * t.y + ny * b.y) + nz * n.y, (nx * t.z + ny * b.z) + nz * n.z};
    const V3 nn = normalize3(w);

    const V3    r = sub3(scale3(2.f * dot3(wo, nn), nn), wo);  // math.reflect3(n, wo), vector4.zig:94-96
    const float a = dot3(geo_n, r);
    if (a >= 0.f) return nn;
    if (dot3(geo_n, wo) < 0.0017453f) return geo_n;  // cos(89.9 degrees)

    const float bb      = dot3(geo_n, nn);
    const float epsilon = 1e-4f;
    V3          tangent = nn;
    if (bb > epsilon) {
        const float distance = __fdiv_rn(fabsf(a), bb);
        tangent              = normalize3(add3(r, scale3(distance, nn)));
    }
    tangent = add3(tangent, scale3(epsilon, geo_n));
    return normalize3(add3(wo, tangent));
}

__device__ __forceinline__ V3 surfaceMapTexel(const ImageSamplerDevice* is, float u, float v, float r) { return imageTexel(*is, u, v, r); }

// Material.sample with the image maps of a Substitute (substitute_material.zig:114-162): colour, roughness, metallic and the normal map are
// all looked up with the vertex's one stochastic_r (texture_sampler.zig:22-76), so shade_b re-reads the texels shade_a read.
template <bool Split, bool Textured>
__device__ __forceinline__ MatSampleD texturedMaterialSample(const SceneDevice& sc, ZygpuMaterial& m, const FragD& frag, V3 wo,
                                                             float stochastic_r, float reg_weight, float reg_alpha, bool caustics,
                                                             float specular_threshold, float ior_outside, int highest_priority) {
    if (Textured) {
        if (ZYGPU_NULL != m.color_map) {  // ts.sample2D_3(self.color, rs, ...), :120
            const V3 c = imageTexel(sc.image_samplers[m.color_map], frag.u, frag.v, stochastic_r);
            m.color[0] = c.x, m.color[1] = c.y, m.color[2] = c.z;
        }
        if (ZYGPU_NULL != m.roughness_map) m.roughness = surfaceMapTexel(sc.image_samplers + m.roughness_map, frag.u, frag.v, stochastic_r).x;  // :122
        if (ZYGPU_NULL != m.metallic_map) m.metallic = surfaceMapTexel(sc.image_samplers + m.metallic_map, frag.u, frag.v, stochastic_r).x;     // :123
    }
    MatSampleD r = materialSample<Split, Textured>(m, frag, wo, reg_weight, reg_alpha, caustics, specular_threshold, ior_outside, highest_priority);
    if (Textured && ZYGPU_NULL != m.normal_map && kSampleSubstitute == r.kind) {  // :157-159: result.super.frame = Frame.init(n)
        const V3 xy = surfaceMapTexel(sc.image_samplers + m.normal_map, frag.u, frag.v, stochastic_r);
        const V3 n  = sampleNormal(wo, frag.t, frag.b, r.n, r.geo_n, xy.x, xy.y);
        V3       t, b;
        orthonormalBasis3(n, t, b);
        r.frame = {t, b, n};
    }
    return r;
}

template <uint32_t Features>
__global__ void __launch_bounds__(kBlock, ZYGPU_SHADE_BLOCKS) shadeAKernel(SceneDevice sc, ZygpuView view, PathState st, PassParams pass, uint32_t round) {
    constexpr bool Split      = 0 != (Features & kFeatureSplit);
    constexpr bool MeshLights = 0 != (Features & kFeatureMeshLights);
    constexpr bool Infinite   = 0 != (Features & kFeatureInfiniteLights);
    constexpr bool Deferred   = 0 != (Features & kFeatureDeferredLights);  // light selection and sampling run in the light kernels
    constexpr bool Textured   = 0 != (Features & kFeatureTextured);
    __shared__ __align__(16) uint32_t sobol_tables[kSobolTableWords];
    const bool      later = Split && round > 0;
    const uint32_t  count = later ? st.counters[9] : st.counters[0];
    if (blockIdx.x * blockDim.x >= count) return;  // no item of any iteration falls to this block: skip the 20 KB table load
    loadSobolTables(sobol_tables);
    const uint32_t* __restrict__ queue = later ? st.queue_s : st.queue_a;
    const uint32_t  iters = (count + gridDim.x * blockDim.x - 1) / (gridDim.x * blockDim.x);
    for (uint32_t it = 0; it < iters; ++it) {
        const uint32_t i      = it * gridDim.x * blockDim.x + blockIdx.x * blockDim.x + threadIdx.x;
        bool           alive  = false;
        bool           multi  = false;
        uint32_t       slot   = 0;
        if (i < count) {
            slot            = queue[i];
            const uint4 smp = st.smp[slot];
            uint32_t    pool = smp.w;
            uint32_t    lane = 0;
            bool        mine = true;  // the slot has a vertex for this round
            if (Split) {
                if (0 == round) {
                    pool  = poolSwap(pool);
                    multi = poolCurCount(pool) > 1;
                }
                mine = poolCurCount(pool) > round;
                lane = poolCurLane(pool, round);
            }
            if (mine) {
            const uint32_t vid = lane * st.capacity + slot;
            LoadedVertex lv = loadVertex(st, vid);
            VertexD&     vertex = lv.v;

            const uint32_t total_depth = vertex.probe_depth;
            const bool     hit         = kEnd != lv.prop;

            SamplerD sampler;
            sampler.sobol.tables = sobol_tables;
            loadSampler(st, slot, smp, pass, view.spp_total, total_depth, sampler);
            if (ZYG_SAMPLER_SOBOL != view.sampler) sampler.use_sobol = false;

            FragD frag;
            frag.prop = kEnd;
            if (hit) shapeFragment(sc, lv.prop, vertex.ray, lv.isec, frag);

            // Context.nextEvent inside a medium: VolumeIntegrator.integrate -> propScatter, the "glass" case
            // (volume_integrator.zig:51-66, 84-130): throughput *= exp(-mu_a * d); topCC = the highest-priority medium
            MediaD media{0, {0, 0, 0}, {0, 0, 0}};
            if (Split && 0 != lv.num_media) {
                media = unpackMedia(st.med[vid], lv.num_media);
                if (hit) {
                    int      priority = -128;
                    uint32_t highest  = 0;
                    for (uint32_t m = 0; m < media.count; ++m) {
                        const int lp = propMaterial(sc, media.prop[m], media.part[m]).priority;
                        if (lp >= priority) {
                            priority = lp;
                            highest  = m;
                        }
                    }
                    const ZygpuMaterial& mm = propMaterial(sc, media.prop[highest], media.part[highest]);
                    const float          nd = -(vertex.ray.tmax - vertex.ray.tmin);
                    lv.throughput = mul3(lv.throughput, {expf(nd * mm.color[0]), expf(nd * mm.color[1]), expf(nd * mm.color[2])});
                }
            }

            if (slot == pass.debug_slot) {
                printf("[gpu] depth %u pc %u sw %g state p%d s%d sg%d | hit prop %u t %.9g media %u | thr %.9g %.9g %.9g | rng %llx round %u lane %u\n",
                       total_depth, 1u << lv.path_count_log2, lv.split_weight, int(0 != (vertex.state & kPrimaryRay)),
                       int(0 != (vertex.state & kSpecular)), int(0 != (vertex.state & kSingular)), lv.prop, vertex.ray.tmax, lv.num_media,
                       lv.throughput.x, lv.throughput.y, lv.throughput.z, (unsigned long long)sampler.rng.state, round, lane);
            }

            // connectLight
            V3 this_light = splat3(0.f);
            if (!(0 == view.caustics_path && 0 != (vertex.state & kSpecular) && 0 == (vertex.state & kPrimaryRay))) {
                vertex.light_split_threshold = splitThreshold(view.split_threshold, lv.vertex_depth);
                if (hit) this_light = evaluateRadiance<MeshLights>(sc, vertex, frag, sampler);
                this_light = add3(this_light, unoccludingEmission<MeshLights>(sc, vertex, sampler));
                if (Infinite && kRayMaxT == vertex.ray.tmax) {  // the ray left the scene: infinite props, pathtracer_mis.zig:313-338
                    for (uint32_t k = 0; k < sc.num_infinite_props; ++k) {
                        const uint32_t  entity = __ldg(sc.infinite_props + k);
                        const ZygpuProp iprop  = sc.props[entity];
                        if (!propVisible(iprop.flags, vertex.probe_depth) || !aabbIntersect(sc.aabbs, entity, vertex.ray)) continue;
                        FragD light_frag;
                        light_frag.prop  = entity;
                        light_frag.trafo = loadTrafo(sc.trafos, entity);
                        HitD isec;
                        if (ZYG_SHAPE_DISTANT == iprop.shape) {
                            if (!distantIntersect(vertex.ray, light_frag.trafo, isec)) continue;
                            distantFragment(vertex.ray, isec, light_frag);
                        } else if (ZYG_SHAPE_CANOPY == iprop.shape) {
                            if (!canopyIntersect(vertex.ray, light_frag.trafo, isec)) continue;
                            canopyFragment(vertex.ray, light_frag);
                        } else {
                            continue;
                        }
                        this_light = add3(this_light, evaluateRadiance<MeshLights>(sc, vertex, light_frag, sampler));
                    }
                }
            }

            const V3 split_throughput = scale3(lv.split_weight, lv.throughput);
            ivalueAdd(st, slot, mul3(split_throughput, this_light), total_depth, 2, 0 == total_depth, 0 != (vertex.state & kSingular));

            bool terminate = !hit || vertex.probe_depth >= view.max_depth_surface || 0 >= view.max_depth_volume;
            // pathtracer_mis.zig:75-84: what a see-through path meets at its end covers the background (only the alpha reads this)
            V3 end_throughput = lv.throughput;
            if (terminate && 0 != (vertex.state & kTransparent)) {
                end_throughput = mul3(end_throughput, {1.f - zmin(this_light.x, 1.f), 1.f - zmin(this_light.y, 1.f), 1.f - zmin(this_light.z, 1.f)});
            }

            if (!terminate) {
                // russianRoulette
                const float r  = sampler.sample1D();
                const float mx = hmax3(lv.throughput);
                const float q  = __fdiv_rn(mx, 0.1f);
                if (q < 1.f) {
                    if (r >= q) {
                        terminate = true;
                    } else {
                        lv.throughput = divs3(lv.throughput, q);
                    }
                }
            }

            if (!terminate) {
                const bool caustics = 0 == (vertex.state & kPrimaryRay) ? 0 != view.caustics_path : true;  // causticsResolve, :343-349

                const V3      wo = neg3(vertex.ray.d);
                ZygpuMaterial m  = sc.materials[__ldg(sc.material_ids + sc.props[frag.prop].parts_start + frag.part)];
                const float   stochastic_r = sampler.sample1D();  // rs.stochastic_r, vertex.zig:165
                if (Textured) st.stoch[vid] = stochastic_r;
                float ior_outside      = 1.f;
                int   highest_priority = -128;
                if (Split) mediaForSample(sc, media, frag, wo, ior_outside, highest_priority);
                const MatSampleD mat_sample = texturedMaterialSample<Split, Textured>(sc, m, frag, wo, stochastic_r, view.regularize_roughness, lv.reg_alpha,
                                                                                      caustics, view.specular_threshold, ior_outside, highest_priority);

                // Worker.commonAOV, worker.zig:209-242 (pathtracer_mis.zig:95-97). A view that records AOVs runs the Textured instances.
                if (Textured && nullptr != st.aov_misc) {
                    if (0 != (vertex.state & kPrimaryRay) && mat_sample.can_evaluate) {
                        // MaterialSample.aovAlbedo, material_sample.zig:38-45; substitute_sample.zig:80-86
                        V3 albedo = splat3(0.f);
                        if (kSampleSubstitute == mat_sample.kind) {
                            albedo = {zlerp(mat_sample.albedo.x, mat_sample.f0.x, mat_sample.metallic), zlerp(mat_sample.albedo.y, mat_sample.f0.y, mat_sample.metallic),
                                      zlerp(mat_sample.albedo.z, mat_sample.f0.z, mat_sample.metallic)};
                        } else if (kSampleGlass == mat_sample.kind) {
                            albedo = splat3(1.f);
                        }
                        const V3 a          = mul3(lv.throughput, albedo);
                        st.aov_albedo[slot] = make_float4(a.x, a.y, a.z, 0.f);
                    }
                    if (0 == vertex.probe_depth) {
                        st.aov_gn[slot]   = make_float4(mat_sample.geo_n.x, mat_sample.geo_n.y, mat_sample.geo_n.z, 0.f);
                        st.aov_sn[slot]   = make_float4(mat_sample.frame.z.x, mat_sample.frame.z.y, mat_sample.frame.z.z, 0.f);
                        const uint32_t id = __ldg(sc.material_ids + sc.props[frag.prop].parts_start + frag.part);
                        st.aov_misc[slot] = make_float4(__fsqrt_rn(mat_sample.ax), vertex.ray.tmax, float(1u + id), 0.f);
                    }
                }

                vertex.light_split_threshold = splitThreshold(view.split_threshold, vertex.probe_depth);

                // sampleLights. Every path owns `shadow_stride` consecutive shadow records (slot * stride + k): picks and
                // their samples are generated in the reference's order and stored in that order.
                uint32_t num_records = 0;
                bool     request     = false;
                if (Deferred && mat_sample.can_evaluate) {
                    const float    select = sampler.sample1D();
                    const uint32_t bits   = (mat_sample.translucent ? 1u : 0u) | (dot3(mat_sample.geo_n, frag.geo_n) < 0.f ? 2u : 0u) |
                                          (vertex.light_split_threshold != view.split_threshold ? 4u : 0u) | (total_depth << 8);
                    st.ls_p[slot] = make_float4(frag.p.x, frag.p.y, frag.p.z, select);
                    st.ls_g[slot] = make_float4(frag.geo_n.x, frag.geo_n.y, frag.geo_n.z, __uint_as_float(bits));
                    request       = true;
                }
                if (!Deferred && mat_sample.can_evaluate) {
                    const V3    p           = frag.p;
                    const V3    n           = mat_sample.geo_n;
                    const bool  translucent = mat_sample.translucent;
                    const float select      = sampler.sample1D();

                    // the picks first, then one loop over them: with the sampling code inside the callback the compiler keeps one
                    // out-of-line copy of it per call site of Tree.randomLight and a closure of references, which moves the scene and
                    // path records into local memory (measured: shade_a of the instanced scene 0.8 -> 1.6 ms per launch)
                    LightPickD     picks[kMaxLightPicks];
                    uint32_t       num_picks = 0;
                    lightTreeRandomLight(sc, p, n, translucent, select, vertex.light_split_threshold, [&](LightPickD pick) {
                        if (num_picks < kMaxLightPicks) picks[num_picks++] = pick;
                    });
                    for (uint32_t pi = 0; pi < num_picks; ++pi) {
                        const LightPickD pick = picks[pi];
                        const ZygpuLight light = sc.lights[pick.offset];
                        const TrafoD     trafo = loadTrafo(sc.trafos, light.prop);
                        const uint32_t   shape = sc.props[light.prop].shape;
                        if (Infinite && ZYG_SHAPE_DISTANT == shape) {  // Distant.sampleTo, distant.zig:78-107
                            const float radius = trafo.scale.x;
                            if (radius <= 0.f) continue;
                            float u0, u1;
                            sampler.sample2D(u0, u1);
                            float lx, ly;
                            diskConcentric(u0, u1, lx, ly);
                            const V3 ws  = scale3(radius, trafo.transformVector({lx, ly, 0.f}));
                            const V3 dir = normalize3(sub3(ws, trafo.r2));
                            if (dot3(dir, n) <= 0.f && !translucent) continue;
                            if (num_records < st.shadow_stride) {
                                const size_t rec    = size_t(slot) * st.shadow_stride + num_records;
                                const V3     origin = frag.offsetP(dir);
                                const float  pdf    = __fdiv_rn(1.f, distantSolidAngle(radius));
                                st.sh_o[rec]  = make_float4(origin.x, origin.y, origin.z, pdf * pick.pdf);
                                st.sh_p[rec]  = make_float4(0.f, 0.f, 0.f, __uint_as_float(pick.offset | 0x80000000u));
                                st.sh_wi[rec] = make_float4(dir.x, dir.y, dir.z, 0.f);
                                num_records += 1;
                            } else {
                                st.counters[3] = 1;
                            }
                            continue;
                        }
                        if (Infinite && ZYG_SHAPE_CANOPY == shape) {  // Canopy.sampleMaterialTo, canopy.zig:94-131
                            if (ZYG_LIGHT_PROP_IMAGE != light.light_class) continue;
                            float u0, u1;
                            sampler.sample2D(u0, u1);
                            V3    dir;
                            float su, sv, pdf;
                            if (!canopySampleMaterialTo(sc.image_samplers[light.sampler], trafo, n, translucent, u0, u1, dir, su, sv, pdf)) continue;
                            if (num_records < st.shadow_stride) {
                                const size_t rec    = size_t(slot) * st.shadow_stride + num_records;
                                const V3     origin = frag.offsetP(dir);
                                st.sh_o[rec]  = make_float4(origin.x, origin.y, origin.z, pdf * pick.pdf);
                                st.sh_p[rec]  = make_float4(su, sv, 0.f, __uint_as_float(pick.offset | 0x80000000u));  // uvw of the sample
                                st.sh_wi[rec] = make_float4(dir.x, dir.y, dir.z, 0.f);
                                num_records += 1;
                            } else {
                                st.counters[3] = 1;
                            }
                            continue;
                        }
                        if (MeshLights && ZYG_SHAPE_TRIANGLE_MESH == shape && ZYGPU_NULL != light.sampler) {
                            num_records = meshLightSampleTo(sc, st, slot, light, pick, trafo, frag, n, translucent,
                                                            vertex.light_split_threshold, sampler, num_records);
                            continue;
                        }
                        if (ZYG_SHAPE_SPHERE == shape) {  // Sphere.sampleTo, sphere.zig:323-393
                            SphereLightD sl;
                            sl.init(trafo, p);
                            if (!sl.valid) continue;
                            const uint32_t ns = lightNumSamples(light, vertex.light_split_threshold);
                            for (uint32_t k = 0; k < ns; ++k) {
                                float u0, u1;
                                sampler.sample2D(u0, u1);
                                V3    lp, wn, dir;
                                float pdf;
                                if (!sl.sample(trafo, p, n, translucent, u0, u1, lp, wn, dir, pdf)) continue;
                                if (num_records < st.shadow_stride) {
                                    const size_t rec       = size_t(slot) * st.shadow_stride + num_records;
                                    const V3     origin    = frag.offsetP(dir);
                                    const V3     light_pos = offsetRay(lp, wn);
                                    st.sh_o[rec]  = make_float4(origin.x, origin.y, origin.z, (float(ns) * pdf) * pick.pdf);
                                    st.sh_p[rec]  = make_float4(light_pos.x, light_pos.y, light_pos.z, __uint_as_float(pick.offset));
                                    st.sh_wi[rec] = make_float4(dir.x, dir.y, dir.z, 0.f);
                                    num_records += 1;
                                } else {
                                    st.counters[3] = 1;
                                }
                            }
                            continue;
                        }
                        if (Textured && ZYG_SHAPE_DISK == shape) {  // Disk.sampleTo, disk.zig:252-332
                            DiskLightD dl;
                            dl.init(trafo, p);
                            if (!dl.valid) continue;
                            const uint32_t ns = lightNumSamples(light, vertex.light_split_threshold);
                            for (uint32_t k = 0; k < ns; ++k) {
                                V3    lp, wn, dir;
                                float pdf;
                                if (!dl.sample(trafo, p, n, 0 != light.two_sided, translucent, float(ns), sampler, lp, wn, dir, pdf)) continue;
                                if (num_records < st.shadow_stride) {
                                    const size_t rec       = size_t(slot) * st.shadow_stride + num_records;
                                    const V3     origin    = frag.offsetP(dir);
                                    const V3     light_pos = offsetRay(lp, wn);
                                    st.sh_o[rec]  = make_float4(origin.x, origin.y, origin.z, pdf * pick.pdf);
                                    st.sh_p[rec]  = make_float4(light_pos.x, light_pos.y, light_pos.z, __uint_as_float(pick.offset));
                                    st.sh_wi[rec] = make_float4(dir.x, dir.y, dir.z, 0.f);
                                    num_records += 1;
                                } else {
                                    st.counters[3] = 1;
                                }
                            }
                            continue;
                        }
                        if (ZYG_SHAPE_RECTANGLE == shape && ZYG_LIGHT_PROP_IMAGE == light.light_class) {  // Rectangle.sampleMaterialTo
                            const uint32_t            ns   = lightNumSamples(light, vertex.light_split_threshold);
                            const ImageSamplerDevice& is   = sc.image_samplers[light.sampler];
                            const float               area = trafo.scale.x * trafo.scale.y;
                            for (uint32_t k = 0; k < ns; ++k) {
                                float u0, u1;
                                sampler.sample2D(u0, u1);
                                float su, sv, rs_pdf;
                                imageSample(is, u0, u1, su, sv, rs_pdf);
                                if (0.f == rs_pdf) continue;
                                const V3 ws   = trafo.objectToWorldPoint({-1.f * su + 0.5f, -1.f * sv + 0.5f, 0.f});
                                const V3 axis = sub3(ws, p);
                                V3       wn   = trafo.r2;
                                if (0 != light.two_sided && dot3(wn, axis) > 0.f) wn = neg3(wn);
                                const float sl  = squaredLength3(axis);
                                const float t   = __fsqrt_rn(sl);
                                const V3    dir = divs3(axis, t);
                                const float c   = -dot3(wn, dir);
                                if (c < kDotMin || (dot3(dir, n) <= 0.f && !translucent)) continue;
                                if (num_records < st.shadow_stride) {
                                    const size_t rec       = size_t(slot) * st.shadow_stride + num_records;
                                    const V3     origin    = frag.offsetP(dir);
                                    const V3     light_pos = offsetRay(ws, wn);
                                    st.sh_o[rec]  = make_float4(origin.x, origin.y, origin.z, __fdiv_rn(float(ns) * rs_pdf * sl, c * area) * pick.pdf);
                                    st.sh_p[rec]  = make_float4(light_pos.x, light_pos.y, light_pos.z, __uint_as_float(pick.offset));
                                    st.sh_wi[rec] = make_float4(dir.x, dir.y, dir.z, 0.f);
                                    if (nullptr != st.sh_uv) st.sh_uv[rec] = make_float2(su, sv);
                                    num_records += 1;
                                } else {
                                    st.counters[3] = 1;
                                }
                            }
                            continue;
                        }
                        if (ZYG_SHAPE_RECTANGLE != shape) continue;

                        // Rectangle.sampleTo, rectangle.zig:305-357
                        const uint32_t ns  = lightNumSamples(light, vertex.light_split_threshold);
                        const float    nsf = float(ns);
                        SphQuadD       squad;
                        squad.init(trafo.scale, trafo.worldToFramePoint(p));
                        const float sample_pdf = nsf * squad.pdf(trafo.scale);

                        for (uint32_t k = 0; k < ns; ++k) {
                            float u0, u1;
                            sampler.sample2D(u0, u1);

                            const V3 ls  = squad.sample(u0, u1);
                            const V3 ws  = trafo.frameToWorldPoint(ls);
                            const V3 dir = normalize3(sub3(ws, p));

                            V3 wn = trafo.r2;
                            if (0 != light.two_sided && dot3(wn, dir) > 0.f) wn = neg3(wn);

                            if (-dot3(wn, dir) < kDotMin || 0.f == squad.S || (dot3(dir, n) <= 0.f && !translucent)) continue;

                            if (num_records < st.shadow_stride) {
                                const size_t rec       = size_t(slot) * st.shadow_stride + num_records;
                                const V3     origin    = frag.offsetP(dir);
                                const V3     light_pos = offsetRay(ws, wn);  // Shape.shadowRay, shape.zig:401-416
                                st.sh_o[rec]  = make_float4(origin.x, origin.y, origin.z, sample_pdf * pick.pdf);
                                st.sh_p[rec]  = make_float4(light_pos.x, light_pos.y, light_pos.z, __uint_as_float(pick.offset));
                                st.sh_wi[rec] = make_float4(dir.x, dir.y, dir.z, 0.f);
                                num_records += 1;
                            } else {
                                st.counters[3] = 1;  // more light samples than the host reserved: reported by zygpu_render
                            }
                        }
                    }
                }

                st.sh_n[slot] = num_records;
                st.thr[vid]   = make_float4(lv.throughput.x, lv.throughput.y, lv.throughput.z, vertex.bxdf_pdf);
                alive         = true;
                if (nullptr != st.queue_r && 0 != num_records) {
                    const uint32_t base = atomicAdd(&st.counters[10], num_records);
                    for (uint32_t k = 0; k < num_records; ++k) st.queue_r[base + k] = slot * st.shadow_stride + k;
                }
                if (Deferred && request) st.queue_l[atomicAdd(&st.counters[11], 1u)] = slot;
            }
            if (Split && !alive) pool = poolFree(pool, lane);
            if (!alive && 0 != view.alpha_transparency) alphaAdd(st, slot, 0 != (vertex.state & kTransparent), end_throughput, lv.split_weight);
            storeSampler(st, slot, sampler, pool);
            }
        }
        queuePush(st.queue_b, &st.counters[1], alive, slot);
        if (Split && 0 == round) queuePush(st.queue_s, &st.counters[9], multi, slot);
    }
}

// ---- deferred light sampling ---------------------------------------------------------------------------------------
//
// With many lights the work of PathtracerMIS.sampleLights varies wildly between vertices: far from the lights the tree is
// descended once, near them the adaptive split returns dozens of picks, and a picked mesh light descends its own tree. Inside
// shade_a a warp would wait for its slowest lane (measured: 2.8 of 32 lanes active). So shade_a only leaves a request, and
// two persistent kernels whose lanes fetch new work as soon as they finish do the rest:
//
//   lightSelectPersistent   Tree.randomLight (light_tree.zig:346-447), one tree node per step and lane -> picks
//   lightSamplePersistent   Light.sampleTo for one pick per step and lane, in pick order -> shadow records
//
// The sampler is only touched by the second kernel, pick by pick in the reference's order, so the stream of a vertex is the
// one the inline path produces.

__device__ __forceinline__ V3 offsetPoint(V3 p, V3 geo_n, V3 w) {  // Fragment.offsetP with offset() == 0, intersection.zig:112-116
    const V3 nn = dot3(geo_n, w) > 0.f ? geo_n : neg3(geo_n);
    return offsetRay(fmas3(0.f, nn, p), nn);
}

__global__ void __launch_bounds__(128) lightSelectPersistent(SceneDevice sc, ZygpuView view, PathState st, uint32_t* __restrict__ work_counter) {
    constexpr uint32_t kFull   = 0xffffffffu;
    const uint32_t     lane    = threadIdx.x & 31u;
    const uint32_t     n_items = st.counters[11];
    const TreeD        tr      = sceneTree(sc);

    struct Value {
        float    pdf, random;
        uint32_t node, depth;
    };

    bool     active = false, exhausted = false;
    uint32_t slot = 0, num_picks = 0, end = 0;
    V3       p = {0.f, 0.f, 0.f}, n = {0.f, 0.f, 0.f};
    bool     total_sphere = false;
    float    threshold    = 0.f;
    Value    t{0.f, 0.f, 0, 0};
    Value    stack[12];

    const uint32_t max_split_depth = sc.lt_max_split_depth;

    for (;;) {
        const uint32_t idle = __ballot_sync(kFull, !active);
        if (0 != idle && !exhausted) {
            uint32_t base = 0;
            if (lane == uint32_t(__ffs(int(idle))) - 1u) base = atomicAdd(work_counter, uint32_t(__popc(idle)));
            base = __shfl_sync(kFull, base, __ffs(int(idle)) - 1);
            if (base + uint32_t(__popc(idle)) >= n_items) exhausted = true;
            const uint32_t index = base + uint32_t(__popc(idle & ((1u << lane) - 1u)));
            if (!active && index < n_items) {
                slot              = st.queue_l[index];
                const float4 lp   = st.ls_p[slot];
                const float4 lg   = st.ls_g[slot];
                const uint32_t fl = __float_as_uint(lg.w);
                p                 = {lp.x, lp.y, lp.z};
                n                 = 0 != (fl & 2u) ? V3{-lg.x, -lg.y, -lg.z} : V3{lg.x, lg.y, lg.z};
                total_sphere      = 0 != (fl & 1u);
                threshold         = 0 != (fl & 4u) ? kLowThreshold : view.split_threshold;
                const float random = lp.w;
                num_picks          = 0;
                end                = 0;

                // Tree.randomLight up to the descent, light_tree.zig:353-381
                float      ip    = 0.f;
                const bool split = threshold > 0.f;
                bool       done  = false;
                if (split && sc.lt_num_infinite < kMaxLightPicks - 1) {
                    for (uint32_t i = 0; i < sc.lt_num_infinite; ++i) {
                        st.picks[size_t(slot) * kMaxLightPicks + num_picks++] = make_uint2(__ldg(sc.lt_mapping + i), __float_as_uint(1.f));
                    }
                } else {
                    ip = sc.lt_infinite_weight;
                    if (random < sc.lt_infinite_guard) {
                        const uint32_t l  = dist1dSample(sc.lt_infinite_cdf, sc.lt_num_infinite + 1, random);
                        const float    lp = __ldg(sc.lt_infinite_cdf + l + 1) - __ldg(sc.lt_infinite_cdf + l);
                        st.picks[size_t(slot) * kMaxLightPicks + num_picks++] = make_uint2(__ldg(sc.lt_mapping + l), __float_as_uint(lp * ip));
                        done = true;
                    }
                }
                if (done || 0 == sc.lt_num_nodes) {
                    st.pick_n[slot] = num_picks;
                } else {
                    const float pd = 1.f - ip;
                    t              = {pd, __fdiv_rn(random - ip, pd), 0, split ? 0 : max_split_depth};
                    stack[end++]   = t;
                    active         = true;
                }
            }
        }
        if (0 == __ballot_sync(kFull, active)) {
            if (exhausted) break;
            continue;
        }
        if (active) {  // one iteration of the descent loop, light_tree.zig:396-444
            const LightNodeD node = loadLightNode(tr, t.node);
            if (0 != (node.meta & 1u)) {
                const bool     do_split = t.depth < max_split_depth && lightNodeSplit(node, p, threshold);
                const uint32_t c0       = node.meta >> 2;
                const uint32_t c1       = c0 + 1;
                if (do_split) {
                    t.depth += 1;
                    t.node       = c0;
                    stack[end++] = {t.pdf, t.random, c1, t.depth};
                } else {
                    t.depth = max_split_depth;

                    float p0 = lightNodeWeight(loadLightNode(tr, c0), p, n, total_sphere);
                    float p1 = lightNodeWeight(loadLightNode(tr, c1), p, n, total_sphere);

                    const float pt = p0 + p1;
                    if (0.f == pt) {
                        t = stack[--end];
                    } else {
                        p0 = __fdiv_rn(p0, pt);
                        p1 = __fdiv_rn(p1, pt);
                        if (t.random < p0) {
                            t.node = c0;
                            t.pdf *= p0;
                            t.random = __fdiv_rn(t.random, p0);
                        } else {
                            t.node = c1;
                            t.pdf *= p1;
                            t.random = zmin(__fdiv_rn(t.random - p0, p1), 1.f);
                        }
                    }
                }
            } else {
                const LightPickD pick = lightNodeRandomLight(sc, tr, node, p, n, total_sphere, t.random);
                if (pick.pdf > 0.f && num_picks < kMaxLightPicks) {
                    st.picks[size_t(slot) * kMaxLightPicks + num_picks++] = make_uint2(pick.offset, __float_as_uint(pick.pdf * t.pdf));
                }
                t = stack[--end];
            }
            if (0 == end) {  // `while (!stack.empty())`: the entry pushed first is popped last
                st.pick_n[slot] = num_picks;
                active          = false;
            }
        }
    }
}

// The kernel is long and branchy and its warps mostly wait for instruction fetch (ncu: no_instruction is the top stall, issue
// slots 13 % busy at 122 registers / 4 blocks per SM): more resident warps hide that better than registers help, so it is
// compiled for more blocks per SM than its registers ask for (8 blocks = 64 registers: 310 -> 265 ms per 4-spp pass of config 4, 12 and
// 16 blocks are slower). Since the mesh-light walks run in their own short loop the spills of 64 registers cost more than the two
// blocks give: 6 blocks = 80 registers (288.5 -> 280.8 ms per 8-spp frame).
#ifndef ZYGPU_LIGHT_BLOCKS
#define ZYGPU_LIGHT_BLOCKS 6
#endif
#ifndef ZYGPU_WALK_REFILL
#define ZYGPU_WALK_REFILL 8
#endif
#ifndef ZYGPU_WALK_LEAF_RATIO
#define ZYGPU_WALK_LEAF_RATIO 3u  // leaf iterations are taken when this many times more walks wait at a leaf than at an inner node
#endif
// Scenes with infinite lights run it twice: their picks come first in a vertex's list (Tree.randomLight, light_tree.zig:353-371), so
// phase 1 takes the Distant / Canopy picks of every vertex and phase 2 the finite ones, each on the sampler state the other left.
// A warp then works on one class of light at a time (and each instance holds half the code): lanes that fetch a new vertex no longer
// start with a sky sample while their neighbours are inside a mesh light. Phase 0 = every pick in one launch.
template <bool MeshLights, bool Infinite, int Phase>
__global__ void __launch_bounds__(128, ZYGPU_LIGHT_BLOCKS) lightSamplePersistent(SceneDevice sc, ZygpuView view, PathState st, PassParams pass,
                                                             uint32_t* __restrict__ work_counter) {
    constexpr bool kInfinitePicks = Infinite && 2 != Phase;  // this instance handles Distant / Canopy picks
    constexpr bool kFinitePicks   = 1 != Phase;
    __shared__ __align__(16) uint32_t sobol_tables[kSobolTableWords];
    loadSobolTables(sobol_tables);

    constexpr uint32_t kFull   = 0xffffffffu;
    const uint32_t     lane    = threadIdx.x & 31u;
    const uint32_t     n_items = st.counters[11];

    bool     active = false, exhausted = false;
    uint32_t slot = 0, pick_i = 0, pick_count = 0, num_records = 0;
    V3       p = {0.f, 0.f, 0.f}, geo_n = {0.f, 0.f, 0.f}, n = {0.f, 0.f, 0.f};
    bool     translucent = false;
    float    threshold   = 0.f;
    uint32_t pool_word   = 0;
    SamplerD sampler;
    sampler.sobol.tables = sobol_tables;

    // a triangle-mesh pick in progress: PrimitiveTree.randomLight advances by one iteration per step of this loop
    constexpr bool kWalks   = kFinitePicks && MeshLights;
    bool           walking = false, walk_leaf = false;
    PrimitiveWalkD walk;
    LightPickD     walk_pick{0, 0.f};
    V3             walk_op = {0.f, 0.f, 0.f}, walk_on = {0.f, 0.f, 0.f};

    for (;;) {
        const uint32_t idle = __ballot_sync(kFull, !active);
        if (0 != idle && !exhausted) {
            uint32_t base = 0;
            if (lane == uint32_t(__ffs(int(idle))) - 1u) base = atomicAdd(work_counter, uint32_t(__popc(idle)));
            base = __shfl_sync(kFull, base, __ffs(int(idle)) - 1);
            if (base + uint32_t(__popc(idle)) >= n_items) exhausted = true;
            const uint32_t index = base + uint32_t(__popc(idle & ((1u << lane) - 1u)));
            if (!active && index < n_items) {
                slot              = st.queue_l[index];
                const float4 lp   = st.ls_p[slot];
                const float4 lg   = st.ls_g[slot];
                const uint32_t fl = __float_as_uint(lg.w);
                p                 = {lp.x, lp.y, lp.z};
                geo_n             = {lg.x, lg.y, lg.z};
                n                 = 0 != (fl & 2u) ? neg3(geo_n) : geo_n;
                translucent       = 0 != (fl & 1u);
                threshold         = 0 != (fl & 4u) ? kLowThreshold : view.split_threshold;
                const uint32_t pn = st.pick_n[slot];
                pick_i            = 2 == Phase ? pn >> 16 : 0u;  // phase 1 leaves the index of the first finite pick there
                pick_count        = pn & 0xffffu;
                num_records       = 2 == Phase ? st.sh_n[slot] : 0u;
                const uint4 smp   = st.smp[slot];
                pool_word         = smp.w;
                loadSampler(st, slot, smp, pass, view.spp_total, (fl >> 8) & 0xffu, sampler);
                if (ZYG_SAMPLER_SOBOL != view.sampler) sampler.use_sobol = false;
                active = true;
            }
        }
        if (0 == __ballot_sync(kFull, active)) {
            if (exhausted) break;
            continue;
        }
        if (active && !(kWalks && walking) && pick_i < pick_count) {  // Light.sampleTo for one pick, pathtracer_mis.zig:214-250
            const uint2      pk = st.picks[size_t(slot) * kMaxLightPicks + pick_i];
            const LightPickD pick{pk.x, __uint_as_float(pk.y)};
            pick_i += 1;

            const ZygpuLight light = sc.lights[pick.offset];
            const TrafoD     trafo = loadTrafo(sc.trafos, light.prop);
            const uint32_t   shape = sc.props[light.prop].shape;
            if (1 == Phase && ZYG_SHAPE_DISTANT != shape && ZYG_SHAPE_CANOPY != shape) {
                // the first finite pick ends phase 1: phase 2 resumes here
                pick_i -= 1;
                st.pick_n[slot] = pick_count | (pick_i << 16);
                st.sh_n[slot]   = num_records;
                storeSampler(st, slot, sampler, pool_word);
                active = false;
            } else if (kInfinitePicks && ZYG_SHAPE_DISTANT == shape) {  // Distant.sampleTo, distant.zig:78-107
                const float radius = trafo.scale.x;
                if (radius > 0.f) {
                    float u0, u1;
                    sampler.sample2D(u0, u1);
                    float lx, ly;
                    diskConcentric(u0, u1, lx, ly);
                    const V3 ws  = scale3(radius, trafo.transformVector({lx, ly, 0.f}));
                    const V3 dir = normalize3(sub3(ws, trafo.r2));
                    if (!(dot3(dir, n) <= 0.f && !translucent)) {
                        if (num_records < st.shadow_stride) {
                            const size_t rec    = size_t(slot) * st.shadow_stride + num_records;
                            const V3     origin = offsetPoint(p, geo_n, dir);
                            const float  pdf    = __fdiv_rn(1.f, distantSolidAngle(radius));
                            st.sh_o[rec]  = make_float4(origin.x, origin.y, origin.z, pdf * pick.pdf);
                            st.sh_p[rec]  = make_float4(0.f, 0.f, 0.f, __uint_as_float(pick.offset | 0x80000000u));
                            st.sh_wi[rec] = make_float4(dir.x, dir.y, dir.z, 0.f);
                            num_records += 1;
                        } else {
                            st.counters[3] = 1;
                        }
                    }
                }
            } else if (kInfinitePicks && ZYG_SHAPE_CANOPY == shape) {  // Canopy.sampleMaterialTo, canopy.zig:94-131
                if (ZYG_LIGHT_PROP_IMAGE == light.light_class) {
                    float u0, u1;
                    sampler.sample2D(u0, u1);
                    V3    dir;
                    float su, sv, pdf;
                    if (canopySampleMaterialTo(sc.image_samplers[light.sampler], trafo, n, translucent, u0, u1, dir, su, sv, pdf)) {
                        if (num_records < st.shadow_stride) {
                            const size_t rec    = size_t(slot) * st.shadow_stride + num_records;
                            const V3     origin = offsetPoint(p, geo_n, dir);
                            st.sh_o[rec]  = make_float4(origin.x, origin.y, origin.z, pdf * pick.pdf);
                            st.sh_p[rec]  = make_float4(su, sv, 0.f, __uint_as_float(pick.offset | 0x80000000u));
                            st.sh_wi[rec] = make_float4(dir.x, dir.y, dir.z, 0.f);
                            num_records += 1;
                        } else {
                            st.counters[3] = 1;
                        }
                    }
                }
            } else if (kFinitePicks && MeshLights && ZYG_SHAPE_TRIANGLE_MESH == shape && ZYGPU_NULL != light.sampler) {
                // Mesh.sampleTo, triangle_mesh.zig:492-608: the draw that picks the triangles, then the walk below
                walk_pick = pick;
                walk_op   = trafo.worldToObjectPoint(p);
                walk_on   = trafo.worldToObjectNormal(n);
                walk.start(sampler.sample1D(), threshold);
                walking   = true;
                walk_leaf = lightNodeIsLeaf(primitiveTree(sc.mesh_samplers[light.sampler]), 0);
            } else if (kFinitePicks && ZYG_SHAPE_SPHERE == shape) {  // Sphere.sampleTo, sphere.zig:323-393
                SphereLightD sl;
                sl.init(trafo, p);
                const uint32_t ns = sl.valid ? lightNumSamples(light, threshold) : 0;
                for (uint32_t k = 0; k < ns; ++k) {
                    float u0, u1;
                    sampler.sample2D(u0, u1);
                    V3    lp, wn, dir;
                    float pdf;
                    if (!sl.sample(trafo, p, n, translucent, u0, u1, lp, wn, dir, pdf)) continue;
                    if (num_records < st.shadow_stride) {
                        const size_t rec       = size_t(slot) * st.shadow_stride + num_records;
                        const V3     origin    = offsetPoint(p, geo_n, dir);
                        const V3     light_pos = offsetRay(lp, wn);
                        st.sh_o[rec]  = make_float4(origin.x, origin.y, origin.z, (float(ns) * pdf) * pick.pdf);
                        st.sh_p[rec]  = make_float4(light_pos.x, light_pos.y, light_pos.z, __uint_as_float(pick.offset));
                        st.sh_wi[rec] = make_float4(dir.x, dir.y, dir.z, 0.f);
                        num_records += 1;
                    } else {
                        st.counters[3] = 1;
                    }
                }
            } else if (kFinitePicks && ZYG_SHAPE_DISK == shape) {  // Disk.sampleTo, disk.zig:252-332
                DiskLightD dl;
                dl.init(trafo, p);
                const uint32_t ns = dl.valid ? lightNumSamples(light, threshold) : 0;
                for (uint32_t k = 0; k < ns; ++k) {
                    V3    lp, wn, dir;
                    float pdf;
                    if (!dl.sample(trafo, p, n, 0 != light.two_sided, translucent, float(ns), sampler, lp, wn, dir, pdf)) continue;
                    if (num_records < st.shadow_stride) {
                        const size_t rec       = size_t(slot) * st.shadow_stride + num_records;
                        const V3     origin    = offsetPoint(p, geo_n, dir);
                        const V3     light_pos = offsetRay(lp, wn);
                        st.sh_o[rec]  = make_float4(origin.x, origin.y, origin.z, pdf * pick.pdf);
                        st.sh_p[rec]  = make_float4(light_pos.x, light_pos.y, light_pos.z, __uint_as_float(pick.offset));
                        st.sh_wi[rec] = make_float4(dir.x, dir.y, dir.z, 0.f);
                        num_records += 1;
                    } else {
                        st.counters[3] = 1;
                    }
                }
            } else if (kFinitePicks && ZYG_SHAPE_RECTANGLE == shape && ZYG_LIGHT_PROP_IMAGE == light.light_class) {  // Rectangle.sampleMaterialTo
                const uint32_t            ns   = lightNumSamples(light, threshold);
                const ImageSamplerDevice& is   = sc.image_samplers[light.sampler];
                const float               area = trafo.scale.x * trafo.scale.y;
                for (uint32_t k = 0; k < ns; ++k) {
                    float u0, u1;
                    sampler.sample2D(u0, u1);
                    float su, sv, rs_pdf;
                    imageSample(is, u0, u1, su, sv, rs_pdf);
                    if (0.f == rs_pdf) continue;
                    const V3 ws   = trafo.objectToWorldPoint({-1.f * su + 0.5f, -1.f * sv + 0.5f, 0.f});
                    const V3 axis = sub3(ws, p);
                    V3       wn   = trafo.r2;
                    if (0 != light.two_sided && dot3(wn, axis) > 0.f) wn = neg3(wn);
                    const float sl  = squaredLength3(axis);
                    const float t   = __fsqrt_rn(sl);
                    const V3    dir = divs3(axis, t);
                    const float c   = -dot3(wn, dir);
                    if (c < kDotMin || (dot3(dir, n) <= 0.f && !translucent)) continue;
                    if (num_records < st.shadow_stride) {
                        const size_t rec       = size_t(slot) * st.shadow_stride + num_records;
                        const V3     origin    = offsetPoint(p, geo_n, dir);
                        const V3     light_pos = offsetRay(ws, wn);
                        st.sh_o[rec]  = make_float4(origin.x, origin.y, origin.z, __fdiv_rn(float(ns) * rs_pdf * sl, c * area) * pick.pdf);
                        st.sh_p[rec]  = make_float4(light_pos.x, light_pos.y, light_pos.z, __uint_as_float(pick.offset));
                        st.sh_wi[rec] = make_float4(dir.x, dir.y, dir.z, 0.f);
                        if (nullptr != st.sh_uv) st.sh_uv[rec] = make_float2(su, sv);
                        num_records += 1;
                    } else {
                        st.counters[3] = 1;
                    }
                }
            } else if (kFinitePicks && ZYG_SHAPE_RECTANGLE == shape) {  // Rectangle.sampleTo, rectangle.zig:305-357
                const uint32_t ns  = lightNumSamples(light, threshold);
                const float    nsf = float(ns);
                SphQuadD       squad;
                squad.init(trafo.scale, trafo.worldToFramePoint(p));
                const float sample_pdf = nsf * squad.pdf(trafo.scale);
                for (uint32_t k = 0; k < ns; ++k) {
                    float u0, u1;
                    sampler.sample2D(u0, u1);

                    const V3 ls  = squad.sample(u0, u1);
                    const V3 ws  = trafo.frameToWorldPoint(ls);
                    const V3 dir = normalize3(sub3(ws, p));

                    V3 wn = trafo.r2;
                    if (0 != light.two_sided && dot3(wn, dir) > 0.f) wn = neg3(wn);

                    if (-dot3(wn, dir) < kDotMin || 0.f == squad.S || (dot3(dir, n) <= 0.f && !translucent)) continue;

                    if (num_records < st.shadow_stride) {
                        const size_t rec       = size_t(slot) * st.shadow_stride + num_records;
                        const V3     origin    = offsetPoint(p, geo_n, dir);
                        const V3     light_pos = offsetRay(ws, wn);
                        st.sh_o[rec]  = make_float4(origin.x, origin.y, origin.z, sample_pdf * pick.pdf);
                        st.sh_p[rec]  = make_float4(light_pos.x, light_pos.y, light_pos.z, __uint_as_float(pick.offset));
                        st.sh_wi[rec] = make_float4(dir.x, dir.y, dir.z, 0.f);
                        num_records += 1;
                    } else {
                        st.counters[3] = 1;
                    }
                }
            }
        }
        if (kWalks) {
            // The walks stay in this short loop until ZYGPU_WALK_REFILL lanes have finished theirs: every trip through the code around
            // it (fetching vertices and picks, the other shapes, the write-back) is paid in instruction fetches by the whole warp.
            // A leaf iteration (Node.randomLight over up to four triangles + the triangle sample) costs several inner ones, so the
            // lanes that reached a leaf wait for the others to arrive and the warp takes the leaves together.
            uint32_t       walkers = __ballot_sync(kFull, walking);
            const uint32_t n0      = uint32_t(__popc(walkers));
            const uint32_t leave   = n0 > ZYGPU_WALK_REFILL ? n0 - ZYGPU_WALK_REFILL : 0u;
            while (0 != walkers) {
                const uint32_t leaves  = __ballot_sync(kFull, walking && walk_leaf);
                const bool     do_leaf = uint32_t(__popc(leaves)) >= ZYGPU_WALK_LEAF_RATIO * uint32_t(__popc(walkers & ~leaves));
                if (walking && walk_leaf == do_leaf) {
                    const ZygpuLight light = sc.lights[walk_pick.offset];
                    const TreeD      tr    = primitiveTree(sc.mesh_samplers[light.sampler]);
                    walking = walk.step(sc, tr, walk_op, walk_on, translucent, threshold, [&](LightPickD sp) {
                        FragD frag;  // the shading point and what offsets from it
                        frag.p      = p;
                        frag.geo_n  = geo_n;
                        num_records = meshLightTriangleSample(sc, st, slot, light, walk_pick, loadTrafo(sc.trafos, light.prop), frag, n, walk_op,
                                                              walk_on, translucent, sp, sampler, num_records);
                    });
                    if (walking) walk_leaf = lightNodeIsLeaf(tr, walk.t.node);
                }
                walkers = __ballot_sync(kFull, walking);
                if (uint32_t(__popc(walkers)) <= leave) break;
            }
        }
        if (active && !(kWalks && walking) && pick_i >= pick_count) {
            if (1 == Phase) st.pick_n[slot] = pick_count | (pick_count << 16);  // no finite pick: phase 2 only queues the records
            st.sh_n[slot] = num_records;
            storeSampler(st, slot, sampler, pool_word);
            if (1 != Phase && nullptr != st.queue_r && 0 != num_records) {
                const uint32_t base = atomicAdd(&st.counters[10], num_records);
                for (uint32_t k = 0; k < num_records; ++k) st.queue_r[base + k] = slot * st.shadow_stride + k;
            }
            active = false;
        }
    }
}

// The rest of PathtracerMIS.li: evaluateLight after the visibility test (pathtracer_mis.zig:252-277), the direct-light
// add (:116-117), mat_sample.sample and the next vertex (:121-166).
template <bool Split, bool Textured>
__global__ void __launch_bounds__(kBlock, Split ? 3 : ZYGPU_SHADE_BLOCKS) shadeBKernel(SceneDevice sc, ZygpuView view, PathState st, PassParams pass, uint32_t round) {
    __shared__ __align__(16) uint32_t sobol_tables[kSobolTableWords];
    const uint32_t count = st.counters[1];
    if (blockIdx.x * blockDim.x >= count) return;  // see shade_a
    loadSobolTables(sobol_tables);
    const LutsD    luts{sc.luts};
    const uint32_t iters = (count + gridDim.x * blockDim.x - 1) / (gridDim.x * blockDim.x);
    for (uint32_t it = 0; it < iters; ++it) {
        const uint32_t i     = it * gridDim.x * blockDim.x + blockIdx.x * blockDim.x + threadIdx.x;
        bool           alive = false;  // the slot got its first vertex of the next generation
        uint32_t       slot  = 0;
        uint32_t       child_id[2] = {0, 0};
        bool           child_on[2] = {false, false};
        if (i < count) {
            slot                 = st.queue_b[i];
            const uint4    smp   = st.smp[slot];
            uint32_t       pool  = smp.w;
            const uint32_t lane  = Split ? poolCurLane(pool, round) : 0;
            const uint32_t vid   = lane * st.capacity + slot;
            LoadedVertex   lv     = loadVertex(st, vid);
            const VertexD& vertex = lv.v;

            const uint32_t total_depth = vertex.probe_depth;

            SamplerD sampler;
            sampler.sobol.tables = sobol_tables;
            loadSampler(st, slot, smp, pass, view.spp_total, total_depth, sampler);
            if (ZYG_SAMPLER_SOBOL != view.sampler) sampler.use_sobol = false;

            FragD frag;
            shapeFragment(sc, lv.prop, vertex.ray, lv.isec, frag);

            MediaD media{0, {0, 0, 0}, {0, 0, 0}};
            if (Split && 0 != lv.num_media) media = unpackMedia(st.med[vid], lv.num_media);

            const bool          caustics   = 0 == (vertex.state & kPrimaryRay) ? 0 != view.caustics_path : true;
            const V3      wo = neg3(vertex.ray.d);
            ZygpuMaterial m  = sc.materials[__ldg(sc.material_ids + sc.props[frag.prop].parts_start + frag.part)];
            float               ior_outside      = 1.f;
            int                 highest_priority = -128;
            if (Split) mediaForSample(sc, media, frag, wo, ior_outside, highest_priority);
            // the same texels shade_a's material sample read
            const MatSampleD mat_sample = texturedMaterialSample<Split, Textured>(sc, m, frag, wo, Textured ? st.stoch[vid] : 0.f, view.regularize_roughness,
                                                                                  lv.reg_alpha, caustics, view.specular_threshold, ior_outside,
                                                                                  highest_priority);

            const uint32_t max_splits = Split ? maxSplits(lv.path_count_log2, 0 != (vertex.state & kPrimaryRay), total_depth) : 1;

            // evaluateLight for the visible records
            V3             next_light = splat3(0.f);
            const uint32_t num        = st.sh_n[slot];
            for (uint32_t k = 0; k < num; ++k) {
                const size_t rec = size_t(slot) * st.shadow_stride + k;
                const float4 wi4 = st.sh_wi[rec];
                if (0.f == wi4.w) continue;
                const float4 o4 = st.sh_o[rec];
                const float4 p4 = st.sh_p[rec];
                const V3     wi = {wi4.x, wi4.y, wi4.z};

                // Light.evaluateTo, light.zig:119-132
                const float         stochastic_r = sampler.sample1D();
                const uint32_t      light_id = __float_as_uint(p4.w) & 0x7FFFFFFFu;
                const ZygpuLight    light    = sc.lights[light_id];
                const TrafoD        ltrafo   = loadTrafo(sc.trafos, light.prop);
                const ZygpuMaterial lm       = sc.materials[__ldg(sc.material_ids + sc.props[light.prop].parts_start + light.part)];
                const float area     = 0.f != lm.emission_normalize ? shapeArea(sc.props[light.prop].shape, ltrafo.scale) : 1.f;
                // an image-mapped (PROP_IMAGE) light left the uvw of its sample in the record's position lanes (infinite lights:
                // the lanes are free) or in sh_uv (finite lights)
                const float2 luv = 0 != (__float_as_uint(p4.w) & 0x80000000u) || nullptr == st.sh_uv ? make_float2(p4.x, p4.y) : st.sh_uv[rec];
                const V3    radiance = ZYGPU_NULL != lm.emission_map
                                           ? emittanceRadianceMapped(lm, wi, ltrafo, area, false,
                                                                     imageTexel(sc.image_samplers[lm.emission_map], luv.x, luv.y, stochastic_r))
                                           : emittanceRadiance(lm, wi, ltrafo, area, false);

                const BxdfResult bxdf_result = mat_sample.template evaluate<Split, Textured>(luts, wi, max_splits);

                const float light_pdf = o4.w;
                const float weight    = predividedPowerHeuristic(light_pdf, bxdf_result.pdf);

                next_light = add3(next_light, mul3(scale3(weight, radiance), bxdf_result.reflection));
            }

            const V3 split_throughput = scale3(lv.split_weight, lv.throughput);
            ivalueAdd(st, slot, mul3(split_throughput, next_light), total_depth, 1, false, false);

            BxdfSample     sample_results[Split ? 2 : 1];
            const uint32_t path_count = mat_sample.template sample<Split, Textured>(luts, sampler, max_splits, sample_results);

            if (Split) pool = poolFree(pool, lane);  // the parent lives in registers from here on
            if (slot == pass.debug_slot) {
                for (uint32_t c = 0; c < path_count; ++c) {
                    printf("[gpu]   sample %u/%u event %u sw %g pdf %g wi %.9g %.9g %.9g\n", c, path_count, sample_results[c].event,
                           sample_results[c].split_weight, sample_results[c].pdf, sample_results[c].wi.x, sample_results[c].wi.y, sample_results[c].wi.z);
                }
            }

            // no sample: the vertex ends here (Pool.consume finds it terminated, vertex.zig:250-268)
            if (0 == path_count && 0 != view.alpha_transparency) alphaAdd(st, slot, 0 != (vertex.state & kTransparent), lv.throughput, lv.split_weight);

            for (uint32_t c = 0; c < (Split ? path_count : min(path_count, 1u)); ++c) {
                const BxdfSample& sample_result = sample_results[c];

                // Vertex.State.update, vertex.zig:30-43
                uint32_t state = vertex.state;
                if (kScatterSpecular == sample_result.scattering) {
                    state |= kSpecular;
                    state = 0.f == sample_result.reg_alpha ? (state | kSingular) : (state & ~kSingular);
                    if (0 != (state & kPrimaryRay)) state |= kStartedSpecular;
                } else if (kEventStraight != sample_result.event) {
                    state &= ~(kSpecular | kSingular | kPrimaryRay);
                }

                uint32_t vertex_depth = lv.vertex_depth;
                float    bxdf_pdf     = vertex.bxdf_pdf;
                V3       origin       = vertex.origin;
                V3       geo_n        = vertex.geo_n;
                float    reg_alpha    = lv.reg_alpha;
                if (kEventStraight != sample_result.event) {
                    state        = mat_sample.translucent ? (state | kTranslucent) : (state & ~kTranslucent);
                    vertex_depth = vertex.probe_depth;
                    bxdf_pdf     = sample_result.pdf;
                    origin       = frag.p;
                    geo_n        = mat_sample.geo_n;
                    reg_alpha    = sample_result.reg_alpha;
                }

                const V3 throughput = mul3(lv.throughput, divs3(sample_result.reflection, sample_result.pdf));

                const V3 next_o = frag.offsetP(sample_result.wi);  // Fragment.offsetRay, intersection.zig:118-120

                if (!(kEventTransmission == sample_result.event || kEventStraight == sample_result.event)) state &= ~kTransparent;

                uint32_t cvid = vid;
                uint32_t pcl2 = 0, num_media = 0;
                if (Split) {
                    const uint32_t clane = poolAlloc(pool);
                    if (clane > 3u) break;  // cannot happen: path_count keeps the live vertices of a sample at 4 or fewer
                    alive = alive || 0 == poolNextCount(pool);
                    pool  = poolAppendNext(pool, clane);
                    cvid  = clane * st.capacity + slot;

                    pcl2 = lv.path_count_log2 + (path_count > 1 ? 1u : 0u);  // path_count *= number of samples (1 or 2)

                    MediaD next_media = media;
                    if (kEventTransmission == sample_result.event) {  // Vertex.interfaceChange, vertex.zig:95-110
                        if (frag.sameHemisphere(sample_result.wi)) {
                            mediaRemove(next_media, frag.prop, frag.part);
                        } else {
                            mediaPush(next_media, frag.prop, frag.part);
                        }
                    }
                    num_media    = next_media.count;
                    st.med[cvid] = packMedia(next_media);
                    child_id[c]  = cvid;
                    child_on[c]  = true;
                } else {
                    alive = true;
                }

                st.ray_o[cvid]  = make_float4(next_o.x, next_o.y, next_o.z,
                                              __uint_as_float(packFlags(state, vertex.probe_depth + 1, vertex_depth, pcl2, num_media)));
                st.ray_d[cvid]  = make_float4(sample_result.wi.x, sample_result.wi.y, sample_result.wi.z, kRayMaxT);
                st.thr[cvid]    = make_float4(throughput.x, throughput.y, throughput.z, bxdf_pdf);
                st.prev_p[cvid] = make_float4(origin.x, origin.y, origin.z, reg_alpha);
                st.prev_n[cvid] = make_float4(geo_n.x, geo_n.y, geo_n.z, lv.split_weight * sample_result.split_weight);
            }
            sampler.incrementPadding();
            storeSampler(st, slot, sampler, pool);
        }
        queuePush(st.queue_a, &st.counters[4], alive, slot);
        if (Split) {
            queuePush(st.queue_t, &st.counters[7], child_on[0], child_id[0]);
            queuePush(st.queue_t, &st.counters[7], child_on[1], child_id[1]);
        }
    }
}

// Counter bookkeeping between the stages (single thread; the queue lengths never leave the device).
__global__ void beginGenerationKernel(PathState st) {  // after extend: the trace queue is consumed, nothing is queued yet
    st.counters[1]  = 0;
    st.counters[4]  = 0;
    st.counters[7]  = 0;
    st.counters[9]  = 0;
    st.counters[10] = 0;
    st.counters[11] = 0;
}
__global__ void beginRoundKernel(PathState st) {
    st.counters[1]  = 0;
    st.counters[10] = 0;
    st.counters[11] = 0;
}
__global__ void advanceKernel(PathState st) {
    st.counters[0] = st.counters[4];
    st.counters[1] = 0;
    st.counters[4]  = 0;
    st.counters[9]  = 0;
    st.counters[10] = 0;
    st.counters[11] = 0;
}

// Sensor.addSample for every sample of the pass, gathered per film pixel (sensor.zig:168-385, buffer_opaque.zig:39-45).
__device__ __forceinline__ V3 clampColor(V3 color, float mx) {  // sensor.zig:615-624
    const float mc = hmax3(color);
    if (mc > mx) return scale3(__fdiv_rn(mx, mc), color);
    return color;
}

__device__ __forceinline__ float filterEval(const ZygpuView& view, float s) {  // sensor.zig:626-628
    const float    cx     = zmin(fabsf(s), view.filter_range_end);
    const float    o      = cx * view.filter_inverse_interval;
    const uint32_t offset = uint32_t(o);
    const float    t      = o - float(offset);
    return zlerp(view.filter[offset], view.filter[min(offset + 1, 29u)], t);
}

__global__ void __launch_bounds__(kBlock) filmKernel(ZygpuView view, PathState st, PassParams pass, float4* film, float* film_alpha) {
    const int32_t  w  = view.resolution[0];
    const int32_t  h  = view.resolution[1];
    const int32_t  fr = view.filter_radius_int;
    const uint32_t padded = pass.padded_w * pass.padded_h;

    for (uint32_t pixel = blockIdx.x * blockDim.x + threadIdx.x; pixel < uint32_t(w * h); pixel += gridDim.x * blockDim.x) {
        const int32_t x = int32_t(pixel % uint32_t(w));
        const int32_t y = int32_t(pixel / uint32_t(w));
        // Sensor.add bounds test: pixels outside the crop receive nothing (sensor.zig:559-562)
        if (x < view.crop[0] || x >= view.crop[2] || y < view.crop[1] || y >= view.crop[3]) continue;

        float4 value = film[pixel];
        float  alpha = nullptr != film_alpha ? film_alpha[pixel] : 0.f;

        for (uint32_t s = 0; s < pass.samples_in_pass; ++s) {
            for (int32_t dy = -fr; dy <= fr; ++dy) {
                for (int32_t dx = -fr; dx <= fr; ++dx) {
                    // the sample's pixel q = (x + dx, y + dy) must have been rendered: crop extended by the filter radius
                    const int32_t qx = x + dx, qy = y + dy;
                    if (qx < view.crop[0] - fr || qx >= view.crop[2] + fr || qy < view.crop[1] - fr || qy >= view.crop[3] + fr) continue;
                    const uint32_t slot = s * padded + uint32_t(qy + fr) * pass.padded_w + uint32_t(qx + fr);

                    const float4 e  = st.acc_e[slot];
                    const float4 d  = st.acc_d[slot];
                    const float4 in = st.acc_i[slot];

                    const V3 emission = clampColor({e.x, e.y, e.z}, view.clamp_emission);
                    const V3 direct   = clampColor({d.x, d.y, d.z}, view.clamp_direct);
                    const V3 indirect = clampColor({in.x, in.y, in.z}, view.clamp_indirect);
                    const V3 composed = add3(add3(emission, direct), indirect);

                    float weight = 1.f;
                    if (fr > 0) {
                        const float ox = e.w - 0.5f;
                        const float oy = d.w - 0.5f;
                        weight         = filterEval(view, ox + float(dx)) * filterEval(view, oy + float(dy));
                    }
                    // Opaque.addPixel
                    value.x += weight * composed.x;
                    value.y += weight * composed.y;
                    value.z += weight * composed.z;
                    value.w += weight;
                    alpha += weight * in.w;  // Transparent.addPixel: the fourth lane of the colour, buffer_transparent.zig:45-54
                }
            }
        }
        film[pixel] = value;
        if (nullptr != film_alpha) film_alpha[pixel] = alpha;
    }
}

// The AOV half of Sensor.addSample (sensor.zig:197-219, 244-275, 328-377) on the gather layout of filmKernel: colour-like classes are
// filtered sums (aov.Buffer.addPixel), Depth keeps the smallest value of the pixel's own samples (lessPixel), MaterialId the value of the
// own sample with the largest centre weight, the first one on ties (overwritePixel: a thread sees its samples in sample order).
__global__ void __launch_bounds__(kBlock) aovFilmKernel(ZygpuView view, PathState st, PassParams pass, AovFilm aov) {
    const int32_t  w  = view.resolution[0];
    const int32_t  h  = view.resolution[1];
    const int32_t  fr = view.filter_radius_int;
    const uint32_t padded = pass.padded_w * pass.padded_h;

    for (uint32_t pixel = blockIdx.x * blockDim.x + threadIdx.x; pixel < uint32_t(w * h); pixel += gridDim.x * blockDim.x) {
        const int32_t x = int32_t(pixel % uint32_t(w));
        const int32_t y = int32_t(pixel / uint32_t(w));
        if (x < view.crop[0] || x >= view.crop[2] || y < view.crop[1] || y >= view.crop[3]) continue;

        float4 value[ZYG_AOV_NUM_CLASSES];
        for (uint32_t c = 0; c < ZYG_AOV_NUM_CLASSES; ++c) value[c] = nullptr != aov.layers[c] ? aov.layers[c][pixel] : make_float4(0.f, 0.f, 0.f, 0.f);

        for (uint32_t s = 0; s < pass.samples_in_pass; ++s) {
            for (int32_t dy = -fr; dy <= fr; ++dy) {
                for (int32_t dx = -fr; dx <= fr; ++dx) {
                    const int32_t qx = x + dx, qy = y + dy;
                    if (qx < view.crop[0] - fr || qx >= view.crop[2] + fr || qy < view.crop[1] - fr || qy >= view.crop[3] + fr) continue;
                    const uint32_t slot = s * padded + uint32_t(qy + fr) * pass.padded_w + uint32_t(qx + fr);

                    const float4 e  = st.acc_e[slot];
                    const float4 d  = st.acc_d[slot];
                    const float4 in = st.acc_i[slot];

                    float weight = 1.f;
                    if (fr > 0) weight = filterEval(view, (e.w - 0.5f) + float(dx)) * filterEval(view, (d.w - 0.5f) + float(dy));

                    auto add = [&](uint32_t c, V3 v) {
                        value[c].x += weight * v.x;
                        value[c].y += weight * v.y;
                        value[c].z += weight * v.z;
                        value[c].w += weight;
                    };
                    if (nullptr != aov.layers[ZYG_AOV_EMISSION]) add(ZYG_AOV_EMISSION, clampColor({e.x, e.y, e.z}, view.clamp_emission));
                    if (nullptr != aov.layers[ZYG_AOV_DIRECT]) add(ZYG_AOV_DIRECT, clampColor({d.x, d.y, d.z}, view.clamp_direct));
                    if (nullptr != aov.layers[ZYG_AOV_INDIRECT]) add(ZYG_AOV_INDIRECT, clampColor({in.x, in.y, in.z}, view.clamp_indirect));
                    if (nullptr != aov.layers[ZYG_AOV_ALBEDO]) {
                        const float4 a = st.aov_albedo[slot];
                        add(ZYG_AOV_ALBEDO, {a.x, a.y, a.z});
                    }
                    if (nullptr != aov.layers[ZYG_AOV_GEOMETRIC_NORMAL]) {
                        const float4 n = st.aov_gn[slot];
                        add(ZYG_AOV_GEOMETRIC_NORMAL, {n.x, n.y, n.z});
                    }
                    if (nullptr != aov.layers[ZYG_AOV_SHADING_NORMAL]) {
                        const float4 n = st.aov_sn[slot];
                        add(ZYG_AOV_SHADING_NORMAL, {n.x, n.y, n.z});
                    }
                    const float4 misc = st.aov_misc[slot];
                    // insert1 sets lane 0 only; lanes 1 and 2 keep the class default 0 (aov_value.zig:68-78)
                    if (nullptr != aov.layers[ZYG_AOV_ROUGHNESS]) add(ZYG_AOV_ROUGHNESS, {misc.x, 0.f, 0.f});
                    if (0 == dx && 0 == dy) {
                        if (nullptr != aov.layers[ZYG_AOV_DEPTH] && misc.y < value[ZYG_AOV_DEPTH].x) value[ZYG_AOV_DEPTH].x = misc.y;
                        if (nullptr != aov.layers[ZYG_AOV_MATERIAL_ID] && weight > value[ZYG_AOV_MATERIAL_ID].w) {
                            value[ZYG_AOV_MATERIAL_ID].x = misc.z;
                            value[ZYG_AOV_MATERIAL_ID].w = weight;
                        }
                    }
                }
            }
        }
        for (uint32_t c = 0; c < ZYG_AOV_NUM_CLASSES; ++c) {
            if (nullptr != aov.layers[c]) aov.layers[c][pixel] = value[c];
        }
    }
}

// aov.Buffer.clear, aov_buffer.zig:39-49
__global__ void __launch_bounds__(kBlock) aovClearKernel(AovFilm aov, uint32_t num_pixels) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < num_pixels; i += gridDim.x * blockDim.x) {
        for (uint32_t c = 0; c < ZYG_AOV_NUM_CLASSES; ++c) {
            if (nullptr == aov.layers[c]) continue;
            const float d    = ZYG_AOV_DEPTH == c ? FLT_MAX : 0.f;
            aov.layers[c][i] = make_float4(d, d, d, 0.f);
        }
    }
}

// aov.Buffer.resolve, aov_buffer.zig:51-82, by Class.encoding (aov_value.zig:32-40)
__global__ void __launch_bounds__(kBlock) resolveAovKernel(uint32_t aov_class, const float4* layer, float4* rgba, uint32_t num_pixels) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < num_pixels; i += gridDim.x * blockDim.x) {
        const float4 p = layer[i];
        float4       out;
        if (ZYG_AOV_ALBEDO == aov_class || aov_class >= ZYG_AOV_EMISSION) {  // Color: |rgb| / weight, AP1 -> sRGB
            const V3 c    = {__fdiv_rn(fabsf(p.x), p.w), __fdiv_rn(fabsf(p.y), p.w), __fdiv_rn(fabsf(p.z), p.w)};
            const V3 srgb = add3(add3(scale3(c.x, {1.70505155f, -0.13025714f, -0.02400328f}), scale3(c.y, {-0.62179068f, 1.14080289f, -0.12896877f})),
                                 scale3(c.z, {-0.08325840f, -0.01054853f, 1.15297171f}));
            out           = make_float4(srgb.x, srgb.y, srgb.z, 1.f);
        } else if (ZYG_AOV_GEOMETRIC_NORMAL == aov_class || ZYG_AOV_SHADING_NORMAL == aov_class) {  // Normal
            out = make_float4(__fdiv_rn(p.x, p.w), __fdiv_rn(p.y, p.w), __fdiv_rn(p.z, p.w), 1.f);
        } else if (ZYG_AOV_ROUGHNESS == aov_class) {  // Float
            out = make_float4(__fdiv_rn(p.x, p.w), 0.f, 0.f, 1.f);
        } else {  // Depth, Id
            out = make_float4(p.x, 0.f, 0.f, 1.f);
        }
        rgba[i] = out;
    }
}

// ---- denoise (src/it/denoise.zig) -------------------------------------------------------------------------------------------------------
// The inputs `it denoise` reads from the exported files, taken from the buffers they were exported from: colour = the resolved beauty in
// AP1 (Opaque.resolveTonemap without the matrix to sRGB primaries, which the tool's image loader undoes again), normal = the resolved
// ShadingNormal layer ("_n"), albedo = the resolved Albedo layer in AP1. The depth input only feeds `dd`, which the reference overwrites
// with 1 (denoise.zig:229-230): it is not read.
__device__ __forceinline__ V3 denoiseColor(const ZygpuView& view, const float4* film, int32_t x, int32_t y) {
    const float4 p = film[size_t(y) * size_t(view.resolution[0]) + size_t(x)];
    return scale3(view.exposure_factor, {fabsf(__fdiv_rn(p.x, p.w)), fabsf(__fdiv_rn(p.y, p.w)), fabsf(__fdiv_rn(p.z, p.w))});
}
__device__ __forceinline__ float denoiseLuma(V3 c) { return powf(hmax3(c), 1.f / 2.2f); }

__global__ void __launch_bounds__(kBlock) denoiseKernel(ZygpuView view, const float4* __restrict__ film, const float4* __restrict__ normal,
                                                        const float4* __restrict__ albedo, const float* __restrict__ weights, int32_t radius,
                                                        float4* __restrict__ rgba) {
    const int32_t w = view.resolution[0], h = view.resolution[1];
    for (uint32_t pixel = blockIdx.x * blockDim.x + threadIdx.x; pixel < uint32_t(w * h); pixel += gridDim.x * blockDim.x) {
        const int32_t px = int32_t(pixel % uint32_t(w)), py = int32_t(pixel / uint32_t(w));

        auto normalAt = [&](int32_t x, int32_t y) -> V3 {
            const float4 p = normal[size_t(y) * size_t(w) + size_t(x)];
            return {__fdiv_rn(p.x, p.w), __fdiv_rn(p.y, p.w), __fdiv_rn(p.z, p.w)};
        };
        auto albedoAt = [&](int32_t x, int32_t y) -> V3 {
            const float4 p = albedo[size_t(y) * size_t(w) + size_t(x)];
            return {__fdiv_rn(fabsf(p.x), p.w), __fdiv_rn(fabsf(p.y), p.w), __fdiv_rn(fabsf(p.z), p.w)};
        };

        const V3 ref_color  = denoiseColor(view, film, px, py);
        const V3 ref_n      = normalAt(px, py);
        const V3 ref_albedo = albedoAt(px, py);

        // estimateNoise, :375-451: coefficient of variation of the 3 x 3 neighbourhood's gamma-encoded brightness
        float sum = 0.f;
        float l[9];
        for (int32_t y = -1, k = 0; y <= 1; ++y) {
            for (int32_t x = -1; x <= 1; ++x, ++k) {
                l[k] = denoiseLuma(denoiseColor(view, film, min(max(px + x, 0), w - 1), min(max(py + y, 0), h - 1)));
                sum += l[k];
            }
        }
        const float norm    = __fdiv_rn(1.f, 9.f);
        const float mean    = sum * norm;
        float       dif_sum = 0.f;
        for (int k = 0; k < 9; ++k) {
            const float dif = l[k] - mean;
            dif_sum += dif * dif;
        }
        const float std_dev        = __fsqrt_rn(norm * dif_sum);
        const float coef           = mean > 0.f ? __fdiv_rn(std_dev, mean) : 0.f;
        const float noise_estimate = zmin(coef * 20.f * zmin(mean, 1.f), 1.f);

        // filter, :175-246
        V3       result = splat3(0.f);
        uint32_t tap    = 0;
        for (int32_t y = -radius; y <= radius; ++y) {
            for (int32_t x = -radius; x <= radius; ++x) {
                const int32_t sx = min(max(px + x, 0), w - 1), sy = min(max(py + y, 0), h - 1);
                const float   weight = weights[tap++];

                const V3    f_n         = normalAt(sx, sy);
                const V3    f_albedo    = albedoAt(sx, sy);
                const float dot_n       = saturate(dot3(ref_n, f_n));
                const V3    da          = sub3(ref_albedo, f_albedo);
                const float dist_albedo = zmin(__fsqrt_rn(dot3(da, da)), 1.f);
                const float strength    = 1.f * (dot_n * dot_n) * (1.f - dist_albedo) * noise_estimate;

                const V3 color = lerp3(ref_color, denoiseColor(view, film, sx, sy), splat3(strength));
                result         = add3(result, scale3(weight, color));
            }
        }
        const V3 srgb = add3(add3(scale3(result.x, {1.70505155f, -0.13025714f, -0.02400328f}), scale3(result.y, {-0.62179068f, 1.14080289f, -0.12896877f})),
                             scale3(result.z, {-0.08325840f, -0.01054853f, 1.15297171f}));
        rgba[pixel]   = make_float4(srgb.x, srgb.y, srgb.z, 1.f);
    }
}

// Opaque.resolveTonemap with the Linear tonemapper, buffer_opaque.zig:73-79, tonemapper.zig:36-39, aces.zig:19-27
// Transparent.resolveTonemap, buffer_transparent.zig:82-93: the same colour, alpha = |sum of weight * alpha / weight|
__global__ void __launch_bounds__(kBlock) resolveKernel(ZygpuView view, const float4* film, const float* film_alpha, float4* rgba, uint32_t num_pixels) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < num_pixels; i += gridDim.x * blockDim.x) {
        const float4 p = film[i];
        const V3     c = {fabsf(__fdiv_rn(p.x, p.w)), fabsf(__fdiv_rn(p.y, p.w)), fabsf(__fdiv_rn(p.z, p.w))};
        const V3     s = scale3(view.exposure_factor, c);
        const V3     srgb = add3(add3(scale3(s.x, {1.70505155f, -0.13025714f, -0.02400328f}), scale3(s.y, {-0.62179068f, 1.14080289f, -0.12896877f})),
                                 scale3(s.z, {-0.08325840f, -0.01054853f, 1.15297171f}));
        rgba[i] = make_float4(srgb.x, srgb.y, srgb.z, nullptr != film_alpha ? fabsf(__fdiv_rn(film_alpha[i], p.w)) : 1.f);
    }
}

}  // namespace

cudaError_t uploadSobolDirections() {
    // Joe & Kuo direction numbers (new-joe-kuo-6.21201) for dimensions 1-5: degree s, coefficients a, initial m_i.
    static const uint32_t S[5]    = {0, 1, 2, 3, 3};
    static const uint32_t A[5]    = {0, 0, 1, 1, 2};
    static const uint32_t M[5][3] = {{0, 0, 0}, {1, 0, 0}, {1, 3, 0}, {1, 3, 1}, {1, 1, 1}};
    uint32_t              d[5][32];
    for (uint32_t i = 0; i < 32; ++i) d[0][i] = 1u << (31 - i);
    for (uint32_t j = 1; j < 5; ++j) {
        const uint32_t s = S[j];
        for (uint32_t i = 0; i < 32; ++i) {
            if (i < s) {
                d[j][i] = M[j][i] << (31 - i);
            } else {
                uint32_t v = d[j][i - s] ^ (d[j][i - s] >> s);
                for (uint32_t k = 1; k < s; ++k) v ^= ((A[j] >> (s - 1 - k)) & 1u) * d[j][i - k];
                d[j][i] = v;
            }
        }
    }
    static uint32_t tables[kSobolTableWords];
    for (uint32_t byte = 0; byte < 4; ++byte) {
        for (uint32_t dim = 0; dim < 5; ++dim) {
            for (uint32_t value = 0; value < 256; ++value) {
                uint32_t x = 0;
                for (uint32_t j = 0; j < 8; ++j) {
                    if (0 != ((value >> j) & 1u)) x ^= d[dim][8 * byte + j];
                }
                tables[(byte * 5 + dim) * 256 + value] = x;
            }
        }
    }
    return cudaMemcpyToSymbol(d_sobol_tables, tables, sizeof(tables));
}

cudaError_t launchGenerate(const ZygpuView& view, const PathState& st, const PassParams& pass, cudaStream_t stream) {
    generateKernel<<<gridFor(pass.num_paths, 16), kBlock, 0, stream>>>(view, st, pass);
    return cudaGetLastError();
}
cudaError_t launchShadeA(const SceneDevice& scene, const ZygpuView& view, const PathState& st, const PassParams& pass, uint32_t max_items,
                         uint32_t round, cudaStream_t stream) {
    // triangle-mesh lights are always sampled by the light kernels (capi/zygpu_render.cu): no instance samples them inline
    const uint32_t features = (st.lanes > 1 ? kFeatureSplit : 0u) | (scene.num_mesh_samplers > 0 ? kFeatureMeshLights : 0u) |
                              (scene.num_infinite_props > 0 ? kFeatureInfiniteLights : 0u) | (nullptr != st.queue_l ? kFeatureDeferredLights : 0u) |
                              (nullptr != st.stoch ? kFeatureTextured : 0u);
    if (0 != (features & kFeatureMeshLights) && 0 == (features & kFeatureDeferredLights)) return cudaErrorInvalidValue;
    const uint32_t grid = gridFor(max_items, shadeGrid(st.lanes > 1));
    if (st.lanes <= 1) round = 0;
#define ZYGPU_SHADE_A(F) \
    case F: shadeAKernel<F><<<grid, kBlock, 0, stream>>>(scene, view, st, pass, round); break;
    switch (features) {
        ZYGPU_SHADE_A(0) ZYGPU_SHADE_A(1) ZYGPU_SHADE_A(4) ZYGPU_SHADE_A(5) ZYGPU_SHADE_A(8) ZYGPU_SHADE_A(9) ZYGPU_SHADE_A(10) ZYGPU_SHADE_A(11)
        ZYGPU_SHADE_A(12) ZYGPU_SHADE_A(13) ZYGPU_SHADE_A(14) ZYGPU_SHADE_A(15)
        ZYGPU_SHADE_A(16) ZYGPU_SHADE_A(17) ZYGPU_SHADE_A(20) ZYGPU_SHADE_A(21) ZYGPU_SHADE_A(24) ZYGPU_SHADE_A(25) ZYGPU_SHADE_A(26) ZYGPU_SHADE_A(27)
        ZYGPU_SHADE_A(28) ZYGPU_SHADE_A(29) ZYGPU_SHADE_A(30) ZYGPU_SHADE_A(31)
        default: return cudaErrorInvalidValue;
    }
#undef ZYGPU_SHADE_A
    return cudaGetLastError();
}
cudaError_t launchLightStages(const SceneDevice& scene, const ZygpuView& view, const PathState& st, const PassParams& pass, uint32_t max_items,
                              cudaStream_t stream) {
    if (nullptr == st.queue_l) return cudaSuccess;
    cudaError_t err = cudaMemsetAsync(st.counters + 12, 0, 3 * sizeof(uint32_t), stream);
    if (cudaSuccess != err) return err;

    static int resident_select = 0, resident_sample = 0;
    if (0 == resident_select) {
        int per_sm = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, lightSelectPersistent, 128, 0);
        resident_select = std::max(per_sm, 1) * numSms();
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, lightSamplePersistent<true, true, 0>, 128, 0);
        resident_sample = std::max(per_sm, 1) * numSms();
    }
    const uint32_t needed = (max_items + 127) / 128;
    lightSelectPersistent<<<std::max(1u, std::min<uint32_t>(uint32_t(resident_select), needed)), 128, 0, stream>>>(scene, view, st, st.counters + 12);

    const uint32_t grid = std::max(1u, std::min<uint32_t>(uint32_t(resident_sample), needed));
    const bool     ml = scene.num_mesh_samplers > 0, inf = scene.num_infinite_props > 0;
    static const bool phases = nullptr == getenv("ZYGPU_LIGHT_PHASES") || 0 != atoi(getenv("ZYGPU_LIGHT_PHASES"));
    if (ml && inf && phases) {
        lightSamplePersistent<true, true, 1><<<grid, 128, 0, stream>>>(scene, view, st, pass, st.counters + 13);
        lightSamplePersistent<true, true, 2><<<grid, 128, 0, stream>>>(scene, view, st, pass, st.counters + 14);
    } else if (ml && inf) {
        lightSamplePersistent<true, true, 0><<<grid, 128, 0, stream>>>(scene, view, st, pass, st.counters + 13);
    } else if (ml) {
        lightSamplePersistent<true, false, 0><<<grid, 128, 0, stream>>>(scene, view, st, pass, st.counters + 13);
    } else if (inf && phases) {
        lightSamplePersistent<false, true, 1><<<grid, 128, 0, stream>>>(scene, view, st, pass, st.counters + 13);
        lightSamplePersistent<false, true, 2><<<grid, 128, 0, stream>>>(scene, view, st, pass, st.counters + 14);
    } else if (inf) {
        lightSamplePersistent<false, true, 0><<<grid, 128, 0, stream>>>(scene, view, st, pass, st.counters + 13);
    } else {
        lightSamplePersistent<false, false, 0><<<grid, 128, 0, stream>>>(scene, view, st, pass, st.counters + 13);
    }
    return cudaGetLastError();
}
cudaError_t launchBeginGeneration(const PathState& st, cudaStream_t stream) {
    beginGenerationKernel<<<1, 1, 0, stream>>>(st);
    return cudaGetLastError();
}
cudaError_t launchBeginRound(const PathState& st, cudaStream_t stream) {
    beginRoundKernel<<<1, 1, 0, stream>>>(st);
    return cudaGetLastError();
}
cudaError_t launchEndGeneration(const PathState& st, cudaStream_t stream) {
    advanceKernel<<<1, 1, 0, stream>>>(st);
    return cudaGetLastError();
}
cudaError_t launchShadeB(const SceneDevice& scene, const ZygpuView& view, const PathState& st, const PassParams& pass, uint32_t max_items,
                         uint32_t round, cudaStream_t stream) {
    if (st.lanes > 1) {
        if (nullptr != st.stoch) {
            shadeBKernel<true, true><<<gridFor(max_items, shadeGrid(true)), kBlock, 0, stream>>>(scene, view, st, pass, round);
        } else {
            shadeBKernel<true, false><<<gridFor(max_items, shadeGrid(true)), kBlock, 0, stream>>>(scene, view, st, pass, round);
        }
    } else {
        if (nullptr != st.stoch) {
            shadeBKernel<false, true><<<gridFor(max_items, shadeGrid(false)), kBlock, 0, stream>>>(scene, view, st, pass, 0);
        } else {
            shadeBKernel<false, false><<<gridFor(max_items, shadeGrid(false)), kBlock, 0, stream>>>(scene, view, st, pass, 0);
        }
        advanceKernel<<<1, 1, 0, stream>>>(st);
    }
    return cudaGetLastError();
}
cudaError_t launchFilm(const ZygpuView& view, const PathState& st, const PassParams& pass, float4* film, float* film_alpha, cudaStream_t stream) {
    filmKernel<<<gridFor(uint32_t(view.resolution[0] * view.resolution[1]), 16), kBlock, 0, stream>>>(view, st, pass, film, film_alpha);
    return cudaGetLastError();
}
cudaError_t launchMeshLightProps(const MeshDevice& mesh, const MeshSamplerDevice& sampler, float4* props, cudaStream_t stream) {
    if (0 == sampler.num_triangles) return cudaSuccess;
    meshLightPropsKernel<<<(sampler.num_triangles + 127u) / 128u, 128, 0, stream>>>(mesh, sampler, props);
    return cudaGetLastError();
}

cudaError_t launchAovClear(const AovFilm& aov, uint32_t num_pixels, cudaStream_t stream) {
    aovClearKernel<<<gridFor(num_pixels, 16), kBlock, 0, stream>>>(aov, num_pixels);
    return cudaGetLastError();
}
cudaError_t launchAovFilm(const ZygpuView& view, const PathState& st, const PassParams& pass, const AovFilm& aov, cudaStream_t stream) {
    aovFilmKernel<<<gridFor(uint32_t(view.resolution[0] * view.resolution[1]), 16), kBlock, 0, stream>>>(view, st, pass, aov);
    return cudaGetLastError();
}
cudaError_t launchResolveAov(uint32_t aov_class, const float4* layer, float4* rgba, uint32_t num_pixels, cudaStream_t stream) {
    resolveAovKernel<<<gridFor(num_pixels, 16), kBlock, 0, stream>>>(aov_class, layer, rgba, num_pixels);
    return cudaGetLastError();
}
cudaError_t launchDenoise(const ZygpuView& view, const float4* film, const float4* normal, const float4* albedo, const float* weights, int32_t radius,
                          float4* rgba, cudaStream_t stream) {
    denoiseKernel<<<gridFor(uint32_t(view.resolution[0] * view.resolution[1]), 16), kBlock, 0, stream>>>(view, film, normal, albedo, weights, radius, rgba);
    return cudaGetLastError();
}
cudaError_t launchResolve(const ZygpuView& view, const float4* film, const float* film_alpha, float4* rgba, uint32_t num_pixels, cudaStream_t stream) {
    resolveKernel<<<gridFor(num_pixels, 16), kBlock, 0, stream>>>(view, film, film_alpha, rgba, num_pixels);
    return cudaGetLastError();
}

}  // namespace zygpu
