// Device restatement of the shading side of the reference for the render stages (render.cu):
// samplers, analytic shapes, Substitute / Light materials, light tree, light sampling.
// Every function names the reference code it follows; arithmetic order is the reference's
// (see zmath.cuh for the fp rules).
#pragma once

#include "render.cuh"
#include "zmath.cuh"

namespace zygpu {

// ---- samplers --------------------------------------------------------------------------------

// sobol5 (sobol.zig:176-192) XORs one direction number per set bit of the index. The device folds the 32 direction numbers of a
// dimension (sobol.zig:194-245, regenerated in render.cu) into four 256-entry tables, one per byte of the index:
// table[(byte * 5 + dim) * 256 + value] = XOR of the directions of the bits set in `value`. Same integers, 20 lookups per block.
constexpr uint32_t kSobolTableWords = 4 * 5 * 256;
__device__ __align__(16) uint32_t d_sobol_tables[kSobolTableWords];

// Copies the tables into the block's shared memory (16 bytes per load: the copy is a serial prologue of every block, 11 % of
// the stall samples of a 16-spp Cornell pass with word loads); every thread of the block must call it once before sampling.
// `shared_tables` must be 16-byte aligned.
__device__ __forceinline__ void loadSobolTables(uint32_t* shared_tables) {
    const uint4* src = reinterpret_cast<const uint4*>(d_sobol_tables);
    uint4*       dst = reinterpret_cast<uint4*>(shared_tables);
    for (uint32_t i = threadIdx.x; i < kSobolTableWords / 4; i += blockDim.x) dst[i] = src[i];
    __syncthreads();
}

__device__ __forceinline__ uint32_t sobolHash(uint32_t i) {  // sobol.zig:107-124
    uint32_t x = i ^ (i >> 16);
    x *= 0x7feb352du;
    x ^= x >> 15;
    x *= 0x846ca68bu;
    x ^= x >> 16;
    return x;
}
__device__ __forceinline__ uint32_t hashCombine(uint32_t seed, uint32_t v) { return seed ^ (v + (seed << 6) + (seed >> 2)); }
__device__ __forceinline__ uint32_t laineKarras(uint32_t i, uint32_t seed) {  // sobol.zig:142-174
    uint32_t x = i ^ (i * 0x3d20adeau);
    x += seed;
    x *= (seed >> 16) | 1u;
    x ^= x * 0x05526c56u;
    x ^= x * 0x53a22864u;
    return x;
}
__device__ __forceinline__ uint32_t nestedUniformScramble(uint32_t x, uint32_t seed) {  // :136-140
    return __brev(laineKarras(__brev(x), seed));
}

struct SobolD {  // sobol.zig:8-105
    float           buffer[5];
    uint32_t        sample, dimension, block_seed, run_seed;
    const uint32_t* tables;  // shared-memory copy of d_sobol_tables

    __device__ void fill(uint32_t s) {  // incrementSeed body, :36-56
        const float    S = 1.f / 4294967296.f;
        const uint32_t i = nestedUniformScramble(sample, s);
        // sobol5, :176-192, a byte of the index at a time
        const uint32_t* t0 = tables + (i & 0xffu);
        const uint32_t* t1 = tables + 5 * 256 + ((i >> 8) & 0xffu);
        const uint32_t* t2 = tables + 10 * 256 + ((i >> 16) & 0xffu);
        const uint32_t* t3 = tables + 15 * 256 + (i >> 24);
        const uint32_t  x0 = t0[0] ^ t1[0] ^ t2[0] ^ t3[0];
        const uint32_t  x1 = t0[256] ^ t1[256] ^ t2[256] ^ t3[256];
        const uint32_t  x2 = t0[512] ^ t1[512] ^ t2[512] ^ t3[512];
        const uint32_t  x3 = t0[768] ^ t1[768] ^ t2[768] ^ t3[768];
        const uint32_t  x4 = t0[1024] ^ t1[1024] ^ t2[1024] ^ t3[1024];
        buffer[0] = __uint2float_rn(nestedUniformScramble(x0, hashCombine(s, 0))) * S;
        buffer[1] = __uint2float_rn(nestedUniformScramble(x1, hashCombine(s, 1))) * S;
        buffer[2] = __uint2float_rn(nestedUniformScramble(x2, hashCombine(s, 2))) * S;
        buffer[3] = __uint2float_rn(nestedUniformScramble(x3, hashCombine(s, 3))) * S;
        buffer[4] = __uint2float_rn(nestedUniformScramble(x4, hashCombine(s, 4))) * S;
    }
    __device__ void incrementSeed() {
        block_seed = run_seed;
        fill(block_seed);
        run_seed  = sobolHash(block_seed + 1);
        dimension = 0;
    }
    __device__ void startPixel(uint32_t s, uint32_t seed) {
        sample     = s;
        dimension  = 5;
        run_seed   = sobolHash(seed);
        block_seed = run_seed;
    }
    // state carried between stages: (block_seed, run_seed, dimension); the buffer is rebuilt on demand
    __device__ void restore(uint32_t s, uint32_t bseed, uint32_t rseed, uint32_t dim) {
        sample     = s;
        block_seed = bseed;
        run_seed   = rseed;
        dimension  = dim;
        if (dim < 5) fill(bseed);
    }
};

struct PcgD {  // src/base/random/generator.zig
    uint64_t state, inc;

    __device__ uint32_t randomUint() {
        const uint64_t old = state;
        state              = old * 6364136223846793005ull + inc;
        const uint32_t xrs = uint32_t(((old >> 18) ^ old) >> 27);
        const uint32_t rot = uint32_t(old >> 59);
        return (xrs >> rot) | (xrs << ((0u - rot) & 31));
    }
    __device__ void start(uint64_t s, uint64_t sequence) {
        state = 0;
        inc   = (sequence << 1) | 1;
        randomUint();
        state += s;
        randomUint();
    }
    __device__ float randomFloat() { return __uint_as_float((randomUint() & 0x007FFFFFu) | 0x3F800000u) - 1.f; }
};

struct SamplerD {  // sampler.zig:17-74 (+ Worker.pickSampler, worker.zig:201-207)
    bool   use_sobol;
    SobolD sobol;
    PcgD   rng;

    __device__ float sample1D() {
        if (!use_sobol) return rng.randomFloat();
        if (sobol.dimension >= 5) sobol.incrementSeed();
        return sobol.buffer[sobol.dimension++];
    }
    __device__ void sample2D(float& a, float& b) {
        if (!use_sobol) {
            a = rng.randomFloat();
            b = rng.randomFloat();
            return;
        }
        if (sobol.dimension >= 4) sobol.incrementSeed();
        const uint32_t d = sobol.dimension;
        sobol.dimension  = d + 2;
        a                = sobol.buffer[d];
        b                = sobol.buffer[d + 1];
    }
    __device__ V3 sample3D() {
        if (!use_sobol) {
            V3 r;
            r.x = rng.randomFloat();
            r.y = rng.randomFloat();
            r.z = rng.randomFloat();
            return r;
        }
        if (sobol.dimension >= 3) sobol.incrementSeed();
        const uint32_t d = sobol.dimension;
        sobol.dimension  = d + 3;
        return {sobol.buffer[d], sobol.buffer[d + 1], sobol.buffer[d + 2]};
    }
    __device__ void incrementPadding() {
        if (use_sobol) sobol.dimension = 5;
    }
};

// ---- shapes ----------------------------------------------------------------------------------

struct HitD {  // shape/intersection.zig:46-61 (trafo is re-read from the prop)
    float    t, u, v;
    uint32_t primitive;
};

struct FragD {  // shape/intersection.zig:63-124
    V3       p, geo_n, t, b, n;
    float    u, v;  // uvw[0..1]; uvw[3] (ray offset) is 0 for every shape in scope
    uint32_t prop, part, primitive;
    TrafoD   trafo;

    __device__ bool sameHemisphere(V3 w) const { return dot3(geo_n, w) > 0.f; }
    __device__ V3   offsetP(V3 w) const {  // :112-116, offset() == 0: @mulAdd(0, n, p) == p
        const V3 nn = sameHemisphere(w) ? geo_n : neg3(geo_n);
        return offsetRay(fmas3(0.f, nn, p), nn);
    }
};

// Rectangle.intersect, rectangle.zig:30-62
__device__ __forceinline__ bool rectangleIntersect(const RayT& ray, const TrafoD& trafo, HitD& isec) {
    const V3    n     = trafo.r2;
    const float d     = dot3(n, trafo.position);
    const float hit_t = __fdiv_rn(-(dot3(n, ray.o) - d), dot3(n, ray.d));

    if (hit_t >= ray.tmin && ray.tmax >= hit_t) {
        const V3 p = rayPoint(ray, hit_t);
        const V3 k = sub3(p, trafo.position);
        const V3 t = neg3(trafo.r0);

        const float u = __fdiv_rn(dot3(t, k), 0.5f * trafo.scale.x);
        if (u > 1.f || u < -1.f) return false;

        const V3    b = neg3(trafo.r1);
        const float v = __fdiv_rn(dot3(b, k), 0.5f * trafo.scale.y);
        if (v > 1.f || v < -1.f) return false;

        isec.u         = u;
        isec.v         = v;
        isec.t         = hit_t;
        isec.primitive = 0;
        return true;
    }
    return false;
}

// Rectangle.fragment, rectangle.zig:102-124
__device__ __forceinline__ void rectangleFragment(const RayT& ray, const HitD& isec, FragD& frag) {
    const V3 p = rayPoint(ray, isec.t);
    const V3 n = frag.trafo.r2;
    const V3 t = neg3(frag.trafo.r0);
    const V3 b = neg3(frag.trafo.r1);

    frag.p     = p;
    frag.t     = t;
    frag.b     = b;
    frag.n     = n;
    frag.geo_n = n;
    if (frag.trafo.scale.z < 0.f) {
        const V3    k = sub3(p, frag.trafo.position);
        const float u = dot3(t, k) * 2.f;
        const float v = dot3(b, k) * 2.f;
        frag.u        = 0.5f * (u + 1.f);
        frag.v        = 0.5f * (v + 1.f);
    } else {
        frag.u = 0.5f * (isec.u + 1.f);
        frag.v = 0.5f * (isec.v + 1.f);
    }
    frag.part = 0;
}

// Disk.intersect / intersectP, disk.zig:28-58, 115-134
__device__ __forceinline__ bool diskIntersect(const RayT& ray, const TrafoD& trafo, HitD& isec) {
    const V3    normal = trafo.r2;
    const float d      = dot3(normal, trafo.position);
    const float denom  = -dot3(normal, ray.d);
    const float numer  = dot3(normal, ray.o) - d;
    const float hit_t  = __fdiv_rn(numer, denom);

    if (hit_t >= ray.tmin && ray.tmax >= hit_t) {
        const V3    p      = rayPoint(ray, hit_t);
        const V3    k      = sub3(p, trafo.position);
        const float l      = dot3(k, k);
        const float radius = 0.5f * trafo.scale.x;
        if (l <= radius * radius) {
            const V3 sk    = divs3(k, radius);
            isec.u         = -dot3(trafo.r0, sk);
            isec.v         = -dot3(trafo.r1, sk);
            isec.t         = hit_t;
            isec.primitive = 0;
            return true;
        }
    }
    return false;
}

// Disk.fragment, disk.zig:98-113
__device__ __forceinline__ void diskFragment(const RayT& ray, const HitD& isec, FragD& frag) {
    frag.p     = rayPoint(ray, isec.t);
    frag.t     = neg3(frag.trafo.r0);
    frag.b     = neg3(frag.trafo.r1);
    frag.n     = frag.trafo.r2;
    frag.geo_n = frag.trafo.r2;
    frag.u     = 0.5f * (isec.u + 1.f);
    frag.v     = 0.5f * (isec.v + 1.f);
    frag.part  = 0;
}

// Disk as a light: DiskSamplerData + EquiAngularSampling + Disk.sampleTo / pdf, disk.zig:181-332, 492-533 (UseEquiAngularSampling)
struct EquiAngularD {
    float offset, min_t, max_t, scale, scale_sqr, angle_min, angle_extent;

    __device__ __forceinline__ void init(V3 source, V3 origin, V3 direction, float mi, float ma) {
        offset                = __fdiv_rn(dot3(direction, sub3(source, origin)), squaredLength3(direction));
        const V3 foot         = sub3(add3(origin, scale3(offset, direction)), source);
        scale_sqr             = squaredLength3(foot);
        scale                 = __fsqrt_rn(scale_sqr);
        const float inv_scale = 0.f == scale ? 0.f : __fdiv_rn(1.f, scale);
        angle_min             = atanf((mi - offset) * inv_scale);
        const float angle_max = atanf((ma - offset) * inv_scale);
        angle_extent          = angle_max - angle_min;
        min_t                 = mi;
        max_t                 = ma;
    }
    __device__ __forceinline__ float sample(float u, float& t) const {
        const float lt = scale * tanf(angle_min + u * angle_extent);
        const float p  = __fdiv_rn(scale, angle_extent * (scale_sqr + lt * lt));
        t              = zclamp(lt + offset, min_t, max_t);
        return p;
    }
    __device__ __forceinline__ float pdf(float t) const {
        if (min_t <= t && t < max_t) {
            const float lt = t - offset;
            return __fdiv_rn(scale, angle_extent * (scale_sqr + lt * lt));
        }
        return 0.f;
    }
    __device__ __forceinline__ float pdfAndSample(float t, float& u) const {
        const float lt = t - offset;
        u              = saturate(__fdiv_rn(atanf(__fdiv_rn(lt, scale)) - angle_min, angle_extent));
        return __fdiv_rn(scale, angle_extent * (scale_sqr + lt * lt));
    }
};

struct DiskLightD {
    V3           lp, xd, yd;
    float        radius;
    EquiAngularD eas0;
    bool         valid;

    __device__ __forceinline__ void init(const TrafoD& trafo, V3 p) { initLocal(trafo.worldToFramePoint(p), 0.5f * trafo.scale.x); }
    __device__ __forceinline__ void initLocal(V3 local_p, float r) {
        radius      = r;
        lp          = local_p;
        const V3 td = {lp.y, -lp.x, 0.f};
        xd          = (0.f == td.x && 0.f == td.y) ? V3{1.f, 0.f, 0.f} : normalize3(td);
        yd          = {-xd.y, xd.x, 0.f};
        eas0.init(lp, splat3(0.f), yd, -radius, radius);
        valid = 0.f != eas0.angle_extent;
    }
    // one sample of Disk.sampleTo's loop; false = the reference's `continue`
    template <typename Sampler>
    __device__ __forceinline__ bool sample(const TrafoD& trafo, V3 p, V3 n, bool two_sided, bool total_sphere, float nsf, Sampler& sampler, V3& ws,
                                           V3& wn, V3& dir, float& pdf) const {
        float u0, u1;
        sampler.sample2D(u0, u1);
        float x, y;
        diskConcentric(u0, u1, x, y);
        float u    = x;
        float pdf_ = __fdiv_rn(__fsqrt_rn(1.f - u * u), 0.25f * kPi);
        u          = (u + 1.f) * 0.5f;

        float y_coord;
        pdf_ *= eas0.sample(u, y_coord);

        const float x_chord = __fsqrt_rn(radius * radius - y_coord * y_coord);
        if (0.f == x_chord) return false;

        EquiAngularD eas1;
        eas1.init(lp, scale3(y_coord, yd), xd, -x_chord, x_chord);
        if (0.f == eas1.angle_extent) return false;

        float x_coord;
        pdf_ *= eas1.sample(sampler.sample1D(), x_coord);

        const V3 l_direction = sub3(add3(scale3(x_coord, xd), scale3(y_coord, yd)), lp);
        const V3 axis        = trafo.objectToWorldNormal(l_direction);
        ws                   = add3(p, axis);
        wn                   = trafo.r2;
        if (two_sided && dot3(wn, sub3(ws, p)) > 0.f) wn = neg3(wn);

        const float sl = squaredLength3(axis);
        dir            = divs3(axis, __fsqrt_rn(sl));
        const float c  = -dot3(wn, dir);
        if (c < kDotMin || (dot3(dir, n) <= 0.f && !total_sphere)) return false;
        pdf = __fdiv_rn((nsf * pdf_) * sl, c);
        return true;
    }
};

// Disk.pdf, disk.zig:492-533, in the disk's frame: lp = where the ray started, l_point = where it met the disk, c = |n . dir|, sl = the
// squared distance between the two. Out of line and on plain values: it sits on the emission path of every shade kernel instance, and
// inlined its atanf chains raise the spills of the instances that never see a Disk light.
static __device__ __noinline__ float diskLightPdfLocal(V3 lp, V3 l_point, float radius, float c, float sl, float nsf) {
    DiskLightD dl;
    dl.initLocal(lp, radius);
    const float y_coord = dot3(l_point, dl.yd);
    float       u;
    const float eas_pdf = dl.eas0.pdfAndSample(y_coord, u);
    u                   = u * 2.f - 1.f;
    float pdf_          = __fdiv_rn(__fsqrt_rn(1.f - u * u), 0.25f * kPi);
    pdf_ *= eas_pdf;
    const float  x_chord = __fsqrt_rn(dl.radius * dl.radius - y_coord * y_coord);
    EquiAngularD eas1;
    eas1.init(dl.lp, scale3(y_coord, dl.yd), dl.xd, -x_chord, x_chord);
    pdf_ *= eas1.pdf(dot3(l_point, dl.xd));
    return __fdiv_rn((nsf * pdf_) * sl, c);
}

// AABB.intersectP on the unit cube, aabb.zig:62-84
__device__ __forceinline__ float unitCubeIntersectP(const RayT& ray) {
    const float lx = (-0.5f - ray.o.x) * ray.inv_d.x, ly = (-0.5f - ray.o.y) * ray.inv_d.y, lz = (-0.5f - ray.o.z) * ray.inv_d.z;
    const float ux = (0.5f - ray.o.x) * ray.inv_d.x, uy = (0.5f - ray.o.y) * ray.inv_d.y, uz = (0.5f - ray.o.z) * ray.inv_d.z;

    const float t0x = zmin(lx, ux), t0y = zmin(ly, uy), t0z = zmin(lz, uz);
    const float t1x = zmax(lx, ux), t1y = zmax(ly, uy), t1z = zmax(lz, uz);

    const float imin = zmax(zmax(t0x, t0y), t0z);
    const float imax = zmin(zmin(t1x, t1y), t1z);

    const float tboxmin = zmax(imin, ray.tmin);
    const float tboxmax = zmin(imax, ray.tmax);

    if (tboxmin <= tboxmax) return imin < ray.tmin ? imax : imin;
    return FLT_MAX;
}

__device__ __forceinline__ RayT worldToObjectRay(const TrafoD& trafo, const RayT& ray) {  // composed_transformation.zig:119-126
    return makeRay(trafo.worldToObjectPoint(ray.o), trafo.worldToObjectVector(ray.d), ray.tmin, ray.tmax);
}

// Cube.intersect, cube.zig:24-38
__device__ __forceinline__ bool cubeIntersect(const RayT& ray, const TrafoD& trafo, HitD& isec) {
    const RayT  local_ray = worldToObjectRay(trafo, ray);
    const float hit_t     = unitCubeIntersectP(local_ray);
    if (hit_t < ray.tmax) {
        isec.t         = hit_t;
        isec.primitive = 0;
        return true;
    }
    return false;
}

// Cube.fragment, cube.zig:40-62
__device__ __forceinline__ void cubeFragment(const RayT& ray, const HitD& isec, FragD& frag) {
    const float hit_t = isec.t;
    frag.p            = rayPoint(ray, hit_t);

    const RayT local_ray = worldToObjectRay(frag.trafo, ray);
    const V3   local_p   = rayPoint(local_ray, hit_t);
    const V3   distance  = {fabsf(0.5f - fabsf(local_p.x)), fabsf(0.5f - fabsf(local_p.y)), fabsf(0.5f - fabsf(local_p.z))};

    // indexMinComponent3, vector4.zig:181-187
    uint32_t i;
    if (distance.x < distance.y) {
        i = distance.x < distance.z ? 0 : 2;
    } else {
        i = distance.y < distance.z ? 1 : 2;
    }
    const float lp = 0 == i ? local_p.x : (1 == i ? local_p.y : local_p.z);
    const V3    ri = 0 == i ? frag.trafo.r0 : (1 == i ? frag.trafo.r1 : frag.trafo.r2);
    const float s  = copysignf(1.f, lp);
    const V3    n  = scale3(s, ri);

    frag.part  = 0;
    frag.geo_n = n;
    frag.n     = n;
    frag.u     = 0.f;
    frag.v     = 0.f;
    orthonormalBasis3(n, frag.t, frag.b);
}

// Cube.intersectP, cube.zig:64-69 -> AABB.intersect, aabb.zig:46-60
__device__ __forceinline__ bool cubeIntersectP(const RayT& ray, const TrafoD& trafo) {
    const RayT r = worldToObjectRay(trafo, ray);
    return FLT_MAX != intersectNode(make_float4(-0.5f, -0.5f, -0.5f, 0.f), make_float4(0.5f, 0.5f, 0.5f, 0.f), r);
}

// Sphere.intersect, sphere.zig:28-62
__device__ __forceinline__ bool sphereIntersect(const RayT& ray, const TrafoD& trafo, HitD& isec) {
    const float idl = __fdiv_rn(1.f, length3(ray.d));
    const V3    nd  = scale3(idl, ray.d);

    const V3    v = sub3(trafo.position, ray.o);
    const float b = dot3(nd, v);

    const V3    remedy_term  = sub3(v, scale3(b, nd));
    const float radius       = 0.5f * trafo.scale.x;
    const float discriminant = radius * radius - dot3(remedy_term, remedy_term);

    if (discriminant > 0.f) {
        const float dist = __fsqrt_rn(discriminant);
        const float t0   = (b - dist) * idl;
        if (t0 >= ray.tmin && ray.tmax >= t0) {
            isec.t         = t0;
            isec.primitive = 0;
            return true;
        }
        const float t1 = (b + dist) * idl;
        if (t1 >= ray.tmin && ray.tmax >= t1) {
            isec.t         = t1;
            isec.primitive = 0;
            return true;
        }
    }
    return false;
}

// Sphere.fragment, sphere.zig:64-92
__device__ __forceinline__ void sphereFragment(const RayT& ray, const HitD& isec, FragD& frag) {
    const V3 p = rayPoint(ray, isec.t);
    const V3 n = normalize3(sub3(p, frag.trafo.position));

    frag.p     = p;
    frag.geo_n = n;
    frag.n     = n;
    frag.part  = 0;

    const V3    xyz   = normalize3(frag.trafo.worldToObjectNormal(n));
    const float phi   = -atan2f(xyz.x, xyz.z) + kPi;
    const float theta = acosf(xyz.y);

    float sin_phi, cos_phi;
    sincosf(phi, &sin_phi, &cos_phi);
    const float sin_theta = zmax(sinf(theta), 0.00001f);

    const V3 t = normalize3(frag.trafo.objectToWorldNormal({sin_theta * cos_phi, 0.f, sin_theta * sin_phi}));

    frag.t = t;
    frag.b = neg3(cross3(t, n));
    frag.u = phi * (0.5f * kPiInv);
    frag.v = theta * kPiInv;
}

// Sphere as a light: Sphere.sampleTo / Sphere.pdf, sphere.zig:323-393, 472-487; smpl.conePdfUniform, sampling.zig:103-106
__device__ __forceinline__ float conePdfUniform(float one_minus_cos_theta_max) {
    return __fdiv_rn(1.f, (2.f * kPi) * zmax(one_minus_cos_theta_max, 1.0e-20f));
}
struct SphereLightD {
    V3    z, tx, ty;  // Frame.init(z)
    float l, r;
    bool  valid;

    __device__ void init(const TrafoD& trafo, V3 p) {
        const V3 v = sub3(trafo.position, p);
        l          = length3(v);
        r          = 0.5f * trafo.scale.x;
        valid      = !(l <= (r + 0.0000001f));
        z          = scale3(__fdiv_rn(1.f, l), v);
        orthonormalBasis3(z, tx, ty);
    }
    // one light sample: false when it faces away from n; pdf excludes the sample count and the light pick
    __device__ bool sample(const TrafoD& trafo, V3 p, V3 n, bool total_sphere, float s0, float s1, V3& lp, V3& wn, V3& dir, float& pdf) const {
        const float sin_theta_max           = __fdiv_rn(r, l);
        const float sin2_theta_max          = sin_theta_max * sin_theta_max;
        const float cos_theta_max           = __fsqrt_rn(1.f - sin2_theta_max);
        float       one_minus_cos_theta_max = 1.f - cos_theta_max;

        float cos_theta  = (cos_theta_max - 1.f) * s0 + 1.f;
        float sin2_theta = 1.f - (cos_theta * cos_theta);
        if (sin2_theta_max < 0.00068523f) {
            sin2_theta              = sin2_theta_max * s0;
            cos_theta               = __fsqrt_rn(1.f - sin2_theta);
            one_minus_cos_theta_max = 0.5f * sin2_theta_max;
        }
        const float cos_alpha = zmin(__fdiv_rn(sin2_theta, sin_theta_max) +
                                         cos_theta * __fsqrt_rn(1.f - zmin(__fdiv_rn(sin2_theta, sin2_theta_max), 1.f)),
                                     1.f);
        const float sin_alpha = __fsqrt_rn(1.f - cos_alpha * cos_alpha);
        const float phi       = s1 * (2.f * kPi);
        const float sin_phi = sinf(phi), cos_phi = cosf(phi);
        const V3    w = {-(cos_phi * sin_alpha), -(sin_phi * sin_alpha), -cos_alpha};
        const FrameD frame{tx, ty, z};
        wn  = frame.frameToWorld(w);
        lp  = add3(trafo.position, scale3(r, wn));
        dir = normalize3(sub3(lp, p));
        if (dot3(dir, n) <= 0.f && !total_sphere) return false;
        pdf = conePdfUniform(one_minus_cos_theta_max);
        return true;
    }
};
__device__ __forceinline__ float sphereLightPdf(const TrafoD& trafo, V3 p) {
    const V3    v              = sub3(trafo.position, p);
    const float l2             = squaredLength3(v);
    const float r              = 0.5f * trafo.scale.x;
    const float sin2_theta_max = __fdiv_rn(r * r, l2);
    const float one_minus_cos_theta_max =
        sin2_theta_max < 0.00068523f ? 0.5f * sin2_theta_max : 1.f - __fsqrt_rn(zmax(1.f - sin2_theta_max, 0.f));
    return conePdfUniform(one_minus_cos_theta_max);
}

// ---- emission images --------------------------------------------------------------------------------------------------
// Distribution1D.sample, distribution_1d.zig:50-54, 250-258: the first i in [1, size - 1) with cdf[i] >= r (else size - 1),
// minus one. The reference walks there linearly from a lookup-table start; the cdf is non-decreasing on that range, so a
// binary search lands on the same entry.
__device__ __forceinline__ uint32_t dist1dSample(const float* __restrict__ cdf, uint32_t size, float r) {
    uint32_t lo = 1, hi = size - 1;
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (__ldg(cdf + mid) >= r) {
            hi = mid;
        } else {
            lo = mid + 1;
        }
    }
    return lo - 1;
}
__device__ __forceinline__ void dist1dSampleContinuous(const float* __restrict__ cdf, uint32_t size, float r, float& offset, float& pdf) {  // :62-75
    const uint32_t o = dist1dSample(cdf, size, r);
    const float    c = __ldg(cdf + o + 1);
    const float    v = c - __ldg(cdf + o);
    if (0.f == v) {
        offset = 0.f;
        pdf    = 0.f;
        return;
    }
    const float t = __fdiv_rn(c - r, v);
    offset        = __fdiv_rn(float(o) + t, float(size - 1));
    pdf           = v;
}
__device__ __forceinline__ float dist1dPdfF(const float* __restrict__ cdf, uint32_t size, float u) {  // :81-86
    const uint32_t o = min(uint32_t(u * float(size - 1)), size - 2);
    return __ldg(cdf + o + 1) - __ldg(cdf + o);
}
__device__ __forceinline__ float textureAddress(uint32_t mode, float x) {  // sampler_mode.zig:21-26
    return 0 == mode ? zmin(zmax(x, 0.f), 1.f) : x - floorf(x);
}
__device__ __forceinline__ int32_t textureCoord(uint32_t mode, int32_t c, int32_t end) {  // :35-40, 76-80
    if (0 == mode) return max(min(c, end - 1), 0);
    const int32_t m = c % end;
    return m < 0 ? m + end : m;
}
// ImageImpl.sample / pdf, shape_sampler.zig:128-152 over Distribution2D.sampleContinuous / pdf, distribution_2d.zig:67-87
__device__ __forceinline__ void imageSample(const ImageSamplerDevice& is, float r0, float r1, float& u, float& v, float& pdf) {
    float vp, up;
    dist1dSampleContinuous(is.marginal_cdf, is.height + 1, r1, v, vp);
    const uint32_t c = min(uint32_t(v * float(is.height)), is.height - 1);
    dist1dSampleContinuous(is.conditional_cdf + size_t(c) * (is.width + 1), is.width + 1, r0, u, up);
    pdf = (up * vp) * is.total_weight;
}
__device__ __forceinline__ float imagePdf(const ImageSamplerDevice& is, float u, float v) {
    const float    au = textureAddress(is.address_u, u), av = textureAddress(is.address_v, v);
    const float    v_pdf = dist1dPdfF(is.marginal_cdf, is.height + 1, av);
    const uint32_t c     = min(uint32_t(av * float(is.height)), is.height - 1);
    return (dist1dPdfF(is.conditional_cdf + size_t(c) * (is.width + 1), is.width + 1, au) * v_pdf) * is.total_weight;
}
// ts.sample2D_3 of an image texture, texture_sampler.zig:63-79; Nearest2D.map :99-124, LinearStochastic2D.map :126-170
__device__ __forceinline__ V3 imageTexel(const ImageSamplerDevice& is, float u, float v, float r) {
    const int32_t dx = int32_t(is.width), dy = int32_t(is.height);
    const float   s = is.scale_u * u, t = is.scale_v * v;
    int32_t       x, y;
    if (0 == is.filter) {
        x = min(int32_t(textureAddress(is.address_u, s) * float(dx)), dx - 1);
        y = min(int32_t(textureAddress(is.address_v, t) * float(dy)), dy - 1);
    } else {
        const float ms = textureAddress(is.address_u, s) * float(dx) - 0.5f, mt = textureAddress(is.address_v, t) * float(dy) - 0.5f;
        const float fs = floorf(ms), ft = floorf(mt);
        const float w0 = ms - fs, w1 = mt - ft;
        const float o0 = 1.f - w0, o1 = 1.f - w1;
        x              = int32_t(fs);
        y              = int32_t(ft);
        int32_t index     = 0;
        float   threshold = o0 * o1;
        index += r > threshold ? 1 : 0;
        threshold = __fmaf_rn(w0, o1, threshold);
        index += r > threshold ? 1 : 0;
        threshold = __fmaf_rn(o0, w1, threshold);
        index += r > threshold ? 1 : 0;
        x = textureCoord(is.address_u, x + (index & 1), dx);
        y = textureCoord(is.address_v, y + ((index & 2) >> 1), dy);
    }
    const float* px = is.pixels + 3 * (size_t(y) * is.width + size_t(x));
    return {__ldg(px), __ldg(px + 1), __ldg(px + 2)};
}

// Canopy, shape/canopy.zig:24-62, 164-202
__device__ __forceinline__ bool canopyIntersect(const RayT& ray, const TrafoD& trafo, HitD& isec) {
    if (ray.tmax < kRayMaxT || dot3(ray.d, trafo.r2) < -0.0005f) return false;
    isec.u = isec.v = 0.f;
    isec.primitive  = 0;
    isec.t          = kRayMaxT;
    return true;
}
__device__ __forceinline__ void canopyFragment(const RayT& ray, FragD& frag) {
    const V3    xyz        = normalize3(frag.trafo.transformVectorTransposed(ray.d));
    const float colatitude = acosf(xyz.z);
    const float longitude  = atan2f(-xyz.y, xyz.x);
    const float r          = colatitude * (kPiInv * 2.f);
    const float dx = r * cosf(longitude), dy = r * sinf(longitude);
    frag.u     = 0.5f * dx + 0.5f;
    frag.v     = 0.5f * dy + 0.5f;
    frag.p     = scale3(kRayMaxT, ray.d);
    frag.geo_n = neg3(ray.d);
    frag.t     = frag.trafo.r0;
    frag.b     = frag.trafo.r1;
    frag.n     = neg3(ray.d);
    frag.part  = 0;
}
__device__ __forceinline__ V3 canopyDiskToHemisphere(float ux, float uy) {
    const float longitude  = atan2f(-uy, ux);
    const float r          = __fsqrt_rn(ux * ux + uy * uy);
    const float colatitude = r * (kPi / 2.f);
    const float sin_col = sinf(colatitude), cos_col = cosf(colatitude);
    const float sin_lon = sinf(longitude), cos_lon = cosf(longitude);
    return {sin_col * cos_lon, sin_col * sin_lon, cos_col};
}
// Canopy.sampleMaterialTo, canopy.zig:94-131: false = no sample. `pdf` excludes the light pick.
__device__ __forceinline__ bool canopySampleMaterialTo(const ImageSamplerDevice& is, const TrafoD& trafo, V3 n, bool total_sphere, float r0,
                                                       float r1, V3& dir, float& u, float& v, float& pdf) {
    float ipdf;
    imageSample(is, r0, r1, u, v, ipdf);
    if (0.f == ipdf) return false;
    const float dx = 2.f * u - 1.f, dy = 2.f * v - 1.f;
    if (dx * dx + dy * dy > 1.f) return false;
    dir = trafo.transformVector(canopyDiskToHemisphere(dx, dy));
    if (dot3(dir, n) <= 0.f && !total_sphere) return false;
    pdf = __fdiv_rn(ipdf, 2.f * kPi);
    return true;
}

// Distant, shape/distant.zig:22-76, 139-145
__device__ __forceinline__ float distantSolidAngle(float radius) {
    return (2.f * kPi) * (1.f - __fsqrt_rn(__fdiv_rn(1.f, radius * radius + 1.f)));
}
__device__ __forceinline__ bool distantIntersect(const RayT& ray, const TrafoD& trafo, HitD& isec) {
    const float radius = trafo.scale.x;
    const V3    n      = trafo.r2;
    const float b      = dot3(n, ray.d);
    if (b > 0.f || ray.tmax < kRayMaxT || radius <= 0.f) return false;

    const float det = (b * b) - dot3(n, n) + (radius * radius);
    if (det >= 0.f) {
        const V3 sk    = divs3(sub3(ray.d, n), radius);
        isec.u         = dot3(trafo.r0, sk);
        isec.v         = dot3(trafo.r1, sk);
        isec.primitive = 0;
        isec.t         = kRayMaxT;
        return true;
    }
    return false;
}
__device__ __forceinline__ void distantFragment(const RayT& ray, const HitD& isec, FragD& frag) {
    frag.p     = scale3(kRayMaxT, ray.d);
    frag.geo_n = frag.trafo.r2;
    frag.t     = frag.trafo.r0;
    frag.b     = frag.trafo.r1;
    frag.n     = frag.trafo.r2;
    frag.u     = (isec.u + 1.f) * 0.5f;
    frag.v     = (isec.v + 1.f) * 0.5f;
    frag.part  = 0;
}

// Mesh.fragment, triangle_mesh.zig:310-335 + Data.interpolateData / normal, triangle_data.zig:106-149
__device__ __forceinline__ V3 decompressNormal(const uint16_t* normals, uint32_t i) {  // encoding.zig:91-108
    const uint32_t packed = __ldg(reinterpret_cast<const uint32_t*>(normals) + i);
    const float    o0     = __fmaf_rn(float(packed & 0xffffu), 1.f / 32768.f, -1.f);
    const float    o1     = __fmaf_rn(float(packed >> 16), 1.f / 32768.f, -1.f);
    V3             v      = {o0, o1, -1.f + fabsf(o0) + fabsf(o1)};
    const float    t      = zmax(v.z, 0.f);
    v.x += v.x > 0.f ? -t : t;
    v.y += v.y > 0.f ? -t : t;
    return normalize3(v);
}

__device__ __forceinline__ V3 interpolate3(V3 a, V3 b, V3 c, float u, float v) {  // triangle.zig:142-149
    const float w     = 1.f - u - v;
    const V3    temp0 = fma3(b, splat3(u), scale3(v, c));
    return fma3(a, splat3(w), temp0);
}

__device__ __forceinline__ V3 gramSchmidt(V3 v, V3 w) { return fmas3(-dot3(v, w), w, v); }  // vector4.zig:120-122

__device__ __forceinline__ void meshFragment(const MeshShading& m, const HitD& isec, FragD& frag) {
    const uint32_t prim = isec.primitive;
    frag.part           = __ldg(m.parts + prim);

    const uint32_t ia = __ldg(m.triangles + 3 * size_t(prim) + 0);
    const uint32_t ib = __ldg(m.triangles + 3 * size_t(prim) + 1);
    const uint32_t ic = __ldg(m.triangles + 3 * size_t(prim) + 2);

    const V3 pa = {__ldg(m.positions + 3 * size_t(ia)), __ldg(m.positions + 3 * size_t(ia) + 1), __ldg(m.positions + 3 * size_t(ia) + 2)};
    const V3 pb = {__ldg(m.positions + 3 * size_t(ib)), __ldg(m.positions + 3 * size_t(ib) + 1), __ldg(m.positions + 3 * size_t(ib) + 2)};
    const V3 pc = {__ldg(m.positions + 3 * size_t(ic)), __ldg(m.positions + 3 * size_t(ic) + 1), __ldg(m.positions + 3 * size_t(ic) + 2)};

    const V3 geo_n = normalize3(cross3(sub3(pb, pa), sub3(pc, pa)));
    frag.geo_n     = frag.trafo.objectToWorldNormal(geo_n);

    const V3 p = interpolate3(pa, pb, pc, isec.u, isec.v);

    const float2 uva = __ldg(reinterpret_cast<const float2*>(m.uvs) + ia);
    const float2 uvb = __ldg(reinterpret_cast<const float2*>(m.uvs) + ib);
    const float2 uvc = __ldg(reinterpret_cast<const float2*>(m.uvs) + ic);
    const float  w   = 1.f - isec.u - isec.v;  // triangle.zig:133-140
    frag.u           = __fmaf_rn(uva.x, w, __fmaf_rn(uvb.x, isec.u, uvc.x * isec.v));
    frag.v           = __fmaf_rn(uva.y, w, __fmaf_rn(uvb.y, isec.u, uvc.y * isec.v));

    const V3 nb = decompressNormal(m.normals, ib);
    const V3 na = decompressNormal(m.normals, ia);
    const V3 nc = decompressNormal(m.normals, ic);
    const V3 ni = normalize3(interpolate3(na, nb, nc, isec.u, isec.v));

    // triangle.positionDifferentials, triangle.zig:102-131
    const float duv02x = uva.x - uvc.x, duv02y = uva.y - uvc.y;
    const float duv12x = uvb.x - uvc.x, duv12y = uvb.y - uvc.y;
    const float determinant = duv02x * duv12y - duv02y * duv12x;

    V3       dpdu, dpdv;
    const V3 dp02 = sub3(pa, pc);
    const V3 dp12 = sub3(pb, pc);
    if (0.f == fabsf(determinant)) {
        const V3 ng = normalize3(cross3(sub3(pc, pa), sub3(pb, pa)));
        if (fabsf(ng.x) > fabsf(ng.y)) {
            dpdu = divs3({-ng.z, 0.f, ng.x}, __fsqrt_rn(ng.x * ng.x + ng.z * ng.z));
        } else {
            dpdu = divs3({0.f, ng.z, -ng.y}, __fsqrt_rn(ng.y * ng.y + ng.z * ng.z));
        }
        dpdv = cross3(ng, dpdu);
    } else {
        const float invdet = __fdiv_rn(1.f, determinant);
        dpdu               = scale3(invdet, fmas3(duv12y, dp02, scale3(-duv02y, dp12)));
        dpdv               = scale3(invdet, fmas3(-duv12x, dp02, scale3(duv02x, dp12)));
    }

    const V3 t = normalize3(gramSchmidt(dpdu, ni));
    const V3 b = normalize3(gramSchmidt(dpdv, ni));

    frag.p = frag.trafo.objectToWorldPoint(p);
    frag.t = frag.trafo.objectToWorldNormal(t);
    frag.b = frag.trafo.objectToWorldNormal(b);
    frag.n = frag.trafo.objectToWorldNormal(ni);
}

// ---- mesh light sampling: triangle_mesh.zig:390-487, triangle.zig:82-100, sampling.zig:39-47 -----------------------

__device__ __forceinline__ V3 orthogonalize(V3 a, V3 b) { return normalize3(fmas3(-dot3(a, b), a, b)); }

__device__ __forceinline__ void barycentricCoords(V3 dir, V3 a, V3 b, V3 c, float& u, float& v) {
    const V3    e1      = sub3(b, a);
    const V3    e2      = sub3(c, a);
    const V3    tvec    = neg3(a);
    const V3    pvec    = cross3(dir, e2);
    const V3    qvec    = cross3(tvec, e1);
    const float e1_d_pv = dot3(e1, pvec);
    const float tv_d_pv = dot3(tvec, pvec);
    const float di_d_qv = dot3(dir, qvec);
    const float inv_det = __fdiv_rn(1.f, e1_d_pv);
    u                   = tv_d_pv * inv_det;
    v                   = di_d_qv * inv_det;
}

__device__ __forceinline__ float sphericalArea(V3 A, V3 B, V3 C, float& cos_alpha, float& alpha) {
    const V3 BA = orthogonalize(A, sub3(B, A));
    const V3 CA = orthogonalize(A, sub3(C, A));
    const V3 AB = orthogonalize(B, sub3(A, B));
    const V3 CB = orthogonalize(B, sub3(C, B));
    const V3 BC = orthogonalize(C, sub3(B, C));
    const V3 AC = orthogonalize(C, sub3(A, C));
    cos_alpha         = zclamp(dot3(BA, CA), -1.f, 1.f);
    alpha             = acosf(cos_alpha);
    const float beta  = acosf(zclamp(dot3(AB, CB), -1.f, 1.f));
    const float gamma = acosf(zclamp(dot3(BC, AC), -1.f, 1.f));
    return alpha + beta + gamma - kPi;
}

// Stratified Sampling of Spherical Triangles, James Arvo
__device__ __forceinline__ bool sampleSpherical(V3 pos, V3 pa, V3 pb, V3 pc, float r0, float r1, V3& dir, float& bu, float& bv, float& pdf) {
    const V3 pap = sub3(pa, pos);
    const V3 pbp = sub3(pb, pos);
    const V3 pcp = sub3(pc, pos);

    const V3 A = normalize3(pap);
    const V3 B = normalize3(pbp);
    const V3 C = normalize3(pcp);

    float       cos_alpha, alpha;
    const float sarea = sphericalArea(A, B, C, cos_alpha, alpha);
    if (0.f == sarea) return false;

    const float cos_c = zclamp(dot3(A, B), -1.f, 1.f);

    const float area_S      = r0 * sarea;
    const float angle_delta = area_S - alpha;
    const float p           = sinf(angle_delta);
    const float q           = cosf(angle_delta);

    const float sin_alpha = __fsqrt_rn(1.f - cos_alpha * cos_alpha);
    const float u         = q - cos_alpha;
    const float v         = p + sin_alpha * cos_c;

    const float s   = zclamp(__fdiv_rn((v * q - u * p) * cos_alpha - v, (v * p + u * q) * sin_alpha), -1.f, 1.f);
    const V3    C_s = add3(scale3(s, A), scale3(__fsqrt_rn(1.f - s * s), orthogonalize(A, C)));

    const float cs_b = dot3(C_s, B);
    const float z    = 1.f - r1 * (1.f - cs_b);
    const V3    P    = add3(scale3(z, B), scale3(__fsqrt_rn(1.f - z * z), orthogonalize(B, C_s)));

    dir = P;
    barycentricCoords(P, pap, pbp, pcp, bu, bv);
    pdf = __fdiv_rn(1.f, sarea);
    return true;
}

__device__ __forceinline__ float pdfSpherical(V3 pos, V3 pa, V3 pb, V3 pc) {
    float       cos_alpha, alpha;
    const float sarea = sphericalArea(normalize3(sub3(pa, pos)), normalize3(sub3(pb, pos)), normalize3(sub3(pc, pos)), cos_alpha, alpha);
    return __fdiv_rn(1.f, sarea);
}

__device__ __forceinline__ void triangleUniform(float u0, float u1, float& x, float& y) {  // E. Heitz
    if (u1 > u0) {
        x = 0.5f * u0;
        y = u1 - x;
        return;
    }
    y = 0.5f * u1;
    x = u0 - y;
}

constexpr float kAreaDistanceRatio = 0.001f;  // triangle_mesh.zig:489

// ---- materials -------------------------------------------------------------------------------

struct LutsD {  // ggx_integral.zig tables in the order of ZygpuScene.ggx_luts
    const float* base;
    __device__ float eM(float n_dot, float alpha) const { return lut2(base, 32, 32, n_dot, alpha); }
    __device__ float eMAvg(float alpha) const { return lut1(base + 1024, 32, alpha); }
    __device__ float e(float n_dot, float alpha, float f0) const { return lut3(base + 1056, 16, 16, 16, n_dot, alpha, f0); }
    __device__ float eAvg(float alpha, float f0) const { return lut2(base + 1056 + 4096, 16, 16, alpha, f0); }
    __device__ float eS(float n_dot, float alpha, float f0) const { return lut3(base + 1056 + 4096 + 256, 16, 16, 16, n_dot, alpha, f0); }
};

struct BxdfResult {  // bxdf.zig:8-19
    V3    reflection;
    float pdf;
};

enum : uint32_t { kScatterDiffuse = 0, kScatterGlossy = 1, kScatterSpecular = 2, kScatterNone = 3 };
enum : uint32_t { kEventReflection = 0, kEventTransmission = 1, kEventStraight = 2 };

struct BxdfSample {  // bxdf.zig:75-82
    V3       reflection, wi;
    float    pdf;
    float    split_weight;
    float    reg_alpha;  // Path
    uint32_t scattering, event;
};

constexpr float kMinRoughness = 0.01314f;  // ggx.zig:14

__device__ __forceinline__ V3 schlickF(V3 f0, float wo_dot_h) {  // fresnel.zig:16-18
    const float p = pow5(1.f - wo_dot_h);
    return {__fmaf_rn(p, 1.f - f0.x, f0.x), __fmaf_rn(p, 1.f - f0.y, f0.y), __fmaf_rn(p, 1.f - f0.z, f0.z)};
}

__device__ __forceinline__ float pdfVisible(float d, float g1_wo) { return __fdiv_rn(0.5f * d, g1_wo); }  // ggx.zig:437-439

// ggx.zig:34-46
__device__ __forceinline__ V3 dspbrMicroEc(const LutsD& luts, V3 f0, float n_dot_wi, float n_dot_wo, float alpha) {
    const float e_wo  = luts.eM(n_dot_wo, alpha);
    const float e_wi  = luts.eM(n_dot_wi, alpha);
    const float e_avg = luts.eMAvg(alpha);

    const float m = __fdiv_rn((1.f - e_wo) * (1.f - e_wi), kPi * (1.f - e_avg));

    const V3 f_avg = {__fmaf_rn(20.f / 21.f, f0.x, 1.f / 21.f), __fmaf_rn(20.f / 21.f, f0.y, 1.f / 21.f), __fmaf_rn(20.f / 21.f, f0.z, 1.f / 21.f)};
    const float om = 1.f - e_avg;
    const V3 f = {__fdiv_rn((f_avg.x * f_avg.x) * e_avg, __fmaf_rn(-f_avg.x, om, 1.f)),
                  __fdiv_rn((f_avg.y * f_avg.y) * e_avg, __fmaf_rn(-f_avg.y, om, 1.f)),
                  __fdiv_rn((f_avg.z * f_avg.z) * e_avg, __fmaf_rn(-f_avg.z, om, 1.f))};
    return scale3(m, f);
}

// Aniso.sample, ggx.zig:393-409
__device__ __forceinline__ V3 sampleVndf(V3 wo, float ax, float ay, float xi0, float xi1, const FrameD& frame, float& n_dot_h) {
    const V3 wo_l = frame.worldToFrame(wo);
    const V3 v    = normalize3({ax * wo_l.x, ay * wo_l.y, wo_l.z});

    const float phi       = (2.f * kPi) * xi0;
    const float z         = __fmaf_rn(1.f - xi1, 1.f + v.z, -v.z);
    const float sin_theta = __fsqrt_rn(saturate(1.f - z * z));
    float       sp, cp;
    sincosf(phi, &sp, &cp);
    const float x = sin_theta * cp;
    const float y = sin_theta * sp;

    const V3 h = add3({x, y, z}, v);
    const V3 m = normalize3({ax * h.x, ay * h.y, h.z});

    n_dot_h = safeClamp(m.z);
    return frame.frameToWorld(m);
}

__device__ __forceinline__ float isoDistribution(float n_dot_h, float a2) {  // ggx.zig:235-238
    const float d = __fmaf_rn(n_dot_h * n_dot_h, a2 - 1.f, 1.f);
    return __fdiv_rn(a2, kPi * d * d);
}
__device__ __forceinline__ void isoVisibilityAndG1Wo(float n_dot_wi, float n_dot_wo, float alpha2, float& vis, float& g1) {  // :240-250
    const float t_wi = __fsqrt_rn(__fmaf_rn(1.f - alpha2, n_dot_wi * n_dot_wi, alpha2));
    const float t_wo = __fsqrt_rn(__fmaf_rn(1.f - alpha2, n_dot_wo * n_dot_wo, alpha2));
    vis              = __fdiv_rn(0.5f, n_dot_wi * t_wo + n_dot_wo * t_wi);
    g1               = t_wo + n_dot_wo;
}
__device__ __forceinline__ float anisoDistribution(float n_dot_h, float x_dot_h, float y_dot_h, float ax, float ay) {  // :411-419
    const float x = __fdiv_rn(x_dot_h * x_dot_h, ax * ax);
    const float y = __fdiv_rn(y_dot_h * y_dot_h, ay * ay);
    const float d = (x + y) + (n_dot_h * n_dot_h);
    return __fdiv_rn(1.f, kPi * (ax * ay) * (d * d));
}
__device__ __forceinline__ void anisoVisibilityAndG1Wo(float t_dot_wi, float t_dot_wo, float b_dot_wi, float b_dot_wo, float n_dot_wi,
                                                       float n_dot_wo, float ax, float ay, float& vis, float& g1) {  // :421-434
    const float t_wo = length3({ax * t_dot_wo, ay * b_dot_wo, n_dot_wo});
    const float t_wi = length3({ax * t_dot_wi, ay * b_dot_wi, n_dot_wi});
    vis              = __fdiv_rn(0.5f, n_dot_wi * t_wo + n_dot_wo * t_wi);
    g1               = t_wo + n_dot_wo;
}

// Aniso.reflectionF (-> Iso.reflectionF when isotropic), ggx.zig:268-305, 73-95
__device__ __forceinline__ BxdfResult ggxReflection(V3 wi, V3 wo, V3 h, float n_dot_wi, float n_dot_wo, float wo_dot_h, float ax,
                                                    float ay, V3 f0, const FrameD& frame) {
    float d, vis, g1;
    if (ax == ay) {
        const float alpha2  = ax * ax;
        const float n_dot_h = saturate(dot3(frame.z, h));
        d                   = isoDistribution(n_dot_h, alpha2);
        isoVisibilityAndG1Wo(n_dot_wi, n_dot_wo, alpha2, vis, g1);
    } else {
        const float n_dot_h = saturate(dot3(frame.z, h));
        const float x_dot_h = dot3(frame.x, h);
        const float y_dot_h = dot3(frame.y, h);
        d                   = anisoDistribution(n_dot_h, x_dot_h, y_dot_h, ax, ay);
        anisoVisibilityAndG1Wo(dot3(frame.x, wi), dot3(frame.x, wo), dot3(frame.y, wi), dot3(frame.y, wo), n_dot_wi, n_dot_wo, ax, ay,
                               vis, g1);
    }
    const V3 f = schlickF(f0, wo_dot_h);
    return {scale3(d * vis, f), pdfVisible(d, g1)};
}

struct MicroD {  // ggx.zig:52-56
    V3    h;
    float n_dot_wi, h_dot_wi;
};

// Aniso.reflect (-> Iso.reflect when isotropic), ggx.zig:307-353, 97-126
__device__ __forceinline__ MicroD ggxReflect(V3 wo, float n_dot_wo, float ax, float ay, float specular_threshold, float xi0, float xi1,
                                             V3 f0, const FrameD& frame, BxdfSample& result) {
    float    n_dot_h;
    const V3 h = sampleVndf(wo, ax, ay, xi0, xi1, frame, n_dot_h);

    float x_dot_h = 0.f, y_dot_h = 0.f;
    if (ax != ay) {
        x_dot_h = dot3(frame.x, h);
        y_dot_h = dot3(frame.y, h);
    }

    const float wo_dot_h = clampDot(wo, h);
    const V3    wi       = normalize3(fmas3(2.f * wo_dot_h, h, neg3(wo)));
    const float n_dot_wi = frame.clampNdot(wi);

    float d, vis, g1;
    if (ax == ay) {
        const float alpha2 = ax * ax;
        d                  = isoDistribution(n_dot_h, alpha2);
        isoVisibilityAndG1Wo(n_dot_wi, n_dot_wo, alpha2, vis, g1);
    } else {
        d = anisoDistribution(n_dot_h, x_dot_h, y_dot_h, ax, ay);
        anisoVisibilityAndG1Wo(dot3(frame.x, wi), dot3(frame.x, wo), dot3(frame.y, wi), dot3(frame.y, wo), n_dot_wi, n_dot_wo, ax, ay,
                               vis, g1);
    }
    const V3 f = schlickF(f0, wo_dot_h);

    result.reflection = scale3(d * vis, f);
    result.wi         = wi;
    result.pdf        = pdfVisible(d, g1);
    const float a     = ax == ay ? ax : ay;  // Path.reflection(alpha | alpha[1], threshold)
    result.reg_alpha  = a;
    result.scattering = a <= specular_threshold ? kScatterSpecular : kScatterGlossy;
    result.event      = kEventReflection;
    return {h, n_dot_wi, wo_dot_h};
}

// diffuse.Micro, diffuse.zig:42-116
__device__ __forceinline__ float diffuseEstimateContribution(const LutsD& luts, float alpha, float f0, float albedo) {
    const float e_avg = luts.eAvg(alpha, f0);
    const float a     = e_avg;
    const float b     = __fdiv_rn(1.f, kPi * (1.f - e_avg)) * albedo;
    return __fdiv_rn(b, a + b);
}
__device__ __forceinline__ V3 diffuseEvaluate(const LutsD& luts, V3 color, float n_dot_wi, float n_dot_wo, float alpha, float f0) {
    const float e_wo  = luts.e(n_dot_wo, alpha, f0);
    const float e_wi  = luts.e(n_dot_wi, alpha, f0);
    const float e_avg = luts.eAvg(alpha, f0);
    return scale3(__fdiv_rn((1.f - e_wo) * (1.f - e_wi), kPi * (1.f - e_avg)), color);
}


// ---- Glass helpers: fresnel.zig:5-7, 31-43; ggx.zig:30-32, 128-257, 441-449 --------------------------------------

__device__ __forceinline__ float safeClampAbs(float x) { return zclamp(fabsf(x), kDotMin, 1.f); }
__device__ __forceinline__ float schlick1(float wo_dot_h, float f0) { return __fmaf_rn(pow5(1.f - wo_dot_h), 1.f - f0, f0); }
__device__ __forceinline__ float fresnelDielectric(float cos_theta_i, float cos_theta_t, float eta_i, float eta_t) {
    const float t0  = eta_t * cos_theta_i;
    const float t1  = eta_i * cos_theta_t;
    const float r_p = __fdiv_rn(t0 - t1, t0 + t1);
    const float t2  = eta_i * cos_theta_i;
    const float t3  = eta_t * cos_theta_t;
    const float r_o = __fdiv_rn(t2 - t3, t2 + t3);
    return 0.5f * (r_p * r_p + r_o * r_o);
}
__device__ __forceinline__ float ilmEpDielectric(const LutsD& luts, float n_dot_wo, float alpha, float f0) {
    return __fdiv_rn(1.f, luts.eS(n_dot_wo, alpha, f0 * 4.f));
}
__device__ __forceinline__ float gSmithCorrelated(float n_dot_wi, float n_dot_wo, float alpha2) {
    const float a = n_dot_wo * __fsqrt_rn(__fmaf_rn(1.f - alpha2, n_dot_wi * n_dot_wi, alpha2));
    const float b = n_dot_wi * __fsqrt_rn(__fmaf_rn(1.f - alpha2, n_dot_wo * n_dot_wo, alpha2));
    return __fdiv_rn(2.f * n_dot_wi * n_dot_wo, a + b);
}
__device__ __forceinline__ float pdfVisibleRefract(float n_dot_wo, float wo_dot_h, float d, float alpha2) {
    const float g1 = __fdiv_rn(2.f * n_dot_wo, n_dot_wo + __fsqrt_rn(alpha2 + (1.f - alpha2) * (n_dot_wo * n_dot_wo)));
    return __fdiv_rn(g1 * wo_dot_h * d, n_dot_wo);
}

struct IorD {  // sample_base.zig:106-117
    float eta_t, eta_i;
};

// Iso.refractionF, ggx.zig:128-159: returns the reflection scalar, pdf and the fresnel term
__device__ __forceinline__ void isoRefractionF(float n_dot_wi, float n_dot_wo, float wi_dot_h, float wo_dot_h, float n_dot_h, float alpha,
                                               IorD ior, float f0, float& refl, float& pdf, float& f) {
    const float alpha2       = alpha * alpha;
    const float abs_wi_dot_h = safeClampAbs(wi_dot_h);
    const float abs_wo_dot_h = safeClampAbs(wo_dot_h);

    const float d = isoDistribution(n_dot_h, alpha2);
    const float g = gSmithCorrelated(n_dot_wi, n_dot_wo, alpha2);

    const float cos_x = ior.eta_i > ior.eta_t ? abs_wi_dot_h : abs_wo_dot_h;
    f                 = 1.f - schlick1(cos_x, f0);

    const float sqr_eta_t = ior.eta_t * ior.eta_t;
    const float factor    = __fdiv_rn(abs_wi_dot_h * abs_wo_dot_h, n_dot_wi * n_dot_wo);
    const float sum       = ior.eta_i * wo_dot_h + ior.eta_t * wi_dot_h;
    const float denom     = sum * sum;

    const float refr = d * g * f;
    refl             = __fdiv_rn(factor * sqr_eta_t, denom) * refr;

    const float p = pdfVisibleRefract(n_dot_wo, abs_wo_dot_h, d, alpha2);
    pdf           = p * __fdiv_rn(abs_wi_dot_h * sqr_eta_t, denom);
}

// Iso.reflectNoFresnel, ggx.zig:161-187
__device__ __forceinline__ float isoReflectNoFresnel(V3 wo, V3 h, float n_dot_wo, float n_dot_h, float wo_dot_h, float alpha,
                                                     float specular_threshold, const FrameD& frame, BxdfSample& result) {
    const V3    wi       = normalize3(fmas3(2.f * wo_dot_h, h, neg3(wo)));
    const float n_dot_wi = frame.clampNdot(wi);
    const float alpha2   = alpha * alpha;

    const float d = isoDistribution(n_dot_h, alpha2);
    float       vis, g1;
    isoVisibilityAndG1Wo(n_dot_wi, n_dot_wo, alpha2, vis, g1);

    result.reflection = splat3(d * vis);
    result.wi         = wi;
    result.pdf        = pdfVisible(d, g1);
    result.reg_alpha  = alpha;
    result.scattering = alpha <= specular_threshold ? kScatterSpecular : kScatterGlossy;
    result.event      = kEventReflection;
    return n_dot_wi;
}

// Iso.refractNoFresnel, ggx.zig:189-233
__device__ __forceinline__ float isoRefractNoFresnel(V3 wo, V3 h, float n_dot_wo, float n_dot_h, float wi_dot_h, float wo_dot_h, float alpha,
                                                     float specular_threshold, IorD ior, const FrameD& frame, BxdfSample& result) {
    const float eta = __fdiv_rn(ior.eta_i, ior.eta_t);

    const float abs_wi_dot_h = safeClampAbs(wi_dot_h);
    const float abs_wo_dot_h = safeClampAbs(wo_dot_h);

    const V3 wi = normalize3(sub3(scale3(__fmaf_rn(eta, abs_wo_dot_h, -abs_wi_dot_h), h), scale3(eta, wo)));

    const float n_dot_wi = frame.clampAbsNdot(wi);
    const float alpha2   = alpha * alpha;

    const float d = isoDistribution(n_dot_h, alpha2);
    const float g = gSmithCorrelated(n_dot_wi, n_dot_wo, alpha2);

    const float refr      = d * g;
    const float factor    = __fdiv_rn(abs_wi_dot_h * abs_wo_dot_h, n_dot_wi * n_dot_wo);
    const float sum       = ior.eta_i * wo_dot_h + ior.eta_t * wi_dot_h;
    const float denom     = sum * sum;
    const float sqr_eta_t = ior.eta_t * ior.eta_t;
    const float pdf       = pdfVisibleRefract(n_dot_wo, abs_wo_dot_h, d, alpha2);

    result.reflection = splat3(__fdiv_rn(factor * sqr_eta_t, denom) * refr);
    result.wi         = wi;
    result.pdf        = pdf * __fdiv_rn(abs_wi_dot_h * sqr_eta_t, denom);
    result.reg_alpha  = alpha;
    result.scattering = alpha <= specular_threshold ? kScatterSpecular : kScatterGlossy;
    result.event      = kEventTransmission;
    return n_dot_wi;
}

enum : uint32_t { kSampleLight = 0, kSampleSubstitute = 1, kSampleGlass = 2 };

// Material sample of {Substitute surface, Light, Glass}: material_sample.zig, substitute_sample.zig:20-410, glass_sample.zig:20-537
struct MatSampleD {
    uint32_t kind;
    bool   lower_priority;
    float  ior, ior_outside, glass_f0;  // Glass
    bool   can_evaluate, avoid_caustics, translucent;
    FrameD frame;
    V3     geo_n, n, wo;
    float  ax, ay;
    V3     albedo, f0;
    float  metallic, specular, specular_threshold, opacity;
    // Coating of a Substitute, substitute/substitute_coating.zig (weight 1: the scale texture is the uniform 1). Only the instances
    // compiled with Coated = true (the ones that read image maps) look at these.
    V3     coat_n, coat_absorption;
    float  coat_thickness, coat_f0, coat_alpha;

    __device__ bool sameHemisphere(V3 v) const { return dot3(geo_n, v) > 0.f; }

    struct CoatResult {
        V3    reflection, attenuation;
        float f, pdf;
    };
    __device__ V3 coatAttenuation3(float distance) const {  // ccoef.attenuation3, collision_coefficients.zig:64-66
        const float nd = -distance;
        return {expf(nd * coat_absorption.x), expf(nd * coat_absorption.y), expf(nd * coat_absorption.z)};
    }
    __device__ V3 coatAttenuation(float n_dot_wi, float n_dot_wo) const {  // substitute_coating.zig:102-108
        const float f = 1.f * schlick1(zmin(n_dot_wi, n_dot_wo), coat_f0);
        const float d = coat_thickness * (__fdiv_rn(1.f, n_dot_wi) + __fdiv_rn(1.f, n_dot_wo));
        return scale3(1.f - f, coatAttenuation3(d));
    }
    // Coating.evaluate, :35-58
    __device__ CoatResult coatEvaluate(const LutsD& luts, V3 wi, V3 h, float wo_dot_h, bool avoid) const {
        const float n_dot_wi = clampDot(coat_n, wi);
        const float n_dot_wo = clampAbsDot(coat_n, wo);
        const V3    att      = coatAttenuation(n_dot_wi, n_dot_wo);
        if (avoid && coat_alpha <= specular_threshold) return {splat3(0.f), att, 0.f, 0.f};
        // ggx.Iso.reflectionF, ggx.zig:73-95
        const float alpha2  = coat_alpha * coat_alpha;
        const float n_dot_h = saturate(dot3(coat_n, h));
        const float d       = isoDistribution(n_dot_h, alpha2);
        float       vis, g1;
        isoVisibilityAndG1Wo(n_dot_wi, n_dot_wo, alpha2, vis, g1);
        const float f  = schlick1(wo_dot_h, coat_f0);
        const float ep = ilmEpDielectric(luts, n_dot_wo, coat_alpha, coat_f0);
        return {splat3(((ep * 1.f) * n_dot_wi) * ((d * vis) * f)), att, f, pdfVisible(d, g1)};
    }
    // Coating.reflect, :60-82
    __device__ V3 coatReflect(const LutsD& luts, V3 h, float n_dot_wo, float n_dot_h, float wo_dot_h, BxdfSample& result) const {
        FrameD cf;
        cf.z = coat_n;
        orthonormalBasis3(coat_n, cf.x, cf.y);
        const float n_dot_wi = isoReflectNoFresnel(wo, h, n_dot_wo, n_dot_h, wo_dot_h, coat_alpha, specular_threshold, cf, result);
        const float ep       = ilmEpDielectric(luts, n_dot_wo, coat_alpha, coat_f0);
        result.reflection    = scale3((ep * 1.f) * n_dot_wi, result.reflection);
        return coatAttenuation(n_dot_wi, n_dot_wo);
    }

    // substitute_sample.zig:236-278
    __device__ BxdfResult baseEvaluate(const LutsD& luts, V3 wi, V3 h, float wo_dot_h, bool force_disable_caustics) const {
        const float n_dot_wi = frame.clampNdot(wi);
        const float n_dot_wo = frame.clampAbsNdot(wo);

        BxdfResult d  = {splat3(0.f), 0.f};
        float      dw = 0.f;

        if (1.f != metallic) {
            const V3    a   = scale3(opacity, albedo);
            const float f0m = hmax3(f0);
            d               = {diffuseEvaluate(luts, a, n_dot_wi, n_dot_wo, ay, f0m), n_dot_wi * kPiInv};
            const float am  = hmax3(albedo);
            dw              = diffuseEstimateContribution(luts, ay, f0m, am);
        }

        if ((force_disable_caustics || avoid_caustics) && ay <= specular_threshold) {
            return {scale3(n_dot_wi, d.reflection), dw * d.pdf};
        }

        const BxdfResult gg  = ggxReflection(wi, wo, h, n_dot_wi, n_dot_wo, wo_dot_h, ax, ay, f0, frame);
        const V3         mms = dspbrMicroEc(luts, f0, n_dot_wi, n_dot_wo, ay);

        const float pdf = dw * d.pdf + (1.f - dw) * gg.pdf;
        return {scale3(n_dot_wi, add3(d.reflection, scale3(specular, add3(gg.reflection, mms)))), pdf};
    }

    // material_sample.zig:56-62 -> substitute_sample.zig:88-145
    // `Glass` = the scene holds Glass materials (compiled out otherwise)
    template <bool Glass, bool Coated = false>
    __device__ BxdfResult evaluate(const LutsD& luts, V3 wi, uint32_t max_splits) const {
        if (kSampleLight == kind) return {splat3(0.f), 0.f};
        if (Glass && kSampleGlass == kind) return glassEvaluate(luts, wi, max_splits);
        if (!sameHemisphere(wo)) return {splat3(0.f), 0.f};
        const V3    h        = normalize3(add3(wo, wi));
        const float wo_dot_h = clampDot(wo, h);
        const BxdfResult base_result = baseEvaluate(luts, wi, h, wo_dot_h, false);
        if (Coated && coat_thickness > 0.f) {  // substitute_sample.zig:138-142
            const CoatResult c   = coatEvaluate(luts, wi, h, wo_dot_h, avoid_caustics);
            const float      pdf = c.f * c.pdf + (1.f - c.f) * base_result.pdf;
            return {add3(c.reflection, mul3(c.attenuation, base_result.reflection)), pdf};
        }
        return base_result;
    }

    // substitute_sample.zig:338-361
    __device__ MicroD diffuseSample(const LutsD& luts, float diffuse_weight, float xi0, float xi1, BxdfSample& result) const {
        const float n_dot_wo = frame.clampAbsNdot(wo);
        const V3    a        = scale3(opacity, albedo);
        const float f0m      = hmax3(f0);

        // diffuse.Micro.reflect, diffuse.zig:82-106
        const V3 is = hemisphereCosine(xi0, xi1);
        const V3 wi = normalize3(frame.frameToWorld(is));
        const V3 h  = normalize3(add3(wo, wi));

        const float h_dot_wi = clampDot(h, wi);
        const float n_dot_wi = frame.clampNdot(wi);

        result.reflection = diffuseEvaluate(luts, a, n_dot_wi, n_dot_wo, ay, f0m);
        result.wi         = wi;
        result.pdf        = n_dot_wi * kPiInv;
        result.reg_alpha  = 1.f;  // Path.diffuseReflection
        result.scattering = kScatterDiffuse;
        result.event      = kEventReflection;

        const BxdfResult gg  = ggxReflection(wi, wo, h, n_dot_wi, n_dot_wo, h_dot_wi, ax, ay, f0, frame);
        const V3         mms = dspbrMicroEc(luts, f0, n_dot_wi, frame.clampNdot(wo), ay);

        result.reflection = scale3(n_dot_wi, add3(result.reflection, scale3(specular, add3(gg.reflection, mms))));
        result.pdf        = diffuse_weight * result.pdf + (1.f - diffuse_weight) * gg.pdf;
        return {h, n_dot_wi, h_dot_wi};
    }

    // substitute_sample.zig:363-410 (no flakes)
    __device__ MicroD glossSample(const LutsD& luts, float diffuse_weight, float xi0, float xi1, BxdfSample& result) const {
        const float n_dot_wo = frame.clampAbsNdot(wo);

        const MicroD micro = ggxReflect(wo, n_dot_wo, ax, ay, specular_threshold, xi0, xi1, f0, frame, result);
        const V3     mms   = dspbrMicroEc(luts, f0, micro.n_dot_wi, frame.clampNdot(wo), ay);

        BxdfResult d = {splat3(0.f), 0.f};
        if (diffuse_weight > 0.f) {
            const V3    a   = scale3(opacity, albedo);
            const float f0m = hmax3(f0);
            d               = {diffuseEvaluate(luts, a, micro.n_dot_wi, n_dot_wo, ay, f0m), micro.n_dot_wi * kPiInv};
        }

        result.reflection = scale3(micro.n_dot_wi, add3(scale3(specular, add3(result.reflection, mms)), d.reflection));
        result.pdf        = (1.f - diffuse_weight) * result.pdf + diffuse_weight * d.pdf;
        return micro;
    }

    // material_sample.zig:64-78 -> substitute_sample.zig:147-234, 280-302. Returns the number of samples (0 or 1).
    template <bool Glass, bool Coated = false>
    __device__ uint32_t sample(const LutsD& luts, SamplerD& sampler, uint32_t max_splits, BxdfSample* results) const {
        if (kSampleLight == kind) return 0;
        if (Glass && kSampleGlass == kind) return glassSample(luts, sampler, max_splits, results);
        if (!sameHemisphere(wo)) return 0;
        BxdfSample& result  = results[0];
        result.split_weight = 1.f;

        if (Coated && coat_thickness > 0.f) {  // coatingSample, substitute_sample.zig:304-336, 412-433
            float xi0, xi1;
            sampler.sample2D(xi0, xi1);
            // Coating.sample, substitute_coating.zig:84-90
            FrameD cf;
            cf.z = coat_n;
            orthonormalBasis3(coat_n, cf.x, cf.y);
            float       n_dot_h;
            const V3    h        = sampleVndf(wo, coat_alpha, coat_alpha, xi0, xi1, cf, n_dot_h);
            const float h_dot_wi = clampDot(wo, h);
            const float f        = schlick1(h_dot_wi, coat_f0);

            const V3    s3 = sampler.sample3D();
            const float p  = s3.x;
            if (p <= f) {  // coatingReflect
                const float n_dot_wo            = clampAbsDot(coat_n, wo);
                const V3    coating_attenuation = coatReflect(luts, h, n_dot_wo, n_dot_h, h_dot_wi, result);
                const BxdfResult base_result    = baseEvaluate(luts, result.wi, h, h_dot_wi, false);
                result.reflection = add3(scale3(f, result.reflection), mul3(coating_attenuation, base_result.reflection));
                result.pdf        = f * result.pdf + (1.f - f) * base_result.pdf;
            } else {
                float dw = 0.f;
                if (1.f != metallic) dw = diffuseEstimateContribution(luts, ay, hmax3(f0), hmax3(albedo));
                const float  p1    = __fdiv_rn(p - f, 1.f - f);
                const MicroD micro = p1 < dw ? diffuseSample(luts, dw, s3.y, s3.z, result) : glossSample(luts, dw, s3.y, s3.z, result);
                // coatingBaseSample
                const CoatResult c = coatEvaluate(luts, result.wi, micro.h, micro.h_dot_wi, avoid_caustics);
                result.reflection  = add3(mul3(c.attenuation, result.reflection), c.reflection);
                result.pdf         = (1.f - f) * result.pdf + f * c.pdf;
            }
            if (0.f == result.pdf) return 0;
            return 1;
        }

        float dw = 0.f;
        if (1.f != metallic) {
            const float f0m = hmax3(f0);
            const float am  = hmax3(albedo);
            dw              = diffuseEstimateContribution(luts, ay, f0m, am);
        }

        const V3    s3 = sampler.sample3D();
        const float p  = s3.x;
        if (p < dw) {
            diffuseSample(luts, dw, s3.y, s3.z, result);
        } else {
            glossSample(luts, dw, s3.y, s3.z, result);
        }
        if (0.f == result.pdf) return 0;
        return 1;
    }

    // ---- Glass (thickness == 0, abbe == 0) ----

    // Sample.evaluate, glass_sample.zig:68-152 (force_disable_caustics = false)
    __device__ BxdfResult glassEvaluate(const LutsD& luts, V3 wi, uint32_t max_splits) const {
        const float alpha = ax;
        const bool  rough = alpha > 0.f;
        if (ior == ior_outside || !rough || lower_priority || (avoid_caustics && alpha <= specular_threshold)) return {splat3(0.f), 0.f};

        const bool  split = max_splits > 1;
        const float s     = specular;

        if (!sameHemisphere(wo)) {
            const IorD io{ior_outside, ior};  // eta_i = self.ior, eta_t = self.ior_outside
            const V3   h = neg3(normalize3(add3(scale3(io.eta_t, wi), scale3(io.eta_i, wo))));

            const float wi_dot_h = dot3(wi, h);
            if (wi_dot_h <= 0.f) return {splat3(0.f), 0.f};

            const float wo_dot_h = dot3(wo, h);
            const float eta      = __fdiv_rn(io.eta_i, io.eta_t);
            const float sint2    = (eta * eta) * (1.f - wo_dot_h * wo_dot_h);
            if (sint2 >= 1.f) return {splat3(0.f), 0.f};

            const float n_dot_wi = frame.clampNdot(wi);
            const float n_dot_wo = frame.clampAbsNdot(wo);
            const float n_dot_h  = saturate(dot3(frame.z, h));

            float refl, pdf, f;
            isoRefractionF(n_dot_wi, n_dot_wo, wi_dot_h, wo_dot_h, n_dot_h, alpha, io, glass_f0, refl, pdf, f);
            const float comp = ilmEpDielectric(luts, n_dot_wo, alpha, glass_f0);

            const float split_pdf = split ? 1.f : f;
            return {splat3((zmin(n_dot_wi, n_dot_wo) * comp * s) * refl), split_pdf * pdf};
        } else if (sameHemisphere(wi)) {
            const float n_dot_wi = frame.clampNdot(wi);
            const float n_dot_wo = frame.clampAbsNdot(wo);

            const V3    h        = normalize3(add3(wo, wi));
            const float wo_dot_h = clampDot(wo, h);

            // Iso.reflectionF, ggx.zig:73-95
            const float alpha2  = alpha * alpha;
            const float n_dot_h = saturate(dot3(frame.z, h));
            const float d       = isoDistribution(n_dot_h, alpha2);
            float       vis, g1;
            isoVisibilityAndG1Wo(n_dot_wi, n_dot_wo, alpha2, vis, g1);
            const float f    = schlick1(wo_dot_h, glass_f0);
            const float refl = (d * vis) * f;
            const float pdf  = pdfVisible(d, g1);

            const float comp = ilmEpDielectric(luts, n_dot_wo, alpha, glass_f0);

            const float split_pdf = split ? 1.f : f;
            return {splat3((n_dot_wi * comp * s) * refl), split_pdf * pdf};
        }
        return {splat3(0.f), 0.f};
    }

    __device__ static BxdfSample singular(V3 reflection, V3 wi, float split_weight, uint32_t event) {
        BxdfSample r;
        r.reflection   = reflection;
        r.wi           = wi;
        r.pdf          = 1.f;
        r.split_weight = split_weight;
        r.reg_alpha    = 0.f;
        r.scattering   = kScatterSpecular;
        r.event        = event;
        return r;
    }

    // specularSample, glass_sample.zig:202-284 (Thin = false, weight = 1)
    __device__ uint32_t glassSpecularSample(SamplerD& sampler, bool split, BxdfSample* buffer) const {
        float eta_i = ior_outside;
        float eta_t = ior;

        if (eta_i == eta_t || lower_priority) {
            buffer[0] = singular(splat3(1.f), neg3(wo), 1.f, kEventTransmission);
            return 1;
        }

        V3 nn = frame.z;
        if (!sameHemisphere(wo)) {
            nn            = neg3(nn);
            const float t = eta_i;
            eta_i         = eta_t;
            eta_t         = t;
        }

        const float n_dot_wo = zmin(fabsf(dot3(nn, wo)), 1.f);
        const float eta      = __fdiv_rn(eta_i, eta_t);
        const float sint2    = (eta * eta) * (1.f - n_dot_wo * n_dot_wo);

        float n_dot_t, f;
        if (sint2 >= 1.f) {
            n_dot_t = 0.f;
            f       = 1.f;
        } else {
            n_dot_t = __fsqrt_rn(1.f - sint2);
            f       = fresnelDielectric(n_dot_wo, n_dot_t, eta_i, eta_t);
        }

        const V3 reflected = normalize3(sub3(scale3(2.f * n_dot_wo, nn), wo));                             // reflect, :429-438
        const V3 refracted = normalize3(sub3(scale3(eta * n_dot_wo - n_dot_t, nn), scale3(eta, wo)));     // thickSpecularRefract, :519-537

        if (split) {
            buffer[0] = singular(splat3(specular), reflected, f, kEventReflection);
            if (1.f == f) return 1;
            buffer[1] = singular(splat3(1.f), refracted, 1.f - f, kEventTransmission);
            return 2;
        }
        const float p = sampler.sample1D();
        if (p <= f) {
            buffer[0] = singular(splat3(specular), reflected, 1.f, kEventReflection);
        } else {
            buffer[0] = singular(splat3(1.f), refracted, 1.f, kEventTransmission);
        }
        return 1;
    }

    // roughSample, glass_sample.zig:286-427 (Thin = false, weight = 1)
    __device__ uint32_t glassRoughSample(const LutsD& luts, SamplerD& sampler, bool split, BxdfSample* buffer) const {
        const IorD quo_ior{ior, ior_outside};  // eta_i = ior_outside, eta_t = ior

        if (fabsf(quo_ior.eta_i - quo_ior.eta_t) <= 2.e-7f || lower_priority) {
            buffer[0] = singular(splat3(1.f), neg3(wo), 1.f, kEventTransmission);
            return 1;
        }

        const float alpha     = ax;
        const bool  same_side = sameHemisphere(wo);

        const FrameD fr = same_side ? frame : FrameD{frame.x, frame.y, neg3(frame.z)};
        const IorD   io = same_side ? quo_ior : IorD{quo_ior.eta_i, quo_ior.eta_t};

        const V3 s3 = sampler.sample3D();

        float    n_dot_h;
        const V3 h = sampleVndf(wo, ax, ay, s3.y, s3.z, fr, n_dot_h);

        const float n_dot_wo = fr.clampAbsNdot(wo);
        const float wo_dot_h = clampDot(wo, h);

        const float eta   = __fdiv_rn(io.eta_i, io.eta_t);
        const float sint2 = (eta * eta) * (1.f - wo_dot_h * wo_dot_h);

        const float s = specular;

        float wi_dot_h, f;
        if (sint2 >= 1.f) {
            wi_dot_h = 0.f;
            f        = 1.f;
        } else {
            wi_dot_h          = __fsqrt_rn(1.f - sint2);
            const float cos_x = io.eta_i > io.eta_t ? wi_dot_h : wo_dot_h;
            f                 = schlick1(cos_x, glass_f0);
        }

        const float ep         = ilmEpDielectric(luts, n_dot_wo, alpha, glass_f0);
        const float r_wo_dot_h = same_side ? -wo_dot_h : wo_dot_h;  // roughRefract, :440-503 (Thin = false)

        if (split) {
            {
                const float n_dot_wi   = isoReflectNoFresnel(wo, h, n_dot_wo, n_dot_h, wo_dot_h, alpha, specular_threshold, fr, buffer[0]);
                buffer[0].reflection   = mul3(buffer[0].reflection, splat3(n_dot_wi * ep * s));
                buffer[0].split_weight = f;
            }
            if (1.f == f) return 1;
            {
                const float n_dot_wi = isoRefractNoFresnel(wo, h, n_dot_wo, n_dot_h, -wi_dot_h, r_wo_dot_h, alpha, specular_threshold, io, fr, buffer[1]);
                if (n_dot_wi < 0.f) return 1;
                buffer[1].reflection   = mul3(buffer[1].reflection, splat3(n_dot_wi * ep));
                buffer[1].split_weight = 1.f - f;
            }
            return 2;
        }

        BxdfSample& result = buffer[0];
        const float p      = s3.x;
        if (p <= f) {
            const float n_dot_wi = isoReflectNoFresnel(wo, h, n_dot_wo, n_dot_h, wo_dot_h, alpha, specular_threshold, fr, result);
            result.reflection    = mul3(result.reflection, splat3(f * n_dot_wi * ep * s));
            result.pdf *= f;
        } else {
            const float n_dot_wi = isoRefractNoFresnel(wo, h, n_dot_wo, n_dot_h, -wi_dot_h, r_wo_dot_h, alpha, specular_threshold, io, fr, result);
            if (n_dot_wi < 0.f) return 0;
            const float omf   = 1.f - f;
            result.reflection = mul3(result.reflection, splat3(omf * n_dot_wi * ep));
            result.pdf *= omf;
        }
        result.split_weight = 1.f;
        return 1;
    }

    // Sample.sample, glass_sample.zig:167-200
    __device__ uint32_t glassSample(const LutsD& luts, SamplerD& sampler, uint32_t max_splits, BxdfSample* buffer) const {
        const bool split = max_splits > 1;
        if (ax > 0.f) return glassRoughSample(luts, sampler, split, buffer);
        return glassSpecularSample(sampler, split, buffer);
    }
};

// Emittance.radiance, emittance.zig:29-59 (uniform emission, no profile)
__device__ __forceinline__ V3 emittanceRadiance(const ZygpuMaterial& m, V3 wi, const TrafoD& trafo, float area, bool in_camera) {
    if (-dot3(wi, trafo.r2) < m.emission_cos_a) return splat3(0.f);
    const float factor    = in_camera ? m.emission_camera_weight : 1.f;
    const V3    intensity = {m.emission[0] * 1.f, m.emission[1] * 1.f, m.emission[2] * 1.f};
    if (0.f != m.emission_normalize) return scale3(__fdiv_rn(factor, area), intensity);
    return scale3(factor, intensity);
}

// the same with an image emission_map: `texel` = ts.sample2D_3(emission_map, rs, ...), emittance.zig:48
__device__ __forceinline__ V3 emittanceRadianceMapped(const ZygpuMaterial& m, V3 wi, const TrafoD& trafo, float area, bool in_camera, V3 texel) {
    if (-dot3(wi, trafo.r2) < m.emission_cos_a) return splat3(0.f);
    const float factor    = in_camera ? m.emission_camera_weight : 1.f;
    const V3    intensity = {m.emission[0] * texel.x, m.emission[1] * texel.y, m.emission[2] * texel.z};
    if (0.f != m.emission_normalize) return scale3(__fdiv_rn(factor, area), intensity);
    return scale3(factor, intensity);
}

// Vertex.sample + Material.sample, vertex.zig:137-181, material.zig:184-194, substitute_material.zig:114-221
// `Coated`: the instance looks at the clear coat of a Substitute (the instances that read image maps; compiled out of the others)
template <bool Glass, bool Coated = false>
__device__ __forceinline__ MatSampleD materialSample(const ZygpuMaterial& m, const FragD& frag, V3 wo, float reg_weight, float reg_alpha,
                                                     bool caustics, float specular_threshold, float ior_outside = 1.f,
                                                     int highest_priority = -128) {
    MatSampleD r;
    r.wo             = wo;
    r.lower_priority = m.priority < highest_priority;
    if (0 != (m.flags & ZYG_MATERIAL_TWO_SIDED) && !frag.sameHemisphere(wo)) {
        r.geo_n = neg3(frag.geo_n);
        r.n     = neg3(frag.n);
    } else {
        r.geo_n = frag.geo_n;
        r.n     = frag.n;
    }
    r.frame          = {frag.t, frag.b, r.n};
    r.avoid_caustics = !caustics;
    r.translucent    = false;

    if (Glass && ZYG_MATERIAL_GLASS == m.type) {  // glass_material.zig:46-73, glass_sample.zig:32-66
        const float rr = 0.f == m.roughness ? 0.f : zmax(m.roughness, kMinRoughness);
        float       a  = rr * rr;
        if (!(0.f == reg_weight || (a <= specular_threshold && !caustics))) {  // Renderstate.regularizeAlpha
            const float k = 1.f - reg_weight * reg_alpha;
            a             = 1.f - ((1.f - a) * k);
        }
        const bool rough = a > 0.f;

        r.kind               = kSampleGlass;
        r.ax = r.ay          = a;
        r.can_evaluate       = rough && m.ior != ior_outside;
        r.translucent        = m.thickness > 0.f;
        r.ior                = m.ior;
        r.ior_outside        = ior_outside;
        const float t        = __fdiv_rn(m.ior - ior_outside, m.ior + ior_outside);
        r.glass_f0           = rough ? t * t : 0.f;
        r.specular           = m.specular;
        r.specular_threshold = specular_threshold;
        return r;
    }
    if (ZYG_MATERIAL_SUBSTITUTE != m.type) {  // Light (and Debug): Base.initTBN(rs, wo, 0, 0, false)
        r.kind         = kSampleLight;
        r.can_evaluate = false;
        r.ax = r.ay = 0.f;
        return r;
    }
    r.kind = kSampleSubstitute;

    const V3    color     = {m.color[0], m.color[1], m.color[2]};
    const float roughness = zmax(m.roughness, kMinRoughness);
    const float metallic  = m.metallic;

    float ax, ay;  // anisotropicAlpha, substitute_material.zig:299-306
    if (m.anisotropy > 0.f) {
        const float rv = zmax(roughness * (1.f - m.anisotropy), kMinRoughness);
        ax             = roughness * roughness;
        ay             = rv * rv;
    } else {
        ax = ay = roughness * roughness;
    }

    const float ior_medium = ior_outside;  // rs.ior: Vertex.iorOutside, vertex.zig:87-93
    // substitute_material.zig:128-134 with the uniform coating scale 1: weight 1
    const bool  coated      = Coated && m.coating_thickness > 0.f;
    const float coating_ior = zlerp(ior_medium, m.coating_ior, 1.f);
    const float ior_outer   = coated ? coating_ior : ior_medium;
    r.coat_thickness        = 0.f;

    // Renderstate.regularizeAlpha, renderstate.zig:58-68
    if (!(0.f == reg_weight || (ax <= specular_threshold && !caustics))) {
        const float k = 1.f - reg_weight * reg_alpha;
        ax            = 1.f - ((1.f - ax) * k);
        ay            = 1.f - ((1.f - ay) * k);
    }
    r.ax           = ax;
    r.ay           = ay;
    r.can_evaluate = m.ior != ior_medium;

    const float t  = __fdiv_rn(m.ior - ior_outer, m.ior + ior_outer);  // Schlick.IorToF0
    const float f0 = t * t;

    r.albedo             = scale3(1.f - metallic, color);
    r.f0                 = lerp3(splat3(f0), color, splat3(metallic));
    r.metallic           = metallic;
    r.specular           = m.specular;
    r.specular_threshold = specular_threshold;
    r.opacity            = 1.f;

    if (coated) {  // :164-181 (no coating normal map: the interpolated normal)
        r.coat_n          = r.n;
        r.coat_absorption = {m.coating_absorption[0], m.coating_absorption[1], m.coating_absorption[2]};
        r.coat_thickness  = 1.f * m.coating_thickness;
        const float ct    = __fdiv_rn(coating_ior - ior_medium, coating_ior + ior_medium);
        r.coat_f0         = ct * ct;
        const float cr    = zmax(m.coating_roughness, kMinRoughness);
        float       ca    = cr * cr;
        if (!(0.f == reg_weight || (ca <= specular_threshold && !caustics))) {
            const float k = 1.f - reg_weight * reg_alpha;
            ca            = 1.f - ((1.f - ca) * k);
        }
        r.coat_alpha = ca;
    }
    return r;
}

}  // namespace zygpu
