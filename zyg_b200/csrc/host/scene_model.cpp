#include "scene_model.hpp"

#include "light_tree_builder.hpp"

#include "bvh_builder.hpp"
#include "mesh_handle.hpp"

#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>

namespace zyg {

namespace {

constexpr float kPi = 3.14159265358979323846f;

constexpr float kMinRoughness = 0.01314f;  // ggx.zig:14
constexpr float kMinAlpha     = kMinRoughness * kMinRoughness;

inline float degreesToRadians(float d) { return d * (kPi / 180.f); }  // math.zig:114-116

Mat3x3 mulMat(const Mat3x3& a, const Mat3x3& b) {  // matrix3x3.zig:104-116
    Mat3x3 m;
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) m.r[i][j] = a.r[i][0] * b.r[0][j] + a.r[i][1] * b.r[1][j] + a.r[i][2] * b.r[2][j];
        m.r[i][3] = 0.f;
    }
    return m;
}

Mat3x3 init9(float m00, float m01, float m02, float m10, float m11, float m12, float m20, float m21, float m22) {
    return {{{{m00, m01, m02, 0.f}}, {{m10, m11, m12, 0.f}}, {{m20, m21, m22, 0.f}}}};
}

struct Mat4x4 {
    Vec4f r[4];
};

Mat4x4 compose(const Mat3x3& basis, Vec4f scale, Vec4f origin) {  // matrix4x4.zig:88-107
    Mat4x4 m;
    for (int i = 0; i < 3; ++i) {
        m.r[i] = {{basis.r[i][0] * scale[i], basis.r[i][1] * scale[i], basis.r[i][2] * scale[i], 0.f}};
    }
    m.r[3] = {{origin[0], origin[1], origin[2], 1.f}};
    return m;
}

AABB transformAabb(const AABB& box, const Mat4x4& m) {  // aabb.zig:147-166
    const Vec4f xa = m.r[0] * splat(box.b[0][0]);
    const Vec4f xb = m.r[0] * splat(box.b[1][0]);
    const Vec4f ya = m.r[1] * splat(box.b[0][1]);
    const Vec4f yb = m.r[1] * splat(box.b[1][1]);
    const Vec4f za = m.r[2] * splat(box.b[0][2]);
    const Vec4f zb = m.r[2] * splat(box.b[1][2]);
    const Vec4f mw = m.r[3];
    return {{min4(xa, xb) + min4(ya, yb) + min4(za, zb) + mw, max4(xa, xb) + max4(ya, yb) + max4(za, zb) + mw}};
}

Vec4f transformVector(const Mat3x3& m, Vec4f v) {  // matrix3x3.zig:118-131
    Vec4f result = splat(v[0]) * m.r[0];
    result       = mulAdd(splat(v[1]), m.r[1], result);
    return mulAdd(splat(v[2]), m.r[2], result);
}

ZygpuAabb packAabb(const AABB& b) {
    ZygpuAabb r;
    for (int i = 0; i < 4; ++i) {
        r.min[i] = b.b[0][i];
        r.max[i] = b.b[1][i];
    }
    return r;
}

Vec4f readVec4f3(const json::Value& v) {  // json.zig:128-138
    if (json::Value::Array == v.kind && v.array.size() >= 3) {
        return {{float(v.array[0].number), float(v.array[1].number), float(v.array[2].number), 0.f}};
    }
    return splat(float(v.number));
}

Vec4f readColor(const json::Value& v) {  // json.zig:251-278 (array / number / {"sRGB": ...})
    if (json::Value::Object == v.kind) {
        Vec4f rgb = splat(0.f);
        if (const json::Value* s = v.get("sRGB")) rgb = readVec4f3(*s);
        return sRGBtoAP1(rgb);
    }
    return sRGBtoAP1(readVec4f3(v));
}

ZygpuMaterial defaultMaterial(uint32_t type) {
    ZygpuMaterial m{};
    m.type                   = type;
    m.flags                  = 0;
    m.priority               = 0;
    m.emission_num_samples   = 1;
    m.emission_cos_a         = -1.f;
    m.emission_camera_weight = 1.f;
    m.emission_normalize     = 0.f;
    m.specular               = 1.f;
    m.ior                    = 1.f;
    m.emission_map           = ZYGPU_NULL;
    m.color_map              = ZYGPU_NULL;
    m.roughness_map = m.metallic_map = m.normal_map = ZYGPU_NULL;
    switch (type) {
        case ZYG_MATERIAL_SUBSTITUTE:  // substitute_material.zig:41-67
            m.color[0] = m.color[1] = m.color[2] = 0.5f;
            m.roughness = 0.8f;
            m.ior       = 1.46f;
            m.coating_ior       = 1.5f;  // coating_thickness 0: no coat
            m.coating_roughness = 0.2f;
            break;
        case ZYG_MATERIAL_LIGHT:  // light_material.zig:44
            m.emission[0] = m.emission[1] = m.emission[2] = m.emission[3] = 1.f;
            break;
        case ZYG_MATERIAL_GLASS:  // glass_material.zig:28-36
            m.ior                  = 1.46f;
            m.roughness            = 0.f;
            m.attenuation_distance = 1.f;
            break;
        default: break;
    }
    return m;
}

uint32_t readAddress(const json::Value& v) {  // material_provider.zig:624-632
    return (json::Value::String == v.kind && "Clamp" == v.string) ? 0u : 1u;
}

// TextureDescriptor of an image resource: "id" + "sampler" + "scale" (material_provider.zig:440-476, 579-600, 634-678)
void readTextureDescriptor(const json::Value* em, uint32_t& map_image, uint32_t mode[3], float map_scale[2]) {
    {
        if (json::Value::Object == em->kind) {
            if (const json::Value* id = em->get("id")) map_image = uint32_t(id->number);
            if (const json::Value* sa = em->get("sampler")) {
                if (const json::Value* f = sa->get("filter")) {
                    if (json::Value::String == f->kind && "Nearest" == f->string) mode[2] = 0;
                    if (json::Value::String == f->kind && "Linear" == f->string) mode[2] = 1;
                }
                if (const json::Value* a = sa->get("address")) {
                    if (json::Value::Array == a->kind && a->array.size() >= 2) {
                        mode[0] = readAddress(a->array[0]);
                        mode[1] = readAddress(a->array[1]);
                    } else {
                        mode[0] = mode[1] = readAddress(*a);
                    }
                }
            }
            if (const json::Value* sc = em->get("scale")) {
                if (json::Value::Array == sc->kind && sc->array.size() >= 2) {
                    map_scale[0] = float(sc->array[0].number);
                    map_scale[1] = float(sc->array[1].number);
                } else {
                    map_scale[0] = map_scale[1] = float(sc->number);
                }
            }
        }
    }
}

// loadEmittance, material_provider.zig:412-436 (no profile; `emission_map` as an image resource)
void loadEmittance(const json::Value& j, ZygpuMaterial& m, uint32_t& map_image, uint32_t mode[3], float map_scale[2]) {
    if (const json::Value* em = j.get("emission_map")) readTextureDescriptor(em, map_image, mode, map_scale);

    Vec4f color = splat(1.f);
    if (const json::Value* s = j.get("spectrum")) color = readColor(*s);
    const float value = json::readFloatMember(j, "value", 1.f);

    m.emission_normalize = json::readBoolMember(j, "normalize", false) ? 1.f : 0.f;
    for (int i = 0; i < 4; ++i) m.emission[i] = value * color[i];
    // profile_angle = degrees(pi) when there is no profile (emittance.zig:77-80)
    const float profile_angle = kPi * (180.f / kPi);
    m.emission_cos_a          = std::cos(degreesToRadians(json::readFloatMember(j, "angle", profile_angle)));
    m.emission_camera_weight  = json::readFloatMember(j, "camera_weight", 1.f);
    m.emission_num_samples    = std::min(json::readUIntMember(j, "num_samples", 1), 64u);
}

bool anyGreaterZero3(const float v[4]) { return v[0] > 0.f || v[1] > 0.f || v[2] > 0.f; }

}  // namespace

Vec4f sRGBtoAP1(Vec4f srgb) {
    const Vec4f r = splat(srgb[0]), g = splat(srgb[1]), b = splat(srgb[2]);
    return Vec4f{{0.61309732f, 0.07019422f, 0.02061560f, 0.f}} * r + Vec4f{{0.33952285f, 0.91635557f, 0.10956983f, 0.f}} * g +
           Vec4f{{0.04737928f, 0.01345259f, 0.86981512f, 0.f}} * b;
}

Mat3x3 quaternionToMat3x3(Vec4f q) {
    const Vec4f tq = q + q;
    const float xy = tq[1] * q[0];
    const float xz = tq[2] * q[0];
    const float yz = tq[1] * q[2];
    const Vec4f w  = tq * splat(q[3]);

    const Vec4f a = {{q[3], q[0], q[3], q[1]}};
    const Vec4f b = {{q[1], q[2], q[0], q[2]}};
    const Vec4f t = (a + b) * (a - b);

    return init9(t[0] + t[1], xy - w[2], xz + w[1], xy + w[2], t[2] + t[3], yz - w[0], xz - w[1], yz + w[0], t[2] - t[3]);
}

Vec4f quaternionFromMat3x3(const Mat3x3& m) {
    float t;
    Vec4f q;
    if (m.r[2][2] < 0.f) {
        if (m.r[0][0] > m.r[1][1]) {
            t = 1.f + m.r[0][0] - m.r[1][1] - m.r[2][2];
            q = {{t, m.r[0][1] + m.r[1][0], m.r[2][0] + m.r[0][2], m.r[2][1] - m.r[1][2]}};
        } else {
            t = 1.f - m.r[0][0] + m.r[1][1] - m.r[2][2];
            q = {{m.r[0][1] + m.r[1][0], t, m.r[1][2] + m.r[2][1], m.r[0][2] - m.r[2][0]}};
        }
    } else {
        if (m.r[0][0] < -m.r[1][1]) {
            t = 1.f - m.r[0][0] - m.r[1][1] + m.r[2][2];
            q = {{m.r[2][0] + m.r[0][2], m.r[1][2] + m.r[2][1], t, m.r[1][0] - m.r[0][1]}};
        } else {
            t = 1.f + m.r[0][0] + m.r[1][1] + m.r[2][2];
            q = {{m.r[2][1] - m.r[1][2], m.r[0][2] - m.r[2][0], m.r[1][0] - m.r[0][1], t}};
        }
    }
    return q * splat(0.5f / std::sqrt(t));
}

Mat3x3 rotationFromEulerDegrees(Vec4f xyz) {
    const float ax = degreesToRadians(xyz[0]), ay = degreesToRadians(xyz[1]), az = degreesToRadians(xyz[2]);
    const float cx = std::cos(ax), sx = std::sin(ax);
    const float cy = std::cos(ay), sy = std::sin(ay);
    const float cz = std::cos(az), sz = std::sin(az);
    const Mat3x3 rot_x = init9(1.f, 0.f, 0.f, 0.f, cx, -sx, 0.f, sx, cx);  // matrix3x3.zig:31-50
    const Mat3x3 rot_y = init9(cy, 0.f, sy, 0.f, 1.f, 0.f, -sy, 0.f, cy);
    const Mat3x3 rot_z = init9(cz, -sz, 0.f, sz, cz, 0.f, 0.f, 0.f, 1.f);
    return mulMat(mulMat(rot_z, rot_x), rot_y);
}

void decomposeMatrix(const float m[16], Transformation& out) {
    const Vec4f mx = {{m[0], m[1], m[2], m[3]}};
    const Vec4f my = {{m[4], m[5], m[6], m[7]}};
    const Vec4f mz = {{m[8], m[9], m[10], m[11]}};
    const float sx = length3(mx), sy = length3(my), sz = length3(mz);
    const Mat3x3 basis = {{mx / splat(sx), my / splat(sy), mz / splat(sz)}};
    out.scale    = {{sx, sy, sz, 0.f}};
    out.position = {{m[12], m[13], m[14], m[15]}};
    out.rotation = quaternionFromMat3x3(basis);
}

const std::vector<float>& ggxLuts(std::string& error) {
    static std::vector<float> luts;
    static std::string        load_error;
    static bool               tried = false;
    if (!tried) {
        tried = true;
        Dl_info info;
        std::string dir = ".";
        if (dladdr(reinterpret_cast<const void*>(&sRGBtoAP1), &info) && info.dli_fname) {
            const std::string path = info.dli_fname;
            const size_t      s    = path.find_last_of('/');
            if (std::string::npos != s) dir = path.substr(0, s);
        }
        const std::string file = dir + "/data/ggx_luts.f32";
        if (FILE* f = std::fopen(file.c_str(), "rb")) {
            luts.resize(ZYGPU_GGX_LUT_FLOATS);
            const size_t n = std::fread(luts.data(), sizeof(float), luts.size(), f);
            std::fclose(f);
            if (n != luts.size()) {
                luts.clear();
                load_error = file + ": truncated";
            }
        } else {
            load_error = file + ": cannot open";
        }
    }
    error = load_error;
    return luts;
}

SceneModel::SceneModel() : specular_threshold_(kMinAlpha) {
    clamp_[0] = clamp_[1] = clamp_[2] = FLT_MAX;
    // createFallbackMaterial, material_provider.zig:127-129: a Debug material at id 0 (capi.zig:94-101)
    materials_.push_back(defaultMaterial(ZYG_MATERIAL_DEBUG));
    emission_maps_.push_back(EmissionMapRec{});
    color_maps_.push_back(EmissionMapRec{});
    surface_maps_.push_back({});
}

bool SceneModel::updateMaterial(uint32_t id, const json::Value& material) {
    touch();
    if (id >= materials_.size()) return false;
    const json::Value* rendering = material.get("rendering");
    if (!rendering || json::Value::Object != rendering->kind) return false;

    ZygpuMaterial& m = materials_[id];

    for (const auto& entry : rendering->object) {
        const json::Value& v = entry.second;
        if ("Light" == entry.first && ZYG_MATERIAL_LIGHT == m.type) {  // updateLight, material_provider.zig:233-244
            for (const auto& e : v.object) {
                if ("emittance" == e.first) {
                    EmissionMapRec& em = emission_maps_[id];
                    uint32_t mode[3]   = {em.address_u, em.address_v, em.filter};
                    loadEmittance(e.second, m, em.image, mode, em.scale);
                    em.address_u = mode[0];
                    em.address_v = mode[1];
                    em.filter    = mode[2];
                } else if ("two_sided" == e.first) {
                    m.flags = e.second.boolean ? (m.flags | ZYG_MATERIAL_TWO_SIDED) : (m.flags & ~ZYG_MATERIAL_TWO_SIDED);
                } else {
                    warnings_.push_back("material " + std::to_string(id) + ": Light parameter \"" + e.first + "\" is not supported by the device path and is ignored");
                }
            }
        } else if ("Substitute" == entry.first && ZYG_MATERIAL_SUBSTITUTE == m.type) {  // updateSubstitute, :254-330
            for (const auto& e : v.object) {
                const std::string& k = e.first;
                if ("color" == k) {
                    // readValue(.Color): a texture description with an image id becomes the colour map, anything else a
                    // uniform colour (material_provider.zig:741-776)
                    EmissionMapRec& cm = color_maps_[id];
                    cm                 = EmissionMapRec{};
                    if (json::Value::Object == e.second.kind && e.second.get("id")) {
                        uint32_t mode[3] = {cm.address_u, cm.address_v, cm.filter};
                        readTextureDescriptor(&e.second, cm.image, mode, cm.scale);
                        cm.address_u = mode[0];
                        cm.address_v = mode[1];
                        cm.filter    = mode[2];
                    } else {
                        const Vec4f c = readColor(e.second);
                        for (int i = 0; i < 4; ++i) m.color[i] = c[i];
                    }
                } else if ("emittance" == k) {
                    EmissionMapRec em;  // Substitute emission maps are not in scope: parsed and dropped
                    uint32_t       mode[3] = {1, 1, 1};
                    loadEmittance(e.second, m, em.image, mode, em.scale);
                } else if ("roughness" == k || "metallic" == k || "normal" == k) {
                    // readValue(.Roughness / .Metallic / .Normal), material_provider.zig:269-287, 741-800: an image id makes it a map
                    const int       slot = "roughness" == k ? kRoughnessMap : ("metallic" == k ? kMetallicMap : kNormalMap);
                    EmissionMapRec& sm   = surface_maps_[id][slot];
                    sm                   = EmissionMapRec{};
                    if (json::Value::Object == e.second.kind && e.second.get("id")) {
                        uint32_t mode[3] = {sm.address_u, sm.address_v, sm.filter};
                        readTextureDescriptor(&e.second, sm.image, mode, sm.scale);
                        sm.address_u = mode[0];
                        sm.address_v = mode[1];
                        sm.filter    = mode[2];
                    } else if (kRoughnessMap == slot) {
                        m.roughness = float(e.second.number);
                    } else if (kMetallicMap == slot) {
                        m.metallic = float(e.second.number);
                    }  // a uniform "normal" is no normal map (Texture.initUniform2: isUniform, substitute_material.zig:157)
                } else if ("coating" == k && json::Value::Object == e.second.kind) {  // material_provider.zig:303-326
                    Vec4f coating_color{{1.f, 1.f, 1.f, 1.f}};
                    float coating_attenuation_distance = 0.1f;
                    for (const auto& c : e.second.object) {
                        if ("color" == c.first) {
                            coating_color = readColor(c.second);
                        } else if ("attenuation_distance" == c.first) {
                            coating_attenuation_distance = float(c.second.number);
                        } else if ("ior" == c.first) {
                            m.coating_ior = float(c.second.number);
                        } else if ("roughness" == c.first && json::Value::Number == c.second.kind) {
                            m.coating_roughness = float(c.second.number);
                        } else if ("thickness" == c.first) {
                            m.coating_thickness = float(c.second.number);
                        } else {  // normal / scale / roughness maps of the coat
                            warnings_.push_back("material " + std::to_string(id) + ": coating parameter \"" + c.first + "\" is not supported by the device path and is ignored");
                        }
                    }
                    // setCoatingAttenuation -> attenuationCoefficient, collision_coefficients.zig:35-44
                    for (int i = 0; i < 3; ++i) {
                        const float c           = fmin_(fmax_(coating_color[i], 0.01f), 0.991102f);
                        m.coating_absorption[i] = 0.f == coating_attenuation_distance ? 0.f : -std::log(c) / coating_attenuation_distance;
                    }
                } else if ("specular" == k) {
                    m.specular = float(e.second.number);
                } else if ("anisotropy" == k) {
                    m.anisotropy = float(e.second.number);
                } else if ("ior" == k) {
                    m.ior = float(e.second.number);
                } else if ("priority" == k) {
                    m.priority = int32_t(e.second.number);
                } else if ("two_sided" == k) {
                    m.flags = e.second.boolean ? (m.flags | ZYG_MATERIAL_TWO_SIDED) : (m.flags & ~ZYG_MATERIAL_TWO_SIDED);
                } else {  // flakes, surface / rotation / mask maps, attenuation, volumetric_anisotropy, ...
                    warnings_.push_back("material " + std::to_string(id) + ": Substitute parameter \"" + k + "\" is not supported by the device path and is ignored");
                }
            }
        } else if ("Glass" == entry.first && ZYG_MATERIAL_GLASS == m.type) {  // updateGlass, :165-196
            Vec4f attenuation_color = splat(1.f);
            for (const auto& e : v.object) {
                const std::string& k = e.first;
                if ("color" == k || "attenuation_color" == k) {
                    attenuation_color = readColor(e.second);
                } else if ("attenuation_distance" == k) {
                    m.attenuation_distance = float(e.second.number);
                } else if ("roughness" == k) {
                    m.roughness = float(e.second.number);
                } else if ("specular" == k) {
                    m.specular = float(e.second.number);
                } else if ("priority" == k) {
                    m.priority = int32_t(e.second.number);
                } else if ("ior" == k) {
                    m.ior = float(e.second.number);
                } else if ("abbe" == k) {
                    m.abbe = float(e.second.number);
                } else if ("thickness" == k) {
                    m.thickness = float(e.second.number);
                } else {
                    warnings_.push_back("material " + std::to_string(id) + ": Glass parameter \"" + k + "\" is not supported by the device path and is ignored");
                }
            }
            // Glass.setVolumetric -> attenuationCoefficient, collision_coefficients.zig:35-44
            for (int i = 0; i < 3; ++i) {
                const float c = fmin_(fmax_(attenuation_color[i], 0.01f), 0.991102f);
                m.color[i]    = 0.f == m.attenuation_distance ? 0.f : -std::log(c) / m.attenuation_distance;
            }
        }
    }

    // Material.commit: light_material.zig:46-52, substitute_material.zig:69-83
    if (ZYG_MATERIAL_LIGHT == m.type || ZYG_MATERIAL_SUBSTITUTE == m.type) {
        m.flags = anyGreaterZero3(m.emission) ? (m.flags | ZYG_MATERIAL_EMISSIVE) : (m.flags & ~ZYG_MATERIAL_EMISSIVE);
    }
    if (ZYG_MATERIAL_SUBSTITUTE == m.type) {
        m.flags = m.roughness <= specular_threshold_ ? (m.flags | ZYG_MATERIAL_CAUSTIC) : (m.flags & ~ZYG_MATERIAL_CAUSTIC);
    }
    if (ZYG_MATERIAL_GLASS == m.type) {  // glass_material.zig:38-45
        m.flags = m.roughness * m.roughness <= specular_threshold_ ? (m.flags | ZYG_MATERIAL_CAUSTIC) : (m.flags & ~ZYG_MATERIAL_CAUSTIC);
        m.flags = m.thickness > 0.f ? (m.flags | ZYG_MATERIAL_TWO_SIDED) : (m.flags & ~ZYG_MATERIAL_TWO_SIDED);
    }
    return true;
}

int SceneModel::createMaterial(const json::Value& material) {
    touch();
    const json::Value* rendering = material.get("rendering");
    if (!rendering || json::Value::Object != rendering->kind) return -1;  // Error.NoRenderNode

    for (const auto& entry : rendering->object) {
        uint32_t type;
        if ("Debug" == entry.first) {
            type = ZYG_MATERIAL_DEBUG;
        } else if ("Glass" == entry.first) {
            type = ZYG_MATERIAL_GLASS;
        } else if ("Light" == entry.first) {
            type = ZYG_MATERIAL_LIGHT;
        } else if ("Substitute" == entry.first) {
            type = ZYG_MATERIAL_SUBSTITUTE;
        } else {
            continue;
        }
        materials_.push_back(defaultMaterial(type));
        emission_maps_.push_back(EmissionMapRec{});
        color_maps_.push_back(EmissionMapRec{});
        surface_maps_.push_back({});
        updateMaterial(uint32_t(materials_.size() - 1), material);
        return int(materials_.size() - 1);
    }
    return -1;  // Error.UnknownMaterial
}

int SceneModel::createImage(uint32_t id, uint32_t format, uint32_t num_channels, uint32_t width, uint32_t height, uint32_t depth,
                            uint32_t pixel_stride, const uint8_t* data) {
    touch();
    // capi.zig:29-35 Format: UInt8 0, UInt16 1, UInt32 2, Float16 3, Float32 4
    const uint32_t bpc = 0 == format ? 1u : ((1 == format || 3 == format) ? 2u : 4u);
    // capi.zig:268-284: UInt8 x 1 / 2 / 3 and Float32 x 1 / 2 / 3 (Float32 x 4 and volumes are outside the scope)
    if (num_channels < 1 || num_channels > 3 || !(0 == format || 4 == format) || 0 == width || 0 == height || 1 != depth || !data) return -1;
    ImageRec img;
    img.width    = width;
    img.height   = height;
    img.format   = format;
    img.channels = num_channels;
    img.pixels.assign(size_t(width) * height * 3, 0.f);
    // Cache.store, resource/cache.zig: an id inside the cache replaces that entry, anything else appends
    uint32_t slot = id < images_.size() ? id : uint32_t(images_.size());
    if (slot == images_.size()) images_.push_back(ImageRec{});
    images_[slot] = std::move(img);
    if (bpc * num_channels == pixel_stride) updateImage(slot, pixel_stride, data);  // capi.zig:262-264: other strides leave the pixels unset
    return int(slot);
}

int SceneModel::updateImage(uint32_t id, uint32_t pixel_stride, const uint8_t* data) {
    touch();
    if (id >= images_.size() || !data) return -1;
    ImageRec&      img = images_[id];
    const size_t   n   = size_t(img.width) * img.height;
    const uint32_t bpp = (0 == img.format ? 1u : 4u) * img.channels;
    if (bpp != pixel_stride) return 0;  // capi.zig:322: silently ignored
    if (4 == img.format && 3 == img.channels) {
        std::memcpy(img.pixels.data(), data, n * 12);
    } else if (4 == img.format) {  // Float1 / Float2 (texture.zig:172, 183)
        const float* src = reinterpret_cast<const float*>(data);
        for (size_t i = 0; i < n; ++i) {
            for (uint32_t c = 0; c < img.channels; ++c) img.pixels[3 * i + c] = src[img.channels * i + c];
        }
    } else if (1 == img.channels) {  // Texture.Byte1_unorm: enc.cachedUnormToFloat (encoding.zig:10-12)
        for (size_t i = 0; i < n; ++i) img.pixels[3 * i] = float(data[i]) * (1.f / 255.f);
    } else if (2 == img.channels) {
        // Texture.Byte2_snorm: what createTexture makes of a Byte2 image used as a normal map (texture_provider.zig:72), the one use a
        // two-channel image has on this path; enc.snorm8ToFloat (encoding.zig:18-20)
        for (size_t i = 0; i < 2 * n; ++i) img.pixels[3 * (i / 2) + (i & 1)] = std::fmaf(float(data[i]), 1.f / 128.f, -1.f);
    } else {  // Texture.Byte3_sRGB: cachedSrgbToFloat3 then sRGB -> AP1 (texture.zig:196-199, srgb.zig:28-38)
        float table[256];
        for (int i = 0; i < 256; ++i) {
            const float c = float(i) / 255.f;
            table[i]      = c <= 0.f ? 0.f : (c < 0.04045f ? c / 12.92f : (c < 1.f ? std::pow((c + 0.055f) / 1.055f, 2.4f) : 1.f));
        }
        for (size_t i = 0; i < n; ++i) {
            const Vec4f c = sRGBtoAP1({{table[data[3 * i]], table[data[3 * i + 1]], table[data[3 * i + 2]], 0.f}});
            img.pixels[3 * i] = c[0], img.pixels[3 * i + 1] = c[1], img.pixels[3 * i + 2] = c[2];
        }
    }
    return 0;
}

namespace {

// Distribution1D.precomputePdfCdf, distribution_1d.zig:93-131. Returns the integral; `cdf` has n + 1 entries (a zero
// integral leaves the caller's degenerate row, see ZygpuImageSampler.conditional_integral).
float precomputeCdf(const float* data, size_t n, float* cdf) {
    float integral = 0.f;
    for (size_t i = 0; i < n; ++i) integral += data[i];
    if (0.f == integral) {
        for (size_t i = 0; i <= n; ++i) cdf[i] = 1.f;
        return 0.f;
    }
    const float ii = 1.f / integral;
    float       p  = 0.f;
    cdf[0]         = 0.f;
    for (size_t i = 0; i + 1 < n; ++i) {
        const float c = std::fmaf(data[i], ii, p);
        cdf[i + 1]    = c;
        p             = c;
    }
    cdf[n] = 1.f;
    return integral;
}

}  // namespace

uint32_t SceneModel::uvWeightClass(uint32_t shape) {  // Shape.uvWeight, shape.zig:535-542
    return ZYG_SHAPE_CANOPY == shape ? 1u : (ZYG_SHAPE_DOME == shape ? 2u : 0u);
}

// light_material.Material.prepareSampling, light_material.zig:54-119 + LuminanceContext / DistributionContext, :206-272
// (rows are summed in order on one thread; the reference adds per-thread partial sums)
uint32_t SceneModel::imageSampler(uint32_t material, uint32_t shape) {
    const uint32_t wc = uvWeightClass(shape);
    for (size_t i = 0; i < image_samplers_.size(); ++i) {
        if (image_samplers_[i]->material == material && image_samplers_[i]->weight_class == wc) return uint32_t(i);
    }
    const EmissionMapRec& em  = emission_maps_[material];
    const ImageRec&       img = images_[em.image];
    const uint32_t        w = img.width, h = img.height;

    auto rec          = std::make_unique<ImageSamplerRec>();
    rec->material     = material;
    rec->weight_class = wc;

    std::vector<float> luminance(size_t(w) * h);
    const float        idf[2] = {1.f / float(w), 1.f / float(h)};
    Vec4f              avg    = splat(0.f);
    for (uint32_t y = 0; y < h; ++y) {
        const float v = idf[1] * (float(y) + 0.5f);
        for (uint32_t x = 0; x < w; ++x) {
            const float u         = idf[0] * (float(x) + 0.5f);
            float       uv_weight = 1.f;
            if (1 == wc) {  // Canopy.uvWeight, canopy.zig:133-141
                const float dx = 2.f * u - 1.f, dy = 2.f * v - 1.f;
                uv_weight      = (dx * dx + dy * dy) > 1.f ? 0.f : 1.f;
            } else if (2 == wc) {
                uv_weight = std::sin(v * kPi);
            }
            const float* px = &img.pixels[3 * (size_t(y) * w + x)];
            const Vec4f  wr = {{uv_weight * px[0], uv_weight * px[1], uv_weight * px[2], 0.f}};
            avg             = avg + Vec4f{{wr[0], wr[1], wr[2], uv_weight}};
            luminance[size_t(y) * w + x] = fmax_(wr[0], fmax_(wr[1], wr[2]));
        }
    }
    const Vec4f average_emission = avg / splat(avg[3]);
    const ZygpuMaterial& m       = materials_[material];
    rec->total_weight            = avg[3];
    rec->average_emission        = Vec4f{{m.emission[0], m.emission[1], m.emission[2], m.emission[3]}} * average_emission;

    // MIS compensation (:248-272)
    const float al = 0.6f * fmax_(average_emission[0], fmax_(average_emission[1], average_emission[2]));
    rec->conditional_cdf.resize(size_t(h) * (w + 1));
    rec->conditional_integral.resize(h);
    for (uint32_t y = 0; y < h; ++y) {
        float* row = &luminance[size_t(y) * w];
        for (uint32_t x = 0; x < w; ++x) {
            const float l = row[x];
            row[x]        = fmax_(l - al, fmin_(l, 0.0025f));
        }
        rec->conditional_integral[y] = precomputeCdf(row, w, &rec->conditional_cdf[size_t(y) * (w + 1)]);
    }
    // Distribution2D.configure, distribution_2d.zig:52-61
    rec->marginal_cdf.resize(h + 1);
    precomputeCdf(rec->conditional_integral.data(), h, rec->marginal_cdf.data());

    image_samplers_.push_back(std::move(rec));
    return uint32_t(image_samplers_.size() - 1);
}

uint32_t SceneModel::addMesh(const zyg_mesh* mesh, uint32_t num_parts) {
    touch();
    meshes_.push_back({mesh, num_parts, {}, {}});
    return 7 + uint32_t(meshes_.size() - 1);
}

bool SceneModel::shapeFinite(uint32_t shape) const {  // shape.zig:94-99
    return !(ZYG_SHAPE_CANOPY == shape || ZYG_SHAPE_DISTANT == shape || ZYG_SHAPE_DOME == shape);
}

uint32_t SceneModel::createEntity() {
    touch();
    PropRec p;
    p.shape = ZYG_SHAPE_DISTANT;  // scene.zig:257
    props_.push_back(p);
    world_.push_back(Transformation{});
    return uint32_t(props_.size() - 1);
}

uint32_t SceneModel::createPropShape(uint32_t shape_id, const uint32_t* materials, uint32_t num_materials, bool unoccluding) {
    touch();
    PropRec p;
    p.shape = shape_id;

    // Prop.configureShape, prop.zig:94-135
    bool pure_emissive = true;
    bool mono          = num_materials > 0;
    for (uint32_t i = 0; i < num_materials; ++i) {
        const ZygpuMaterial& m = materials_[materials[i]];
        if (ZYG_MATERIAL_LIGHT != m.type) pure_emissive = false;
        if (materials[i] != materials[0]) mono = false;
    }
    const bool volumetric = shapeFinite(shape_id) && mono && materials_[materials[0]].ior < 1.f;
    p.solid               = !volumetric;
    if (unoccluding && shapeFinite(shape_id) && pure_emissive) p.flags |= ZYG_PROP_UNOCCLUDING;

    const uint32_t num_parts = shape_id >= 7 ? meshes_[shape_id - 7].num_parts : 1;
    p.parts_start            = uint32_t(material_ids_.size());
    for (uint32_t i = 0; i < num_parts; ++i) {
        material_ids_.push_back(num_materials > 0 ? materials[std::min(i, num_materials - 1)] : 0);
        light_ids_.push_back(ZYGPU_NULL);
    }

    props_.push_back(p);
    world_.push_back(Transformation{});
    const uint32_t id = uint32_t(props_.size() - 1);

    // Scene.classifyProp, scene.zig:322-340
    if (p.solid) {
        if (shapeFinite(shape_id)) {
            (0 != (p.flags & ZYG_PROP_UNOCCLUDING) ? unoccluding_props_ : finite_props_).push_back(id);
        } else {
            infinite_props_.push_back(id);
        }
    }
    return id;
}

int SceneModel::createPropInstance(uint32_t entity) {
    touch();
    if (entity >= props_.size()) return -1;
    const PropRec p = props_[entity];  // same shape, materials and parts; its own transformation
    props_.push_back(p);
    world_.push_back(Transformation{});
    const uint32_t id = uint32_t(props_.size() - 1);
    if (p.solid) {  // Scene.classifyProp
        if (shapeFinite(p.shape)) {
            (0 != (p.flags & ZYG_PROP_UNOCCLUDING) ? unoccluding_props_ : finite_props_).push_back(id);
        } else {
            infinite_props_.push_back(id);
        }
    }
    return int(id);
}

int SceneModel::createInstancer(const uint32_t* prototypes, uint32_t num_prototypes, const uint32_t* prototype_indices,
                                const Transformation* transformations, uint32_t num_instances) {
    touch();
    if (0 == num_prototypes || !prototypes || (num_instances > 0 && (!prototype_indices || !transformations))) return -1;
    for (uint32_t i = 0; i < num_prototypes; ++i) {
        if (prototypes[i] >= props_.size() || ZYGPU_NULL == props_[prototypes[i]].parts_start) return -1;
    }
    // prototypes are loaded with is_prototype = true: they are not classified into the scene's own trees (scene.zig:299-301)
    auto erase = [](std::vector<uint32_t>& v, uint32_t id) { v.erase(std::remove(v.begin(), v.end(), id), v.end()); };
    for (uint32_t i = 0; i < num_prototypes; ++i) {
        erase(finite_props_, prototypes[i]);
        erase(unoccluding_props_, prototypes[i]);
        erase(infinite_props_, prototypes[i]);
    }
    InstancerRec rec;
    rec.entity = createEntity();
    rec.prototypes.resize(num_instances);
    rec.trafos.assign(transformations, transformations + num_instances);
    for (uint32_t i = 0; i < num_instances; ++i) {
        const uint32_t pi = prototype_indices[i] < num_prototypes ? prototype_indices[i] : 0;  // scene_loader.zig:453-456
        rec.prototypes[i] = prototypes[pi];
    }
    instancers_.push_back(std::move(rec));
    return int(instancers_.back().entity);
}

bool SceneModel::createLight(uint32_t entity) {  // scene.zig:342-372
    touch();
    if (entity >= props_.size()) return false;
    const PropRec& p         = props_[entity];
    const uint32_t num_parts = p.shape >= 7 ? meshes_[p.shape - 7].num_parts : 1;
    for (uint32_t i = 0; i < num_parts; ++i) {
        const ZygpuMaterial& m = materials_[material_ids_[p.parts_start + i]];
        if (0 == (m.flags & ZYG_MATERIAL_EMISSIVE)) continue;
        ZygpuLight l{};
        l.prop        = entity;
        l.part        = i;
        l.light_class = ZYG_LIGHT_PROP;  // becomes PROP_IMAGE in compile when the material has an emission image (scene.zig:356-366)
        l.two_sided   = 0 != (m.flags & ZYG_MATERIAL_TWO_SIDED);
        l.num_samples = m.emission_num_samples;
        lights_.push_back(l);
    }
    return true;
}

bool SceneModel::setWorldTransformation(uint32_t entity, const Transformation& t) {
    touch();
    if (entity >= props_.size()) return false;
    world_[entity] = t;
    return true;
}

bool SceneModel::setVisibility(uint32_t entity, bool in_camera, bool in_reflection, bool /*in_sss*/) {  // prop.zig:78-91
    touch();
    if (entity >= props_.size()) return false;
    uint32_t& f = props_[entity].flags;
    f &= ~(ZYG_PROP_VISIBLE_IN_CAMERA | ZYG_PROP_VISIBLE_IN_REFLECTION | ZYG_PROP_VISIBLE_IN_SHADOW);
    if (in_camera) f |= ZYG_PROP_VISIBLE_IN_CAMERA;
    if (in_reflection) f |= ZYG_PROP_VISIBLE_IN_REFLECTION | ZYG_PROP_VISIBLE_IN_SHADOW;
    return true;
}

void SceneModel::setCamera(uint32_t width, uint32_t height) {  // capi.zig:143-167
    touch();
    resolution_[0] = int32_t(width);
    resolution_[1] = int32_t(height);
    fov_           = degreesToRadians(80.f);
    if (ZYGPU_NULL == camera_entity_) camera_entity_ = createEntity();
}

void SceneModel::loadIntegrators(const json::Value& value) {
    touch();
    if (const json::Value* st = value.get("specular_threshold")) {
        const float s       = float(st->number);
        specular_threshold_ = s * s;
    }
    const json::Value* surface = value.get("surface");
    if (!surface || json::Value::Object != surface->kind) return;
    for (const auto& entry : surface->object) {
        if ("PTMIS" != entry.first) continue;  // only PathtracerMIS is implemented
        const json::Value& v   = entry.second;
        regularize_roughness_  = json::readFloatMember(v, "regularize_roughness", 0.f);
        caustics_path_         = json::readBoolMember(v, "caustics", true);
        max_depth_surface_     = 16;  // take.zig:77
        max_depth_volume_      = 256;
        if (const json::Value* d = v.get("depth")) {  // loadDepth reads "surface" for both, take.zig:254-261
            max_depth_surface_ = json::readUIntMember(*d, "surface", 16) & 0xFFFFu;
            max_depth_volume_  = json::readUIntMember(*d, "surface", 256) & 0xFFFFu;
        }
        // loadLightSampling, take.zig:263-271: without a "light_sampling" node the threshold is the raw default 0.5; only a
        // value read from the node is clamped and raised to the 4th power
        split_threshold_ = 0.5f;
        if (const json::Value* ls = v.get("light_sampling")) {
            float st = json::readFloatMember(*ls, "split_threshold", 0.5f);
            st       = st < 0.f ? 0.f : (st > 1.f ? 1.f : st);
            const float st2  = st * st;
            split_threshold_ = st2 * st2;
        }
        ptmis_           = true;
    }
}

void SceneModel::loadSensor(const json::Value& value) {
    touch();
    clamp_[0] = clamp_[1] = clamp_[2] = FLT_MAX;
    if (const json::Value* c = value.get("clamp")) {
        if (json::Value::Object == c->kind) {
            clamp_[0] = json::readFloatMember(*c, "emission", clamp_[0]);
            clamp_[1] = json::readFloatMember(*c, "direct", clamp_[1]);
            clamp_[2] = json::readFloatMember(*c, "indirect", clamp_[2]);
        }
    }
    alpha_transparency_ = json::readBoolMember(value, "alpha_transparency", false);  // take_loader.zig:194-196
    filter_kind_        = FilterKind::None;
    filter_radius_      = 0.f;
    if (const json::Value* f = value.get("filter")) {
        for (const auto& entry : f->object) {
            if ("Blackman" == entry.first) {
                filter_kind_   = FilterKind::Blackman;
                filter_radius_ = 2.f;
                break;
            }
            if ("Mitchell" == entry.first) {
                filter_kind_   = FilterKind::Mitchell;
                filter_radius_ = 2.f;
                break;
            }
        }
    }
}

// View.loadAOV, take.zig:106-129: {"Albedo": true, "Depth": true, ...} sets or clears the class bits; unknown keys are ignored
void SceneModel::loadAovs(const json::Value& value) {
    touch();
    static const char* const kNames[ZYG_AOV_NUM_CLASSES] = {"Albedo",    "Depth",    "MaterialId", "GeometricNormal", "ShadingNormal",
                                                            "Roughness", "Emission", "Direct",     "Indirect"};
    if (json::Value::Object != value.kind) return;
    for (const auto& entry : value.object) {
        for (uint32_t c = 0; c < ZYG_AOV_NUM_CLASSES; ++c) {
            if (entry.first != kNames[c]) continue;
            const bool on = json::Value::Bool == entry.second.kind ? entry.second.boolean : false;  // json.readBool
            aov_slots_    = on ? (aov_slots_ | (1u << c)) : (aov_slots_ & ~(1u << c));
        }
    }
}

void SceneModel::loadSampler(const json::Value& value) {
    touch();
    sampler_ = ZYG_SAMPLER_SOBOL;
    for (const auto& entry : value.object) {
        spp_ = json::readUIntMember(entry.second, "samples_per_pixel", 1);
        if ("Random" == entry.first) {
            sampler_ = ZYG_SAMPLER_RANDOM;
            return;
        }
    }
}

AABB SceneModel::shapeAabb(uint32_t shape) const {  // shape.zig:108-115
    switch (shape) {
        case ZYG_SHAPE_CANOPY:
        case ZYG_SHAPE_DISTANT:
        case ZYG_SHAPE_DOME: return AABB::empty();
        case ZYG_SHAPE_DISK:
        case ZYG_SHAPE_RECTANGLE: return {{{{-0.5f, -0.5f, 0.f, 0.f}}, {{0.5f, 0.5f, 0.f, 0.f}}}};
        case ZYG_SHAPE_CUBE:
        case ZYG_SHAPE_SPHERE: return {{splat(-0.5f), splat(0.5f)}};
        default: return meshes_[shape - 7].mesh->tree.aabb();
    }
}

// PropBvhBuilder.build + serialize, prop_tree_builder.zig:24-96
void SceneModel::buildPropTree(const std::vector<uint32_t>& indices, std::vector<ZygpuBvhNode>& nodes,
                               std::vector<uint32_t>& out_indices) {
    nodes.clear();
    out_indices.clear();
    if (indices.empty()) return;

    std::vector<Reference> references(indices.size());
    AABB                   bounds = AABB::empty();
    for (size_t i = 0; i < indices.size(); ++i) {
        const ZygpuAabb& b  = flat_aabbs_[indices[i]];
        const Vec4f      mi = {{b.min[0], b.min[1], b.min[2], b.min[3]}};
        const Vec4f      ma = {{b.max[0], b.max[1], b.max[2], b.max[3]}};
        references[i].set(mi, ma, indices[i]);
        bounds.mergeAssign({{mi, ma}});
    }

    BuildResult build;
    buildBinaryBvh(std::move(references), bounds, 16, 64, 4, 0, build);

    nodes.assign(build.build_nodes.size(), ZygpuBvhNode{});
    out_indices.assign(build.reference_ids.size(), 0);

    struct Item {
        uint32_t src, dst;
    };
    uint32_t          current_node = 1, current_prop = 0;
    std::vector<Item> stack{{0, 0}};
    while (!stack.empty()) {
        const Item it = stack.back();
        stack.pop_back();
        const BvhNode& node = build.build_nodes[it.src];
        ZygpuBvhNode   n;
        for (int k = 0; k < 3; ++k) {
            n.min[k] = node.min[k];
            n.max[k] = node.max[k];
        }
        if (0 == node.numIndices()) {
            const uint32_t child0 = current_node;
            n.children_or_start   = child0;
            n.num_indices         = 0;
            current_node += 2;
            stack.push_back({node.children() + 1, child0 + 1});
            stack.push_back({node.children(), child0});
        } else {
            const uint32_t num  = node.numIndices();
            n.children_or_start = current_prop;
            n.num_indices       = num;
            for (uint32_t k = 0; k < num; ++k) out_indices[current_prop + k] = build.reference_ids[node.children() + k];
            current_prop += num;
        }
        nodes[it.dst] = n;
    }
}

// LightTreeBuilder.build, light_tree_builder.zig:281-376
bool SceneModel::buildLightTree(std::string& /*error*/) {
    const uint32_t num_lights = uint32_t(lights_.size());
    light_mapping_.assign(num_lights, 0);
    light_orders_.assign(num_lights, 0);
    light_nodes_.clear();
    light_node_middles_.clear();

    uint32_t lm = 0, order = 0;
    for (uint32_t l = 0; l < num_lights; ++l) {
        if (!shapeFinite(props_[lights_[l].prop].shape)) light_mapping_[lm++] = l;
    }
    const uint32_t num_infinite = lm;
    for (uint32_t l = 0; l < num_lights; ++l) {
        if (shapeFinite(props_[lights_[l].prop].shape)) light_mapping_[lm++] = l;
    }
    float infinite_total_power = 0.f;
    for (uint32_t i = 0; i < num_infinite; ++i) {
        light_orders_[light_mapping_[i]] = order++;
        infinite_total_power += light_aabbs_[light_mapping_[i]].min[3];
    }
    {  // Tree.infinite_light_distribution.configure(infinite_light_powers), light_tree_builder.zig:312-326
        std::vector<float> powers(num_infinite);
        for (uint32_t i = 0; i < num_infinite; ++i) powers[i] = light_aabbs_[light_mapping_[i]].min[3];
        infinite_cdf_.assign(size_t(num_infinite) + 1, 1.f);
        if (num_infinite > 0) precomputeCdf(powers.data(), num_infinite, infinite_cdf_.data());
    }

    ZygpuLightTree& t     = flat_.light_tree;
    t                     = ZygpuLightTree{};
    t.infinite_end        = order;
    t.num_lights          = num_lights;
    t.num_infinite_lights = num_infinite;

    // what the builder reads of a light: Scene.lightAabb / lightCone / lightPower / lightTwoSided, scene.zig:650-664
    std::vector<AABB>    aabbs(num_lights);
    std::vector<Vec4f>   cones(num_lights);
    std::vector<float>   powers(num_lights);
    std::vector<uint8_t> two_sided(num_lights);
    for (uint32_t l = 0; l < num_lights; ++l) {
        const ZygpuAabb& b = light_aabbs_[l];
        aabbs[l]           = {{{{b.min[0], b.min[1], b.min[2], b.min[3]}}, {{b.max[0], b.max[1], b.max[2], b.max[3]}}}};
        cones[l]           = {{light_cones_[l * 4], light_cones_[l * 4 + 1], light_cones_[l * 4 + 2], light_cones_[l * 4 + 3]}};
        powers[l]          = b.min[3];
        two_sided[l]       = lights_[l].two_sided ? 1 : 0;
    }
    const LightSet set{aabbs.data(), cones.data(), powers.data(), two_sided.data(), false, false};

    LightTreeResult tree;
    zyg::buildLightTree(set, light_mapping_, num_infinite, order, light_orders_, tree);
    light_nodes_        = std::move(tree.nodes);
    light_node_middles_ = std::move(tree.node_middles);
    t.max_split_depth   = tree.max_split_depth;
    t.num_nodes         = uint32_t(light_nodes_.size());
    if (t.num_nodes > 0) {
        for (int k = 0; k < 4; ++k) {
            t.bounds.min[k] = tree.bounds.b[0][k];
            t.bounds.max[k] = tree.bounds.b[1][k];
        }
    }

    const uint32_t num_finite = num_lights - num_infinite;
    const float    p0         = infinite_total_power;
    const float    p1         = 0 == num_finite ? 0.f : tree.root_power;
    const float    pt         = p0 + p1;
    t.infinite_weight = (0 == num_lights || 0.f == pt) ? 0.f : p0 / pt;
    t.infinite_guard  = 0 == num_finite ? (0 == num_infinite ? 0.f : 1.1f) : t.infinite_weight;

    t.nodes         = light_nodes_.data();
    t.node_middles  = light_node_middles_.data();
    t.light_orders  = light_orders_.data();
    t.light_mapping = light_mapping_.data();
    t.infinite_cdf  = infinite_cdf_.data();
    return true;
}

bool SceneModel::compile(std::string& error) {
    if (ZYGPU_NULL == camera_entity_) {
        error = "no camera";  // Error.NoCameraProp, driver.zig:122-124
        return false;
    }
    const std::vector<float>& luts = ggxLuts(error);
    if (luts.empty()) return false;

    // The device intersects Rectangle, Cube, Disk, Sphere, triangle meshes, Distant and Canopy. A Dome prop would be silently invisible and
    // a Disk light with an emission map (Disk.sampleMaterialTo, disk.zig:334-412) would be sampled uniformly: refuse the scene instead,
    // like thin or dispersive Glass at upload.
    for (const PropRec& p : props_) {
        if (ZYGPU_NULL != p.shape && p.shape >= 7 && !meshes_[p.shape - 7].mesh) {
            error = "shape " + std::to_string(p.shape) + " has no triangle tree (its build failed)";
            return false;
        }
    }
    for (const std::vector<uint32_t>* list : {&finite_props_, &unoccluding_props_, &infinite_props_}) {
        for (uint32_t id : *list) {
            if (ZYG_SHAPE_DOME == props_[id].shape) {
                error = "prop " + std::to_string(id) + ": the shape Dome is not supported by the device path";
                return false;
            }
        }
    }
    for (const InstancerRec& ir : instancers_) {
        for (uint32_t proto : ir.prototypes) {
            if (ZYG_SHAPE_DOME == props_[proto].shape) {
                error = "prop " + std::to_string(proto) + ": the shape Dome is not supported by the device path";
                return false;
            }
        }
    }

    uint32_t num_props = uint32_t(props_.size());
    for (const InstancerRec& ir : instancers_) num_props += uint32_t(ir.prototypes.size());
    const Vec4f origin = world_[camera_entity_].position;  // Scene.propWorldPosition, driver.zig:163

    flat_props_.resize(num_props);
    flat_trafos_.resize(num_props);
    flat_aabbs_.resize(num_props);

    // one flat prop from a composed transformation: rotation rows with the scale in lane 3, position in world space
    auto emitProp = [&](uint32_t i, const PropRec& p, Mat3x3 rot, Vec4f position) {
        // Space.calculateWorldBounds, space.zig:60-97 (static prop)
        const Vec4f scale  = {{rot.r[0][3], rot.r[1][3], rot.r[2][3], 1.f}};
        AABB        bounds = transformAabb(shapeAabb(p.shape), compose(rot, scale, position));
        bounds.translate(-origin);
        bounds.cacheRadius();
        flat_aabbs_[i] = packAabb(bounds);

        // Space.transformationAtMaybeStatic, space.zig:103-111
        ZygpuTrafo& ft = flat_trafos_[i];
        for (int r = 0; r < 3; ++r) {
            for (int c = 0; c < 4; ++c) ft.r[r][c] = rot.r[r][c];
        }
        const Vec4f pos = position + (-origin);
        for (int c = 0; c < 4; ++c) ft.position[c] = pos[c];

        ZygpuProp& fp  = flat_props_[i];
        fp.shape       = p.shape >= 7 ? uint32_t(ZYG_SHAPE_TRIANGLE_MESH) : p.shape;
        fp.mesh        = p.shape >= 7 ? p.shape - 7 : ZYGPU_NULL;
        fp.flags       = p.flags;
        fp.parts_start = p.parts_start;
    };
    auto composed = [](const Transformation& t) {  // ComposedTransformation.init, composed_transformation.zig:19-31
        Mat3x3 rot  = quaternionToMat3x3(t.rotation);
        rot.r[0][3] = t.scale[0];
        rot.r[1][3] = t.scale[1];
        rot.r[2][3] = t.scale[2];
        return rot;
    };

    const uint32_t num_entities = uint32_t(props_.size());
    for (uint32_t i = 0; i < num_entities; ++i) emitProp(i, props_[i], composed(world_[i]), world_[i].position);

    // instancers: instance trafo = instancer.transform(instance) (instancer.zig:77-86, composed_transformation.zig:55-68)
    flat_finite_      = finite_props_;
    flat_unoccluding_ = unoccluding_props_;
    uint32_t next     = num_entities;
    for (const InstancerRec& ir : instancers_) {
        const Transformation& st    = world_[ir.entity];
        const Mat3x3          self  = composed(st);
        const Vec4f           sscale = {{st.scale[0], st.scale[1], st.scale[2], 1.f}};
        const Vec4f           a = self.r[0] * splat(sscale[0]), b = self.r[1] * splat(sscale[1]), c = self.r[2] * splat(sscale[2]);
        for (size_t k = 0; k < ir.prototypes.size(); ++k) {
            const Transformation& ot    = ir.trafos[k];
            const Mat3x3          other = composed(ot);
            Mat3x3                rot   = mulMat(other, self);
            rot.r[0][3]                 = st.scale[0] * ot.scale[0];
            rot.r[1][3]                 = st.scale[1] * ot.scale[1];
            rot.r[2][3]                 = st.scale[2] * ot.scale[2];
            // objectToWorldPoint(other.position), composed_transformation.zig:70-94
            Vec4f pos = splat(ot.position[0]) * a;
            pos       = mulAdd(splat(ot.position[1]), b, pos);
            pos       = mulAdd(splat(ot.position[2]), c, pos);
            pos       = pos + st.position;
            pos[3]    = 0.f;

            const PropRec& proto = props_[ir.prototypes[k]];
            emitProp(next, proto, rot, pos);
            if (proto.solid && shapeFinite(proto.shape)) {
                (0 != (proto.flags & ZYG_PROP_UNOCCLUDING) ? flat_unoccluding_ : flat_finite_).push_back(next);
            }
            next += 1;
        }
    }

    buildPropTree(flat_finite_, solid_nodes_, solid_indices_);
    buildPropTree(flat_unoccluding_, unocc_nodes_, unocc_indices_);

    // Scene.propPrepareSampling, scene.zig:402-497 (static props, analytic shapes)
    const uint32_t num_lights = uint32_t(lights_.size());
    light_aabbs_.resize(num_lights);
    light_cones_.resize(size_t(num_lights) * 4);
    flat_part_areas_.assign(material_ids_.size(), 0.f);
    image_samplers_.clear();
    for (size_t m = 0; m < materials_.size(); ++m) {
        const EmissionMapRec& em = emission_maps_[m];
        materials_[m].emission_map = ZYGPU_NULL;
        materials_[m].color_map    = ZYGPU_NULL;
        materials_[m].roughness_map = materials_[m].metallic_map = materials_[m].normal_map = ZYGPU_NULL;
        if (ZYG_MATERIAL_SUBSTITUTE == materials_[m].type && materials_[m].coating_thickness > 0.f && 0 != (materials_[m].flags & ZYG_MATERIAL_EMISSIVE)) {
            // substitute_material.zig:274-292 attenuates the emission by the coat: not on the device path, refused instead of rendered wrongly
            error = "material " + std::to_string(m) + ": an emissive Substitute with a coating is not supported";
            return false;
        }
        static const char* const kMapName[kNumSurfaceMaps] = {"roughness", "metallic", "normal"};
        for (int k = 0; k < kNumSurfaceMaps; ++k) {
            const uint32_t image = surface_maps_[m][k].image;
            if (ZYGPU_NULL == image) continue;
            if (image >= images_.size()) {
                error = "material " + std::to_string(m) + ": " + kMapName[k] + " references image " + std::to_string(image) + " which does not exist";
                return false;
            }
            // texture.zig:166-186: a fetch with the wrong channel count returns zero; refuse instead of rendering something else
            const uint32_t want = kNormalMap == k ? 2u : 1u;
            if (images_[image].channels != want) {
                error = "material " + std::to_string(m) + ": " + kMapName[k] + " map needs an image of " + std::to_string(want) + " channel(s)";
                return false;
            }
        }
        if (ZYGPU_NULL != color_maps_[m].image && color_maps_[m].image >= images_.size()) {
            error = "material " + std::to_string(m) + ": color references image " + std::to_string(color_maps_[m].image) + " which does not exist";
            return false;
        }
        if (ZYGPU_NULL == em.image) continue;
        if (em.image >= images_.size()) {
            error = "material " + std::to_string(m) + ": emission_map references image " + std::to_string(em.image) + " which does not exist";
            return false;
        }
    }
    for (uint32_t l = 0; l < num_lights; ++l) {
        ZygpuLight&           light = lights_[l];
        const PropRec&        p     = props_[light.prop];
        const Transformation& t     = world_[light.prop];

        light_ids_[p.parts_start + light.part] = l;

        // Scene.createLight, scene.zig:356-366: an analytic shape with an emission image is sampled through the image
        const uint32_t        light_material = material_ids_[p.parts_start + light.part];
        const ImageSamplerRec* image_sampler = nullptr;
        light.light_class                    = ZYG_LIGHT_PROP;
        if (ZYG_SHAPE_DISK == p.shape && ZYGPU_NULL != emission_maps_[light_material].image) {
            error = "prop " + std::to_string(light.prop) + ": a Disk light with an emission map is not supported by the device path";
            return false;
        }
        if (p.shape < 7 && ZYGPU_NULL != emission_maps_[light_material].image) {
            light.light_class = ZYG_LIGHT_PROP_IMAGE;
            light.sampler     = imageSampler(light_material, p.shape);
            image_sampler     = image_samplers_[light.sampler].get();
        }

        // ShapeSamplerCache.prepareSampling -> Mesh.prepareSampling -> Part.configure (triangle_mesh.zig:57-149, 722-746)
        if (!image_sampler) light.sampler = ZYGPU_NULL;
        const MeshSamplerData* mesh_sampler = nullptr;
        if (p.shape >= 7) {
            MeshRec& mr = meshes_[p.shape - 7];
            if (mr.primitive_mapping.empty()) meshPartTables(mr.mesh->tree, mr.primitive_mapping, mr.part_areas);
            for (size_t k = 0; k < samplers_.size(); ++k) {
                if (samplers_[k]->mesh == p.shape - 7 && samplers_[k]->part == light.part && samplers_[k]->two_sided == (0 != light.two_sided)) {
                    light.sampler = uint32_t(k);
                }
            }
            if (ZYGPU_NULL == light.sampler) {
                samplers_.push_back(std::make_unique<SamplerRec>());
                SamplerRec& rec = *samplers_.back();
                rec.mesh        = p.shape - 7;
                rec.part        = light.part;
                rec.two_sided   = 0 != light.two_sided;
                buildMeshSampler(mr.mesh->tree, light.part, rec.two_sided, rec.data);
                light.sampler = uint32_t(samplers_.size() - 1);
            }
            mesh_sampler = &samplers_[light.sampler]->data;
            for (uint32_t k = 0; k < mr.num_parts; ++k) flat_part_areas_[p.parts_start + k] = mr.part_areas[k];
        }

        Mat3x3 rot  = quaternionToMat3x3(t.rotation);
        rot.r[0][3] = t.scale[0];
        rot.r[1][3] = t.scale[1];
        rot.r[2][3] = t.scale[2];
        const Vec4f scale = {{t.scale[0], t.scale[1], t.scale[2], 1.f}};
        const Vec4f pos   = t.position + (-origin);

        // sampler.impl.aabb / cone: the emitting triangles' for a mesh part, the shape's otherwise (shape_sampler.zig:105-117)
        AABB bb = transformAabb(mesh_sampler ? mesh_sampler->aabb : shapeAabb(p.shape), compose(rot, scale, pos));
        bb.cacheRadius();

        // shape.cone(), shape.zig:117-122
        const bool  flat_shape = ZYG_SHAPE_DISK == p.shape || ZYG_SHAPE_RECTANGLE == p.shape || ZYG_SHAPE_DISTANT == p.shape;
        const Vec4f part_cone  = mesh_sampler ? mesh_sampler->cone : Vec4f{{0.f, 0.f, 1.f, flat_shape ? 1.f : -1.f}};
        const Vec4f tc         = transformVector(rot, part_cone);
        light_cones_[l * 4 + 0] = tc[0];
        light_cones_[l * 4 + 1] = tc[1];
        light_cones_[l * 4 + 2] = tc[2];
        light_cones_[l * 4 + 3] = part_cone[3];

        // shape.area, shape.zig:143-156
        float extent = 0.f;
        switch (p.shape) {
            case ZYG_SHAPE_RECTANGLE: extent = scale[0] * scale[1]; break;
            case ZYG_SHAPE_SPHERE: extent = (4.f * kPi) * ((0.5f * scale[0]) * (0.5f * scale[0])); break;
            case ZYG_SHAPE_CUBE: extent = 2.f * (scale[0] * scale[1] + scale[0] * scale[2] + scale[1] * scale[2]); break;
            case ZYG_SHAPE_DISK: extent = kPi * ((0.5f * scale[0]) * (0.5f * scale[0])); break;
            case ZYG_SHAPE_DISTANT:  // Distant.solidAngle, distant.zig:143-145
                extent = (2.f * kPi) * (1.f - std::sqrt(1.f / (scale[0] * scale[0] + 1.f)));
                break;
            case ZYG_SHAPE_CANOPY: extent = 2.f * kPi; break;
            default:  // Mesh.area, triangle_mesh.zig:283-286
                if (p.shape >= 7) extent = meshes_[p.shape - 7].part_areas[light.part] * (scale[0] * scale[1]);
                break;
        }

        // Emittance.totalEmission (emittance.zig:61-71) of the average radiance (= emittance value for uniform
        // emission), then Light.power (light.zig:65-75; finite lights)
        const ZygpuMaterial& m     = materials_[material_ids_[p.parts_start + light.part]];
        Vec4f                power = {{m.emission[0], m.emission[1], m.emission[2], 0.f}};
        if (image_sampler) power = image_sampler->average_emission;  // Sampler.averageEmission, shape_sampler.zig:43-48
        if (extent <= 0.f) {
            power = splat(0.f);
        } else if (0.f == m.emission_normalize) {
            power = power * splat(extent);
        }
        if (!shapeFinite(p.shape) && !solid_nodes_.empty()) {
            // Light.power, light.zig:65-75: an infinite light's power scales with the squared extent of the scene box
            // (Scene.aabb = the root of the solid prop tree, scene.zig:173-175)
            const ZygpuBvhNode& root   = solid_nodes_[0];
            const Vec4f         box_extent = {{root.max[0] - root.min[0], root.max[1] - root.min[1], root.max[2] - root.min[2], 0.f}};
            power                          = splat(dot3(box_extent, box_extent)) * power;
        }
        bb.b[0][3] = fmax_(power[0], fmax_(power[1], power[2]));  // hmax3
        light_aabbs_[l] = packAabb(bb);
    }

    // every material with an emission image gets (at least) the sampler of the plain uv weight for its texture lookups
    for (size_t m = 0; m < materials_.size(); ++m) {
        if (ZYGPU_NULL == emission_maps_[m].image) continue;
        uint32_t first = ZYGPU_NULL;
        for (size_t i = 0; i < image_samplers_.size() && ZYGPU_NULL == first; ++i) {
            if (image_samplers_[i]->material == m) first = uint32_t(i);
        }
        materials_[m].emission_map = ZYGPU_NULL != first ? first : imageSampler(uint32_t(m), ZYG_SHAPE_RECTANGLE);
    }
    flat_image_samplers_.clear();
    for (const auto& rec : image_samplers_) {
        const EmissionMapRec& em  = emission_maps_[rec->material];
        const ImageRec&       img = images_[em.image];
        ZygpuImageSampler     is{};
        is.width                = img.width;
        is.height               = img.height;
        is.address_u            = em.address_u;
        is.address_v            = em.address_v;
        is.filter               = em.filter;
        is.total_weight         = rec->total_weight;
        is.scale[0]             = em.scale[0];
        is.scale[1]             = em.scale[1];
        is.pixels               = img.pixels.data();
        is.marginal_cdf         = rec->marginal_cdf.data();
        is.conditional_cdf      = rec->conditional_cdf.data();
        is.conditional_integral = rec->conditional_integral.data();
        flat_image_samplers_.push_back(is);
    }
    for (size_t m = 0; m < materials_.size(); ++m) {  // colour maps: looked up only (no distribution)
        const EmissionMapRec& cm = color_maps_[m];
        if (ZYGPU_NULL == cm.image) continue;
        const ImageRec&   img = images_[cm.image];
        ZygpuImageSampler is{};
        is.width     = img.width;
        is.height    = img.height;
        is.address_u = cm.address_u;
        is.address_v = cm.address_v;
        is.filter    = cm.filter;
        is.scale[0]  = cm.scale[0];
        is.scale[1]  = cm.scale[1];
        is.pixels    = img.pixels.data();
        materials_[m].color_map = uint32_t(flat_image_samplers_.size());
        flat_image_samplers_.push_back(is);
    }
    for (size_t m = 0; m < materials_.size(); ++m) {  // roughness / metallic / normal maps: looked up only
        for (int k = 0; k < kNumSurfaceMaps; ++k) {
            const EmissionMapRec& sm = surface_maps_[m][k];
            if (ZYGPU_NULL == sm.image) continue;
            const ImageRec&   img = images_[sm.image];
            ZygpuImageSampler is{};
            is.width     = img.width;
            is.height    = img.height;
            is.address_u = sm.address_u;
            is.address_v = sm.address_v;
            is.filter    = sm.filter;
            is.scale[0]  = sm.scale[0];
            is.scale[1]  = sm.scale[1];
            is.pixels    = img.pixels.data();
            uint32_t& slot = kRoughnessMap == k ? materials_[m].roughness_map : (kMetallicMap == k ? materials_[m].metallic_map : materials_[m].normal_map);
            slot           = uint32_t(flat_image_samplers_.size());
            flat_image_samplers_.push_back(is);
        }
    }

    flat_ = ZygpuScene{};
    if (!buildLightTree(error)) return false;

    flat_meshes_.clear();
    for (const MeshRec& m : meshes_) flat_meshes_.push_back(m.mesh);

    flat_samplers_.clear();
    for (const auto& rec : samplers_) {
        const MeshSamplerData& d = rec->data;
        ZygpuMeshSampler       ms{};
        for (int k = 0; k < 4; ++k) {
            ms.bounds.min[k] = d.tree.bounds.b[0][k];
            ms.bounds.max[k] = d.tree.bounds.b[1][k];
        }
        ms.num_triangles     = uint32_t(d.triangle_mapping.size());
        ms.num_nodes         = uint32_t(d.tree.nodes.size());
        ms.two_sided         = d.two_sided ? 1 : 0;
        ms.mesh              = rec->mesh;
        ms.nodes             = d.tree.nodes.data();
        ms.node_middles      = d.tree.node_middles.data();
        ms.light_orders      = d.tree.light_orders.data();
        ms.light_mapping     = d.tree.light_mapping.data();
        ms.triangle_mapping  = d.triangle_mapping.data();
        ms.triangle_pdfs     = d.triangle_pdfs.data();
        ms.primitive_mapping = meshes_[rec->mesh].primitive_mapping.data();
        flat_samplers_.push_back(ms);
    }

    flat_.num_props          = num_props;
    flat_.num_parts          = uint32_t(material_ids_.size());
    flat_.num_materials      = uint32_t(materials_.size());
    flat_.num_lights         = num_lights;
    flat_.num_infinite_props = uint32_t(infinite_props_.size());
    flat_.num_meshes         = uint32_t(flat_meshes_.size());
    flat_.props              = flat_props_.data();
    flat_.trafos             = flat_trafos_.data();
    flat_.aabbs              = flat_aabbs_.data();
    flat_.material_ids       = material_ids_.data();
    flat_.light_ids          = light_ids_.data();
    flat_.materials          = materials_.data();
    flat_.lights             = lights_.data();
    flat_.light_aabbs        = light_aabbs_.data();
    flat_.light_cones        = light_cones_.data();
    flat_.solid_bvh          = {uint32_t(solid_nodes_.size()), uint32_t(solid_indices_.size()), solid_nodes_.data(), solid_indices_.data()};
    flat_.unoccluding_bvh    = {uint32_t(unocc_nodes_.size()), uint32_t(unocc_indices_.size()), unocc_nodes_.data(), unocc_indices_.data()};
    flat_.infinite_props     = infinite_props_.data();
    flat_.num_mesh_samplers  = uint32_t(flat_samplers_.size());
    flat_.mesh_samplers      = flat_samplers_.data();
    flat_.mesh_part_areas    = flat_part_areas_.data();
    flat_.num_image_samplers = uint32_t(flat_image_samplers_.size());
    flat_.image_samplers     = flat_image_samplers_.data();
    flat_.meshes             = flat_meshes_.data();
    flat_.ggx_luts           = luts.data();

    // ---- view ----
    ZygpuView& v    = view_;
    v               = ZygpuView{};
    v.resolution[0] = resolution_[0];
    v.resolution[1] = resolution_[1];
    {  // Base.setResolution, camera_base.zig:32-41
        int32_t cc[4] = {0, 0, resolution_[0], resolution_[1]};
        if (crop_[2] >= 0) {
            for (int k = 0; k < 4; ++k) cc[k] = std::max(crop_[k], 0);
            cc[2] = std::min(cc[2], resolution_[0]);
            cc[3] = std::min(cc[3], resolution_[1]);
            cc[0] = std::min(cc[0], cc[2]);
            cc[1] = std::min(cc[1], cc[3]);
        }
        for (int k = 0; k < 4; ++k) v.crop[k] = cc[k];
    }

    {  // Perspective.update, camera_perspective.zig:79-122 (mono)
        const float fr0   = float(resolution_[0]);
        const float fr1   = float(resolution_[1]);
        const float ratio = fr1 / fr0;
        const float z     = 1.f / std::tan(0.5f * fov_);

        const Vec4f left_top    = {{-1.f, ratio, z, 0.f}};
        const Vec4f right_top   = {{1.f, ratio, z, 0.f}};
        const Vec4f left_bottom = {{-1.f, -ratio, z, 0.f}};
        const Vec4f d_x         = (right_top - left_top) / splat(fr0);
        const Vec4f d_y         = (left_bottom - left_top) / splat(fr1);
        for (int c = 0; c < 4; ++c) {
            v.left_top[c]   = left_top[c];
            v.d_x[c]        = d_x[c];
            v.d_y[c]        = d_y[c];
            v.eye_offset[c] = 0.f;
        }
    }
    v.camera_trafo    = flat_trafos_[camera_entity_];
    v.aperture_radius = aperture_radius_;
    v.focus_distance  = focus_distance_;

    v.sampler   = sampler_;
    v.spp_total = spp_;

    v.max_depth_surface    = max_depth_surface_;
    v.max_depth_volume     = max_depth_volume_;
    v.split_threshold      = split_threshold_;
    v.regularize_roughness = regularize_roughness_;
    v.caustics_path        = caustics_path_ ? 1u : 0u;
    v.specular_threshold   = specular_threshold_;

    v.clamp_emission = clamp_[0];
    v.clamp_direct   = clamp_[1];
    v.clamp_indirect = clamp_[2];

    {  // Sensor.init, sensor.zig:106-124 + InterpolatedFunction1DN.init, interpolated_function.zig:100-120
        const float radius      = filter_radius_;
        v.filter_radius_int     = int32_t(std::ceil(radius));
        const float interval    = (radius - 0.f) / float(30 - 1);
        v.filter_range_end      = radius;
        v.filter_inverse_interval = 1.f / interval;

        auto evalFilter = [&](float x) -> float {
            if (FilterKind::Mitchell == filter_kind_) {  // sensor.zig:42-58
                const float b = 1.f / 3.f, c = 1.f / 3.f;
                const float xx = x * x;
                if (x > 1.f) {
                    return ((-b - 6.f * c) * xx * x + (6.f * b + 30.f * c) * xx + (-12.f * b - 48.f * c) * x + (8.f * b + 24.f * c)) / 6.f;
                }
                return ((12.f - 9.f * b - 6.f * c) * xx * x + (-18.f + 12.f * b + 6.f * c) * xx + (6.f - 2.f * b)) / 6.f;
            }
            // Blackman, sensor.zig:27-40
            const float a0 = 0.35875f, a1 = 0.48829f, a2 = 0.14128f, a3 = 0.01168f;
            const float b  = (kPi * (x + radius)) / radius;
            return a0 - a1 * std::cos(b) + a2 * std::cos(2.f * b) - a3 * std::cos(3.f * b);
        };

        float s = 0.f;
        for (int i = 0; i < 30; ++i) {
            v.filter[i] = evalFilter(s);
            s += interval;
        }

        if (radius > 0.f) {
            auto eval = [&](float x) -> float {  // InterpolatedFunction1DN.eval, :131-143
                const float    cx     = fmin_(std::fabs(x), v.filter_range_end);
                const float    o      = cx * v.filter_inverse_interval;
                const uint32_t offset = uint32_t(o);
                const float    t      = o - float(offset);
                const float    u      = 1.f - t;
                return std::fmaf(u, v.filter[offset], t * v.filter[std::min(offset + 1, 29u)]);
            };
            // Sensor.integral(64, radius), sensor.zig:630-645
            const float ival = radius / 64.f;
            float       x    = 0.5f * ival;
            float       sum  = 0.f;
            for (int i = 0; i < 64; ++i) {
                const float a = eval(x) * ival;
                sum += a;
                x += ival;
            }
            const float scale = 1.f / (sum + sum);
            for (int i = 0; i < 30; ++i) v.filter[i] *= scale;
        }
    }
    v.exposure_factor = 1.f;  // Tonemapper.init(.Linear, 0.0): exp2(0)
    v.aov_slots       = aov_slots_;
    v.alpha_transparency = alpha_transparency_ ? 1u : 0u;

    if (!ptmis_) {
        error = "only the PTMIS surface integrator is implemented: call su_integrators_create with {\"surface\":{\"PTMIS\":{}}}";
        return false;
    }
    return true;
}

}  // namespace zyg
