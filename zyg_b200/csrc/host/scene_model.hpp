// Host-side scene model of the backend: what a zyg host keeps between su_* calls (Scene, Take.View,
// the material / shape resources) reduced to the parts the surface-integration path consumes, plus
// `compile`, which flattens it into the ZygpuScene / ZygpuView arrays of include/zygpu_scene.h.
//
// Mirrors, for static scenes:
//   Scene.createEntity / createPropShape / createLight / classifyProp   src/core/scene/scene.zig:254-380
//   Scene.compile, propPrepareSampling                                  scene.zig:185-223, 402-497
//   Space.calculateWorldBounds / transformationAtMaybeStatic            src/core/scene/space.zig:60-111
//   Prop.configureShape / setVisibility                                 src/core/scene/prop/prop.zig:78-135
//   PropBvhBuilder                                                      src/core/scene/prop/prop_tree_builder.zig
//   material JSON                                                       src/core/scene/material/material_provider.zig
//   View.loadIntegrators, Sensor.init, Perspective.update               take.zig:131-271, sensor.zig:106-124,
//                                                                       camera_perspective.zig:79-122
#pragma once

#include "mesh_sampler.hpp"

#include <memory>

#include "../../../include/zygpu_scene.h"
#include "json.hpp"
#include "zmath.hpp"

#include <string>
#include <array>
#include <vector>

struct zyg_mesh;

namespace zyg {

struct Transformation {  // src/base/math/transformation.zig
    Vec4f position{{0.f, 0.f, 0.f, 0.f}};
    Vec4f scale{{1.f, 1.f, 1.f, 1.f}};
    Vec4f rotation{{0.f, 0.f, 0.f, 1.f}};  // quaternion
};

struct Mat3x3 {
    Vec4f r[3];
};

Mat3x3 quaternionToMat3x3(Vec4f q);                    // quaternion.zig:73-112
Vec4f  quaternionFromMat3x3(const Mat3x3& m);          // quaternion.zig:14-37
Mat3x3 rotationFromEulerDegrees(Vec4f xyz);            // json.zig:169-175 (createRotationMatrix)
void   decomposeMatrix(const float m[16], Transformation& out);  // matrix4x4.zig:109-121 + capi.zig:485-503
Vec4f  sRGBtoAP1(Vec4f srgb);                          // src/base/spectrum/aces.zig:9-17

enum class FilterKind { None, Blackman, Mitchell };

class SceneModel {
  public:
    SceneModel();

    // ---- resources ----
    // `rendering` is the JSON object holding one of "Substitute" / "Light" / "Glass" / "Debug"
    // (material_provider.zig:131-161). Returns the material id or -1.
    int  createMaterial(const json::Value& material);
    bool updateMaterial(uint32_t id, const json::Value& material);
    // su_image_create / su_image_update, capi.zig:236-340: the library copies the pixels. Float32 x 3 (image.Float3) and
    // UInt8 x 3 (image.Byte3, read as sRGB like Texture.Byte3_sRGB) are the formats an emission map can have here.
    int  createImage(uint32_t id, uint32_t format, uint32_t num_channels, uint32_t width, uint32_t height, uint32_t depth,
                     uint32_t pixel_stride, const uint8_t* data);
    int  updateImage(uint32_t id, uint32_t pixel_stride, const uint8_t* data);
    // Registers a compiled mesh as a shape resource; returns the shape id (>= 7).
    uint32_t addMesh(const zyg_mesh* mesh, uint32_t num_parts);
    // The mesh of a shape registered with a null handle (an asynchronous build that has finished since).
    void     setMesh(uint32_t shape, const zyg_mesh* mesh) {
        meshes_[shape - 7].mesh = mesh;
        touch();
    }
    // Messages about accepted but ignored input (unsupported material parameters); the caller logs and clears them.
    std::vector<std::string>& warnings() { return warnings_; }
    // Counts the edits: a caller that kept the products of compile() may skip the next compile while it is unchanged.
    uint64_t revision() const { return revision_; }
    uint32_t numShapes() const { return 7 + uint32_t(meshes_.size()); }
    uint32_t numMaterials() const { return uint32_t(materials_.size()); }
    uint32_t fallbackMaterial() const { return 0; }

    // ---- scene ----
    uint32_t createEntity();  // scene.zig:254-260
    uint32_t createPropShape(uint32_t shape_id, const uint32_t* materials, uint32_t num_materials, bool unoccluding);
    int      createPropInstance(uint32_t entity);  // scene.zig:310-320
    // scene_loader.loadInstancer + Scene.createPropInstancer (src/util/scene_loader.zig:401-508, scene.zig:292-304): the
    // prototype entities leave the scene's own prop tree, every instance places one of them with a transformation relative to
    // the instancer entity. Compile flattens the instancer: each instance becomes a prop with the composed transformation
    // (ComposedTransformation.transform, composed_transformation.zig:55-68), so the device sees one two-level layout.
    int      createInstancer(const uint32_t* prototypes, uint32_t num_prototypes, const uint32_t* prototype_indices,
                             const Transformation* transformations, uint32_t num_instances);
    bool     createLight(uint32_t entity);
    bool     setWorldTransformation(uint32_t entity, const Transformation& t);
    bool     setVisibility(uint32_t entity, bool in_camera, bool in_reflection, bool in_sss);
    uint32_t numProps() const { return uint32_t(props_.size()); }

    // ---- view ----
    void setCamera(uint32_t width, uint32_t height);  // su_perspective_camera_create
    void setFov(float radians) {
        fov_ = radians;
        touch();
    }
    void setLens(float aperture_radius, float focus_distance) {
        aperture_radius_ = aperture_radius;
        focus_distance_  = focus_distance;
        touch();
    }
    // The take's camera "crop" (take_loader.zig camera block -> Base.setResolution, camera_base.zig:32-41): x0, y0, x1, y1 with
    // x1 / y1 exclusive; clamped to the resolution at compile time. A negative x1 restores the full frame.
    void setCrop(int32_t x0, int32_t y0, int32_t x1, int32_t y1) {
        crop_[0] = x0, crop_[1] = y0, crop_[2] = x1, crop_[3] = y1;
        touch();
    }
    void setSamplesPerPixel(uint32_t spp) {
        spp_ = spp;
        touch();
    }
    void loadIntegrators(const json::Value& value);  // take.zig:131-148
    void loadSensor(const json::Value& value);       // take_loader.zig:186-232
    void loadSampler(const json::Value& value);      // take_loader.zig:142-158
    void loadAovs(const json::Value& value);         // View.loadAOV, take.zig:106-129
    uint32_t aovSlots() const { return aov_slots_; }
    bool     alphaTransparency() const { return alpha_transparency_; }
    uint32_t cameraEntity() const { return camera_entity_; }
    uint32_t width() const { return uint32_t(resolution_[0]); }
    uint32_t height() const { return uint32_t(resolution_[1]); }
    uint32_t samplesPerPixel() const { return spp_; }

    // Scene.compile + camera.update for the current camera position. The returned records point into
    // storage owned by the model and stay valid until the next compile or edit.
    bool compile(std::string& error);

    const ZygpuScene& scene() const { return flat_; }
    const ZygpuView&  view() const { return view_; }

  private:
    void touch() { revision_ += 1; }
    std::vector<std::string> warnings_;
    uint64_t revision_ = 0;
    int32_t  crop_[4]  = {0, 0, -1, -1};

    struct PropRec {
        uint32_t shape       = ZYGPU_NULL;
        uint32_t flags       = ZYG_PROP_VISIBLE_IN_CAMERA | ZYG_PROP_VISIBLE_IN_REFLECTION | ZYG_PROP_VISIBLE_IN_SHADOW;
        uint32_t parts_start = 0;
        bool     solid       = true;
        bool     classified  = false;
    };
    struct MeshRec {
        const zyg_mesh* mesh;
        uint32_t        num_parts;
        // Mesh.prepareSampling / calculateAreas products, filled when a part of the mesh first becomes a light
        std::vector<uint32_t> primitive_mapping;
        std::vector<float>    part_areas;
    };
    struct SamplerRec {  // ShapeSamplerCache entry (shape_sampler_cache.zig:131-172), keyed by mesh, part and sidedness
        uint32_t        mesh, part;
        bool            two_sided;
        MeshSamplerData data;
    };

    struct InstancerRec {  // prop/instancer.zig:22-50
        uint32_t                    entity;
        std::vector<uint32_t>       prototypes;  // per instance: the prototype entity
        std::vector<Transformation> trafos;      // per instance, relative to the instancer entity
    };
    struct ImageRec {  // image.Float1 / Float2 / Float3 / Byte1 / Byte2 / Byte3
        uint32_t           width = 0, height = 0, format = 0, channels = 3;
        std::vector<float> pixels;  // float triples: RGB in ACEScg, or the 1 / 2 channels of a scalar / normal map in the first slots
    };
    struct EmissionMapRec {  // Emittance.emission_map when it is an image (Texture + Texture.Mode), per material
        uint32_t image     = ZYGPU_NULL;
        uint32_t address_u = 1, address_v = 1, filter = 1;  // Texture.DefaultMode: Repeat, Repeat, LinearStochastic
        float    scale[2]  = {1.f, 1.f};
    };
    struct ImageSamplerRec {  // shape_sampler.ImageImpl of one (material, uv-weight class of the shape) pair
        uint32_t           material, weight_class;
        float              total_weight;
        Vec4f              average_emission;
        std::vector<float> marginal_cdf, conditional_cdf, conditional_integral;
    };
    static uint32_t uvWeightClass(uint32_t shape);
    uint32_t        imageSampler(uint32_t material, uint32_t shape);

    bool shapeFinite(uint32_t shape) const;
    AABB shapeAabb(uint32_t shape) const;
    void buildPropTree(const std::vector<uint32_t>& indices, std::vector<ZygpuBvhNode>& nodes, std::vector<uint32_t>& out_indices);
    bool buildLightTree(std::string& error);

    std::vector<ZygpuMaterial>  materials_;
    std::vector<EmissionMapRec> emission_maps_;  // per material
    std::vector<EmissionMapRec> color_maps_;     // per material: Substitute.color as an image texture
    enum { kRoughnessMap = 0, kMetallicMap = 1, kNormalMap = 2, kNumSurfaceMaps = 3 };
    std::vector<std::array<EmissionMapRec, 3>> surface_maps_;  // per material: Substitute.roughness / metallic / normal_map as images
    std::vector<ImageRec>       images_;
    std::vector<std::unique_ptr<ImageSamplerRec>> image_samplers_;
    std::vector<ZygpuImageSampler>                flat_image_samplers_;
    std::vector<float>                            infinite_cdf_;
    std::vector<MeshRec>        meshes_;
    std::vector<PropRec>        props_;
    std::vector<Transformation> world_;
    std::vector<uint32_t>       material_ids_, light_ids_;
    std::vector<ZygpuLight>     lights_;
    std::vector<uint32_t>       finite_props_, infinite_props_, unoccluding_props_;
    std::vector<InstancerRec>   instancers_;
    std::vector<uint32_t>       flat_finite_, flat_unoccluding_;  // the classified lists with the instancers' instances appended

    // view
    int32_t  resolution_[2]   = {0, 0};
    float    fov_             = 0.f;
    float    aperture_radius_ = 0.f;
    float    focus_distance_  = 0.f;
    uint32_t camera_entity_   = ZYGPU_NULL;
    uint32_t spp_             = 1;
    uint32_t sampler_         = ZYG_SAMPLER_SOBOL;
    uint32_t aov_slots_       = 0;  // aov.Factory.slots
    bool     alpha_transparency_ = false;  // Sensor.Buffer.Class Transparent

    uint32_t max_depth_surface_ = 1, max_depth_volume_ = 1;  // take.zig:43-52 (default AOV integrator)
    float    split_threshold_   = 0.f;
    float    regularize_roughness_ = 0.f;
    bool     caustics_path_     = true;
    bool     ptmis_             = false;
    float    specular_threshold_;

    FilterKind filter_kind_   = FilterKind::Mitchell;  // take.zig:59-64
    float      filter_radius_ = 2.f;
    float      clamp_[3];

    // flattened output
    std::vector<ZygpuProp>      flat_props_;
    std::vector<ZygpuTrafo>     flat_trafos_;
    std::vector<ZygpuAabb>      flat_aabbs_;
    std::vector<ZygpuAabb>      light_aabbs_;
    std::vector<float>          light_cones_;
    std::vector<ZygpuBvhNode>   solid_nodes_, unocc_nodes_;
    std::vector<uint32_t>       solid_indices_, unocc_indices_;
    std::vector<ZygpuLightNode> light_nodes_;
    std::vector<uint32_t>       light_node_middles_, light_orders_, light_mapping_;
    std::vector<const zyg_mesh*> flat_meshes_;
    std::vector<std::unique_ptr<SamplerRec>> samplers_;
    std::vector<ZygpuMeshSampler>            flat_samplers_;
    std::vector<float>                       flat_part_areas_;
    std::vector<float>          luts_;
    ZygpuScene                  flat_{};
    ZygpuView                   view_{};
};

// Locates and reads zyg_b200/data/ggx_luts.f32 (next to the shared library). Empty on failure.
const std::vector<float>& ggxLuts(std::string& error);

}  // namespace zyg
