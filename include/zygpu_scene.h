/* zygpu_scene.h — the flattened ("compiled") scene and view a zyg host hands to the device.
 *
 * zyg compiles its scene graph on the host once per frame (Scene.compile, src/core/scene/scene.zig:185-223)
 * and the hot path only reads the result. These PODs are that result, laid out as plain arrays: the
 * same records the Zig structs hold, camera-relative (space.zig:94,103-110), with every pointer
 * replaced by an index. A Zig host fills them from its own `Scene` / `View`; the C++ host model in
 * zyg_b200/csrc/host/scene_model.cpp (behind the su_* API) fills them the same way.
 *
 * All colours are ACEScg (AP1) like inside zyg (src/base/json.zig:247-249).
 */
#ifndef ZYGPU_SCENE_H
#define ZYGPU_SCENE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ZYGPU_NULL 0xFFFFFFFFu

/* ComposedTransformation (src/core/scene/composed_transformation.zig:15-17): three rotation rows
 * whose 4th lane carries the scale of that axis, then the position. 64 bytes. */
typedef struct ZygpuTrafo {
    float r[3][4];
    float position[4];
} ZygpuTrafo;

/* math.AABB (src/base/math/aabb.zig): bounds[0], bounds[1]; bounds[1][3] = cached radius,
 * bounds[0][3] = light power for light boxes (scene.zig:496). */
typedef struct ZygpuAabb {
    float min[4];
    float max[4];
} ZygpuAabb;

/* bvh.Node (src/core/scene/bvh/node.zig:9-20). */
typedef struct ZygpuBvhNode {
    float    min[3];
    uint32_t children_or_start;
    float    max[3];
    uint32_t num_indices; /* 0 => inner node */
} ZygpuBvhNode;

/* Built-in shape ids of the resource manager (src/core/resource/manager.zig:36-44). */
enum {
    ZYG_SHAPE_CANOPY        = 0,
    ZYG_SHAPE_CUBE          = 1,
    ZYG_SHAPE_DISK          = 2,
    ZYG_SHAPE_DISTANT       = 3,
    ZYG_SHAPE_DOME          = 4,
    ZYG_SHAPE_RECTANGLE     = 5,
    ZYG_SHAPE_SPHERE        = 6,
    ZYG_SHAPE_TRIANGLE_MESH = 7
};

/* Prop.Properties (src/core/scene/prop/prop.zig:24-48), the bits the hot path reads. */
enum {
    ZYG_PROP_VISIBLE_IN_CAMERA     = 1u << 0,
    ZYG_PROP_VISIBLE_IN_REFLECTION = 1u << 1,
    ZYG_PROP_VISIBLE_IN_SHADOW     = 1u << 2,
    ZYG_PROP_EVALUATE_VISIBILITY   = 1u << 4,
    ZYG_PROP_UNOCCLUDING           = 1u << 5
};

typedef struct ZygpuProp {
    uint32_t shape;       /* ZYG_SHAPE_* */
    uint32_t mesh;        /* index into ZygpuScene.meshes when shape == ZYG_SHAPE_TRIANGLE_MESH */
    uint32_t flags;       /* ZYG_PROP_* */
    uint32_t parts_start; /* Scene.prop_parts: first entry in material_ids / light_ids */
} ZygpuProp;

enum { /* Material union tags (src/core/scene/material/material.zig:35-42) */
    ZYG_MATERIAL_DEBUG      = 0,
    ZYG_MATERIAL_GLASS      = 1,
    ZYG_MATERIAL_LIGHT      = 3,
    ZYG_MATERIAL_SUBSTITUTE = 5
};

enum { /* material_base.zig:17-26 */
    ZYG_MATERIAL_TWO_SIDED = 1u << 0,
    ZYG_MATERIAL_CAUSTIC   = 1u << 2,
    ZYG_MATERIAL_EMISSIVE  = 1u << 3
};

/* Substitute / Glass / Light parameters: uniform values plus the image maps in scope (emission map,
 * Substitute colour, roughness, metallic and normal maps) and the clear coat of a Substitute. 144 bytes. */
typedef struct ZygpuMaterial {
    uint32_t type;  /* ZYG_MATERIAL_* */
    uint32_t flags; /* ZYG_MATERIAL_* bits */
    int32_t  priority;
    uint32_t emission_num_samples; /* Emittance.num_samples (emittance.zig:25) */

    float color[4];    /* Substitute.color; Glass: absorption coefficient (glass_material.zig) */
    float emission[4]; /* Emittance.value */

    float roughness; /* un-clamped; ggx.clampRoughness is applied per sample like the reference */
    float metallic;
    float specular;
    float ior;

    float anisotropy;
    float emission_cos_a;
    float emission_camera_weight;
    float emission_normalize; /* 0 / 1 */

    float attenuation_distance;
    float thickness;
    float abbe;
    uint32_t emission_map; /* Emittance.emission_map when it is an image: index into ZygpuScene.image_samplers, else ZYGPU_NULL */

    uint32_t color_map; /* Substitute.color when it is an image (substitute_material.zig:120): index into image_samplers, else ZYGPU_NULL */
    uint32_t roughness_map; /* Substitute.roughness as an image (ts.sample2D_1, substitute_material.zig:122): first channel */
    uint32_t metallic_map;  /* Substitute.metallic as an image (:123): first channel */
    uint32_t normal_map;    /* Substitute.normal_map (hlp.sampleNormal, material_helper.zig:16-79): first two channels = tangent-space xy */

    /* Substitute "coating" (material_provider.zig:303-326, substitute_coating.zig): uniform parameters, scale 1 */
    float coating_absorption[3]; /* coating_absorption_coef = attenuationCoefficient(color, attenuation_distance), substitute_material.zig:81-83 */
    float coating_thickness;     /* 0: no coating */
    float coating_ior;
    float coating_roughness; /* un-clamped */
    float pad[2];
} ZygpuMaterial;

enum { /* Light.Class (src/core/scene/light/light.zig:34-40) */
    ZYG_LIGHT_PORTAL_IMAGE = 0,
    ZYG_LIGHT_PROP         = 1,
    ZYG_LIGHT_PROP_IMAGE   = 2
};

typedef struct ZygpuLight {
    uint32_t prop;
    uint32_t part;
    uint32_t light_class;
    uint32_t two_sided;
    uint32_t num_samples; /* ShapeSampler.num_samples (shape_sampler.zig:33) */
    uint32_t sampler;     /* Light.sampler: index into ZygpuScene.mesh_samplers for a triangle-mesh light, into
                             ZygpuScene.image_samplers for a ZYG_LIGHT_PROP_IMAGE light, else ZYGPU_NULL */
    uint32_t pad[2];
} ZygpuLight;

/* light_tree.Node (src/core/scene/light/light_tree.zig:25-37). 32 bytes. */
typedef struct ZygpuLightNode {
    uint16_t center[4]; /* unorm16 in the tree bounds; w = radius / bounds radius */
    uint16_t cone[4];   /* snorm16 */
    float    power;
    float    variance;
    uint32_t meta; /* bit0 has_children, bit1 two_sided, bits 2..31 children_or_light */
    uint32_t num_lights;
} ZygpuLightNode;

/* shape_sampler.MeshImpl (src/core/scene/shape/shape_sampler.zig:149-262) of one emissive mesh part with its
 * PrimitiveTree (light_tree.zig:520-719): what Mesh.sampleTo / Mesh.pdf (triangle_mesh.zig:492-608, 662-703) read. */
typedef struct ZygpuMeshSampler {
    ZygpuAabb bounds; /* PrimitiveTree.bounds: box of the emitting triangles (object space), radius cached in max[3] */
    uint32_t  num_triangles; /* triangles of the part */
    uint32_t  num_nodes;
    uint32_t  two_sided;
    uint32_t  mesh; /* index into ZygpuScene.meshes */
    const struct ZygpuLightNode* nodes;
    const uint32_t* node_middles;
    const uint32_t* light_orders;      /* per part triangle */
    const uint32_t* light_mapping;     /* tree order -> part triangle */
    const uint32_t* triangle_mapping;  /* part triangle -> BVH-order triangle (Part.triangle_mapping) */
    const float*    triangle_pdfs;     /* Distribution1D.pdfI per part triangle (relative area) */
    const uint32_t* primitive_mapping; /* BVH-order triangle -> index within its part (Mesh.primitive_mapping) */
} ZygpuMeshSampler;

/* An emission image with its importance-sampling tables: Texture (Float3 image + sampler_mode.zig Mode) and
 * shape_sampler.ImageImpl (shape_sampler.zig:128-152) = Distribution2D (src/base/math/distribution_2d.zig) over the
 * MIS-compensated luminance (light_material.zig:54-119, 248-272). The CDFs are the reference's arrays
 * (Distribution1D.precomputePdfCdf, distribution_1d.zig:93-131); its lookup tables (initLut) only accelerate the linear
 * search and are rebuilt by whoever wants them. */
typedef struct ZygpuImageSampler {
    uint32_t width, height;
    uint32_t address_u, address_v; /* Texture.Mode.Address: 0 Clamp, 1 Repeat */
    uint32_t filter;               /* Texture.Mode.Filter: 0 Nearest, 1 LinearStochastic */
    float    total_weight;         /* ImageImpl.total_weight = sum of Shape.uvWeight over the texels */
    float    scale[2];             /* Texture.data.image.scale */
    const float* pixels;           /* width * height float triples: RGB (ACEScg); a 1- or 2-channel image fills the first channels */
    const float* marginal_cdf;     /* height + 1; NULL (with the two arrays below) for an image that is only looked up, never sampled */
    const float* conditional_cdf;  /* height rows of width + 1 */
    const float* conditional_integral; /* height; 0 => that row is the degenerate distribution {1, 1} (distribution_1d.zig:99-110) */
} ZygpuImageSampler;

typedef struct ZygpuLightTree {
    ZygpuAabb             bounds;
    float                 infinite_weight;
    float                 infinite_guard;
    uint32_t              infinite_end;
    uint32_t              max_split_depth;
    uint32_t              num_lights;
    uint32_t              num_infinite_lights;
    uint32_t              num_nodes;
    uint32_t              pad;
    const ZygpuLightNode* nodes;
    const uint32_t*       node_middles;
    const uint32_t*       light_orders;
    const uint32_t*       light_mapping;
    const float*          infinite_cdf; /* Tree.infinite_light_distribution.cdf: num_infinite_lights + 1 entries (light_tree.zig:273) */
} ZygpuLightTree;

/* PropBvh.Tree (src/core/scene/prop/prop_tree.zig:31-36). */
typedef struct ZygpuPropTree {
    uint32_t            num_nodes;
    uint32_t            num_indices;
    const ZygpuBvhNode* nodes;
    const uint32_t*     indices;
} ZygpuPropTree;

struct zyg_mesh;

/* The compiled scene. Every array is host memory owned by the caller; zygpu_upload_scene copies. */
typedef struct ZygpuScene {
    uint32_t num_props;
    uint32_t num_parts; /* length of material_ids / light_ids */
    uint32_t num_materials;
    uint32_t num_lights;
    uint32_t num_infinite_props;
    uint32_t num_meshes;

    const ZygpuProp*     props;
    const ZygpuTrafo*    trafos; /* Space.transformationAtMaybeStatic result: already minus the camera position */
    const ZygpuAabb*     aabbs;  /* Space.aabbs */
    const uint32_t*      material_ids;
    const uint32_t*      light_ids;
    const ZygpuMaterial* materials;

    const ZygpuLight* lights;
    const ZygpuAabb*  light_aabbs;
    const float*      light_cones; /* 4 per light */
    ZygpuLightTree    light_tree;

    ZygpuPropTree   solid_bvh;
    ZygpuPropTree   unoccluding_bvh;
    const uint32_t* infinite_props;

    const struct zyg_mesh* const* meshes; /* compiled meshes referenced by ZygpuProp.mesh */

    uint32_t                num_mesh_samplers;
    const ZygpuMeshSampler* mesh_samplers; /* referenced by ZygpuLight.sampler */
    const float*            mesh_part_areas; /* Part.area (object space) per ZygpuScene part entry (material_ids index), 0 for analytic shapes */

    uint32_t                 num_image_samplers;
    const ZygpuImageSampler* image_samplers; /* referenced by ZygpuMaterial.emission_map and ZygpuLight.sampler */

    /* ggx_integral.zig tables, concatenated: E_m[32*32], E_m_avg[32], E[16^3], E_avg[16*16], E_s[16^3]. */
    const float* ggx_luts;
} ZygpuScene;

#define ZYGPU_GGX_LUT_FLOATS (32 * 32 + 32 + 16 * 16 * 16 + 16 * 16 + 16 * 16 * 16)

enum { ZYG_SAMPLER_RANDOM = 0, ZYG_SAMPLER_SOBOL = 1 };

/* What Take.View + Perspective.update + Sensor.init leave for the hot path
 * (src/core/take/take.zig:40-75, camera_perspective.zig:79-122, sensor.zig:106-124). */
typedef struct ZygpuView {
    int32_t resolution[2];
    int32_t crop[4]; /* x0, y0, x1, y1 (exclusive), camera_base.zig:38-49 */

    float      left_top[4];
    float      d_x[4];
    float      d_y[4];
    float      eye_offset[4];
    ZygpuTrafo camera_trafo; /* camera entity, camera-relative => position 0 */
    float      aperture_radius;
    float      focus_distance;

    uint32_t sampler;   /* ZYG_SAMPLER_* */
    uint32_t spp_total; /* View.num_samples_per_pixel = num_expected_samples (worker.zig:110) */

    uint32_t max_depth_surface;
    uint32_t max_depth_volume;
    float    split_threshold; /* already st^4 (take.zig:263-271) */
    float    regularize_roughness;
    uint32_t caustics_path;
    float    specular_threshold; /* resource manager, = ggx.MinAlpha by default */

    float clamp_emission, clamp_direct, clamp_indirect; /* Sensor.Clamp */

    int32_t filter_radius_int;
    float   filter_range_end;
    float   filter_inverse_interval;
    float   filter[30];      /* InterpolatedFunction1DN(30) samples, already normalised */
    float   exposure_factor; /* Tonemapper.exposure_factor, Linear class */

    uint32_t aov_slots; /* aov.Factory.slots (rendering/sensor/aov/aov_value.zig:84-103): bit c = AOV class c is recorded */
    uint32_t alpha_transparency; /* sensor "alpha_transparency" (take_loader.zig:194-196): the Transparent buffer keeps the alpha of
                                    Pool.transparency (vertex.zig:243-268) next to the colour (buffer_transparent.zig) */
} ZygpuView;

enum { /* aov.Value.Class, aov_value.zig:10-20 */
    ZYG_AOV_ALBEDO           = 0,
    ZYG_AOV_DEPTH            = 1,
    ZYG_AOV_MATERIAL_ID      = 2,
    ZYG_AOV_GEOMETRIC_NORMAL = 3,
    ZYG_AOV_SHADING_NORMAL   = 4,
    ZYG_AOV_ROUGHNESS        = 5,
    ZYG_AOV_EMISSION         = 6,
    ZYG_AOV_DIRECT           = 7,
    ZYG_AOV_INDIRECT         = 8,
    ZYG_AOV_NUM_CLASSES      = 9
};

/* Records that are the reference's own, byte for byte: their sizes are pinned to the numbers the reference checks for itself
 * (src/core/size_test.zig:36-47: ComposedTransformation 64, BvhNode 32, LightNode 32, Pack4f 16). */
#if defined(__cplusplus)
static_assert(sizeof(ZygpuTrafo) == 64, "ComposedTransformation");
static_assert(sizeof(ZygpuBvhNode) == 32, "bvh.Node");
static_assert(sizeof(ZygpuLightNode) == 32, "light_tree.Node");
static_assert(sizeof(ZygpuAabb) == 32, "math.AABB = 2 x Vec4f");
static_assert(sizeof(ZygpuMaterial) == 144, "ZygpuMaterial");
#else
_Static_assert(sizeof(ZygpuTrafo) == 64, "ComposedTransformation");
_Static_assert(sizeof(ZygpuBvhNode) == 32, "bvh.Node");
_Static_assert(sizeof(ZygpuLightNode) == 32, "light_tree.Node");
_Static_assert(sizeof(ZygpuAabb) == 32, "math.AABB = 2 x Vec4f");
_Static_assert(sizeof(ZygpuMaterial) == 144, "ZygpuMaterial");
#endif

#ifdef __cplusplus
}
#endif

#endif /* ZYGPU_SCENE_H */
