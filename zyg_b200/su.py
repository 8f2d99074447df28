"""ctypes mirror of zyg's C API (``su_*``, include/zyg_su.h) as served by ``libzyg_b200.so``.

Written the way ``src/capi-test/test.py`` drives ``libzyg.so``: module-level functions over one
process-global engine, JSON strings for materials and integrators, 4x4 row-major matrices for
transformations. No compute happens in Python.
"""

from __future__ import annotations

import ctypes as C
import json

import numpy as np

from . import lib as _lib

CANOPY, CUBE, DISK, DISTANT, DOME, RECTANGLE, SPHERE = range(7)  # src/core/resource/manager.zig:36-44

_bound = None


def _su() -> C.CDLL:
    global _bound
    if _bound is not None:
        return _bound
    lib = _lib.load_library()
    u32, i32, f32, vp, cp = C.c_uint32, C.c_int32, C.c_float, C.c_void_p, C.c_char_p
    sig = {
        "su_init": [], "su_release": [], "su_mount": [cp],
        "su_perspective_camera_create": [u32, u32], "su_camera_set_fov": [f32], "su_camera_sensor_dimensions": [vp],
        "su_exporters_create": [cp], "su_aovs_create": [cp], "su_sampler_create": [u32], "su_integrators_create": [cp],
        "su_image_create": [u32, u32, u32, u32, u32, u32, u32, vp], "su_image_update": [u32, u32, vp],
        "su_material_create": [u32, cp], "su_material_update": [u32, cp],
        "su_triangle_mesh_create": [u32, u32, vp, u32, vp, u32, vp, u32, vp, u32, vp, u32, vp, u32, C.c_bool],
        "su_prop_create": [u32, u32, vp], "su_prop_create_instance": [u32], "su_light_create": [u32],
        "su_prop_set_transformation": [u32, vp], "su_prop_set_transformation_frame": [u32, u32, vp],
        "su_prop_set_visibility": [u32, u32, u32, u32],
        "su_render_frame": [u32], "su_export_frame": [], "su_start_frame": [u32], "su_render_iterations": [u32],
        "su_resolve_frame": [u32], "su_resolve_frame_to_buffer": [u32, u32, u32, vp],
        "su_copy_framebuffer": [u32, u32, u32, u32, vp], "su_register_log": [vp], "su_register_progress": [vp, vp],
        "zyg_su_sensor_create": [cp], "zyg_su_instancer_create": [u32, vp, u32, vp, vp], "zyg_su_prop_create_unoccluding": [u32, u32, vp],
        "zyg_su_camera_set_lens": [f32, f32], "zyg_su_camera_set_crop": [i32, i32, i32, i32], "zyg_su_set_device": [i32],
        "zyg_su_render_frame_range": [u32, u32, u32], "zyg_su_compile": [vp, vp],
        "zyg_su_write_image": [cp, u32, u32, vp, i32, i32, vp],
        "zyg_su_set_mesh_builder": [i32], "zyg_su_set_light_tree_builder": [i32, u32], "zyg_su_triangle_mesh_refit": [u32, vp, u32, vp, u32],
    }
    for name, argtypes in sig.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = i32
    lib.zyg_su_device.argtypes = []
    lib.zyg_su_device.restype = vp
    lib.zyg_su_mesh.argtypes = [u32]
    lib.zyg_su_mesh.restype = vp
    _bound = lib
    return lib


class SuError(RuntimeError):
    pass


def _ok(rc: int, what: str) -> int:
    if rc < 0:
        raise SuError(f"{what} returned {rc}: {_su().zygpu_last_error().decode()}")
    return rc


MESH_BUILDER = 0  # builders every new engine starts with (set_mesh_builder / set_light_tree_builder): the scene helpers call
LIGHT_TREE_BUILDER = 0  # init() themselves


def init():
    _ok(_su().su_init(), "su_init")
    if 0 != MESH_BUILDER:
        _ok(_su().zyg_su_set_mesh_builder(MESH_BUILDER), "zyg_su_set_mesh_builder")
    if 0 != LIGHT_TREE_BUILDER:
        _ok(_su().zyg_su_set_light_tree_builder(LIGHT_TREE_BUILDER, 0), "zyg_su_set_light_tree_builder")


def release():
    return _su().su_release()


def perspective_camera_create(width: int, height: int) -> int:
    return _ok(_su().su_perspective_camera_create(width, height), "su_perspective_camera_create")


def camera_set_fov(radians: float):
    _ok(_su().su_camera_set_fov(radians), "su_camera_set_fov")


def camera_set_lens(aperture_radius: float, focus_distance: float):
    _ok(_su().zyg_su_camera_set_lens(aperture_radius, focus_distance), "zyg_su_camera_set_lens")


def camera_set_crop(x0: int, y0: int, x1: int, y1: int):
    _ok(_su().zyg_su_camera_set_crop(x0, y0, x1, y1), "zyg_su_camera_set_crop")


def sampler_create(spp: int):
    _su().su_sampler_create(spp)  # returns -1 even on success (capi.zig:215-221)


def integrators_create(desc: dict):
    _ok(_su().su_integrators_create(json.dumps(desc).encode()), "su_integrators_create")


def sensor_create(desc: dict):
    _ok(_su().zyg_su_sensor_create(json.dumps(desc).encode()), "zyg_su_sensor_create")


def image_create(pixels: np.ndarray) -> int:
    """su_image_create for an (H, W, C) float32 (Format.Float32) or uint8 (Format.UInt8) array, C = 1, 2 or 3 ((H, W) = one channel);
    the library copies. uint8 x 3 is read as sRGB, x 1 as unorm (roughness / metallic maps), x 2 as snorm (normal maps)."""
    px = np.ascontiguousarray(pixels)
    if 2 == px.ndim:
        px = px[..., None]
    assert px.ndim == 3 and 1 <= px.shape[2] <= 3 and px.dtype in (np.float32, np.uint8)
    fmt, bpc = (4, 4) if px.dtype == np.float32 else (0, 1)
    return _ok(_su().su_image_create(0xFFFFFFFF, fmt, px.shape[2], px.shape[1], px.shape[0], 1, px.shape[2] * bpc, px.ctypes.data), "su_image_create")


def image_update(image: int, pixels: np.ndarray):
    px = np.ascontiguousarray(pixels)
    channels = 1 if 2 == px.ndim else px.shape[2]
    _ok(_su().su_image_update(image, channels * px.dtype.itemsize, px.ctypes.data), "su_image_update")


def material_create(desc: dict) -> int:
    return _ok(_su().su_material_create(0xFFFFFFFF, json.dumps(desc).encode()), "su_material_create")


def material_update(material: int, desc: dict):
    _ok(_su().su_material_update(material, json.dumps(desc).encode()), "su_material_update")


def triangle_mesh_create(positions, indices, normals=None, uvs=None, parts=None) -> int:
    positions = np.ascontiguousarray(positions, np.float32)
    indices = np.ascontiguousarray(indices, np.uint32).reshape(-1)
    nv, nt = positions.shape[0], indices.size // 3
    normals = None if normals is None else np.ascontiguousarray(normals, np.float32)
    uvs = None if uvs is None else np.ascontiguousarray(uvs, np.float32)
    parts = None if parts is None else np.ascontiguousarray(parts, np.uint32).reshape(-1)
    return _ok(_su().su_triangle_mesh_create(
        0xFFFFFFFF, 0 if parts is None else parts.size // 3, None if parts is None else parts.ctypes.data, nt,
        indices.ctypes.data, nv, positions.ctypes.data, 3, None if normals is None else normals.ctypes.data, 3, None, 0,
        None if uvs is None else uvs.ctypes.data, 2, False), "su_triangle_mesh_create")


def prop_create(shape: int, materials, unoccluding: bool = False) -> int:
    mats = np.ascontiguousarray(materials, np.uint32)
    fn = _su().zyg_su_prop_create_unoccluding if unoccluding else _su().su_prop_create
    return _ok(fn(shape, mats.size, mats.ctypes.data), "su_prop_create")


def prop_create_instance(entity: int) -> int:
    return _ok(_su().su_prop_create_instance(entity), "su_prop_create_instance")


def instancer_create(prototypes, prototype_indices, matrices) -> int:
    """An "Instancer" entity (zyg_su_instancer_create): instance i = prototypes[prototype_indices[i]] placed with the 4x4
    matrices[i] relative to the returned entity."""
    protos = np.ascontiguousarray(prototypes, np.uint32)
    idx = np.ascontiguousarray(prototype_indices, np.uint32)
    m = np.ascontiguousarray(matrices, np.float32).reshape(-1, 16)
    assert m.shape[0] == idx.size
    return _ok(_su().zyg_su_instancer_create(protos.size, protos.ctypes.data, idx.size, idx.ctypes.data, m.ctypes.data),
               "zyg_su_instancer_create")


def light_create(prop: int):
    _ok(_su().su_light_create(prop), "su_light_create")


def prop_set_transformation(prop: int, matrix):
    m = np.ascontiguousarray(matrix, np.float32).reshape(16)
    _ok(_su().su_prop_set_transformation(prop, m.ctypes.data), "su_prop_set_transformation")


def prop_set_visibility(prop: int, in_camera: bool, in_reflection: bool, in_sss: bool = False):
    _ok(_su().su_prop_set_visibility(prop, int(in_camera), int(in_reflection), int(in_sss)), "su_prop_set_visibility")


def render_frame(frame: int = 0):
    _ok(_su().su_render_frame(frame), "su_render_frame")


def render_frame_range(frame: int, iteration: int, num_samples: int):
    _ok(_su().zyg_su_render_frame_range(frame, iteration, num_samples), "zyg_su_render_frame_range")


def start_frame(frame: int = 0):
    _ok(_su().su_start_frame(frame), "su_start_frame")


def render_iterations(n: int):
    _ok(_su().su_render_iterations(n), "su_render_iterations")


def resolve_frame_to_buffer(width: int, height: int, aov: int = 0xFFFFFFFF) -> np.ndarray:
    """su_resolve_frame_to_buffer: the beauty (any aov >= 9) or one AOV class (0..8, aov.Value.Class order); SuError(-2) when the
    class is not recorded."""
    out = np.empty((height, width, 4), np.float32)
    _ok(_su().su_resolve_frame_to_buffer(aov, width, height, out.ctypes.data), "su_resolve_frame_to_buffer")
    return out


AOV_ALBEDO, AOV_DEPTH, AOV_MATERIAL_ID, AOV_GEOMETRIC_NORMAL, AOV_SHADING_NORMAL, AOV_ROUGHNESS, AOV_EMISSION, AOV_DIRECT, AOV_INDIRECT = range(9)


def denoise_frame_to_buffer(sigma: float, width: int, height: int) -> np.ndarray:
    """zyg_su_denoise_frame_to_buffer: `it --denoise sigma` over the frame just rendered (needs the ShadingNormal and Albedo AOVs)."""
    fn = _su().zyg_su_denoise_frame_to_buffer
    fn.argtypes = [C.c_float, C.c_uint32, C.c_uint32, C.c_void_p]
    out = np.empty((height, width, 4), np.float32)
    _ok(fn(sigma, width, height, out.ctypes.data), "zyg_su_denoise_frame_to_buffer")
    return out


def aovs_create(desc: dict):
    """The take's "aov" block, e.g. {"Albedo": true, "Depth": true} (View.loadAOV, take.zig:106-129)."""
    _ok(_su().su_aovs_create(json.dumps(desc).encode()), "su_aovs_create")


HOST_BUILDER, DEVICE_BUILDER = 0, 1


def set_mesh_builder(builder: int):
    """0: host SAH build in reference order (default); 1: LBVH built on the device (zygpu_mesh_build)."""
    _ok(_su().zyg_su_set_mesh_builder(builder), "zyg_su_set_mesh_builder")


def set_light_tree_builder(builder: int, min_lights: int = 0):
    """0: host restatement of the reference's light-tree builder (default); 1: device builder for trees of >= min_lights lights."""
    _ok(_su().zyg_su_set_light_tree_builder(builder, min_lights), "zyg_su_set_light_tree_builder")


def triangle_mesh_refit(shape: int, positions, normals=None):
    p = np.ascontiguousarray(positions, np.float32).reshape(-1, 3)
    n = None if normals is None else np.ascontiguousarray(normals, np.float32).reshape(-1, 3)
    _ok(_su().zyg_su_triangle_mesh_refit(shape, p.ctypes.data, 3, None if n is None else n.ctypes.data, 3), "zyg_su_triangle_mesh_refit")


def exporters_create(desc: dict):
    """The take's "export" block, e.g. {"Image": {"format": "EXR", "bitdepth": 32}} (take.zig:303-331)."""
    _ok(_su().su_exporters_create(json.dumps(desc).encode()), "su_exporters_create")


def export_frame():
    """Writes image_00_<frame:06>.<ext> per exporter into the working directory (driver.zig:224-253)."""
    _ok(_su().su_export_frame(), "su_export_frame")


IMAGE_PNG, IMAGE_EXR, IMAGE_RGBE = 0, 1, 2
IMAGE_ALPHA, IMAGE_HALF, IMAGE_ERROR_DIFFUSION = 1, 2, 4
IMAGE_DEPTH, IMAGE_ID, IMAGE_NORMAL, IMAGE_FLOAT = 2 << 8, 3 << 8, 4 << 8, 5 << 8  # Writer.Encoding of an AOV layer


def write_image(path: str, fmt: int, rgba, flags: int = 0, crop=None):
    """The codecs behind su_export_frame on a caller-owned (H, W, 4) float image; needs no engine and no GPU."""
    img = np.ascontiguousarray(rgba, np.float32)
    c = None if crop is None else np.ascontiguousarray(crop, np.int32)
    _ok(_su().zyg_su_write_image(path.encode(), fmt, flags, img.ctypes.data, img.shape[1], img.shape[0],
                                 None if c is None else c.ctypes.data), "zyg_su_write_image")


def compile_scene():
    """Scene.compile + camera.update on the host. Returns (ZygpuScene*, ZygpuView*) as integers for the
    oracle and the device ABI; valid until the scene is edited."""
    scene, view = C.c_void_p(), C.c_void_p()
    _ok(_su().zyg_su_compile(C.byref(scene), C.byref(view)), "zyg_su_compile")
    return scene.value, view.value


def device_handle():
    return _su().zyg_su_device()


def transformation(position=(0, 0, 0), scale=(1, 1, 1), rotation_deg=(0, 0, 0)) -> np.ndarray:
    """Row-major 4x4 in the layout su_prop_set_transformation expects: rows 0-2 = scaled basis vectors, row 3 =
    position. The rotation follows json.createRotationMatrix (src/base/json.zig:169-175): Rz * Rx * Ry."""
    ax, ay, az = np.radians(np.asarray(rotation_deg, np.float64))
    rx = np.array([[1, 0, 0], [0, np.cos(ax), -np.sin(ax)], [0, np.sin(ax), np.cos(ax)]])
    ry = np.array([[np.cos(ay), 0, np.sin(ay)], [0, 1, 0], [-np.sin(ay), 0, np.cos(ay)]])
    rz = np.array([[np.cos(az), -np.sin(az), 0], [np.sin(az), np.cos(az), 0], [0, 0, 1]])
    rot = rz @ rx @ ry
    m = np.zeros((4, 4), np.float32)
    m[:3, :3] = (rot * np.asarray(scale, np.float64)[:, None]).astype(np.float32)
    m[3, :3] = position
    m[3, 3] = 1
    return m
