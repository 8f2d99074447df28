"""ctypes mirror of include/zygpu_scene.h (ZygpuScene and the records it points to) with numpy views of its arrays.
Test infrastructure: lets host tests read what zyg_su_compile produced without going through a renderer."""

from __future__ import annotations

import ctypes as C

import numpy as np

u32, f32, vp = C.c_uint32, C.c_float, C.c_void_p

PROP_DTYPE = np.dtype([("shape", "<u4"), ("mesh", "<u4"), ("flags", "<u4"), ("parts_start", "<u4")])
AABB_DTYPE = np.dtype([("min", "<f4", 4), ("max", "<f4", 4)])
NODE_DTYPE = np.dtype([("min", "<f4", 3), ("a", "<u4"), ("max", "<f4", 3), ("n", "<u4")])
LIGHT_DTYPE = np.dtype([("prop", "<u4"), ("part", "<u4"), ("light_class", "<u4"), ("two_sided", "<u4"), ("num_samples", "<u4"),
                        ("sampler", "<u4"), ("pad", "<u4", 2)])
LIGHT_NODE_DTYPE = np.dtype([("center", "<u2", 4), ("cone", "<u2", 4), ("power", "<f4"), ("variance", "<f4"), ("meta", "<u4"),
                             ("num_lights", "<u4")])
TRAFO_DTYPE = np.dtype([("r", "<f4", (3, 4)), ("position", "<f4", 4)])

SHAPE_CANOPY, SHAPE_CUBE, SHAPE_DISK, SHAPE_DISTANT, SHAPE_DOME, SHAPE_RECTANGLE, SHAPE_SPHERE, SHAPE_MESH = range(8)
PROP_UNOCCLUDING = 1 << 5


class Aabb(C.Structure):
    _fields_ = [("min", f32 * 4), ("max", f32 * 4)]


class LightTree(C.Structure):
    _fields_ = [("bounds", Aabb), ("infinite_weight", f32), ("infinite_guard", f32), ("infinite_end", u32), ("max_split_depth", u32),
                ("num_lights", u32), ("num_infinite_lights", u32), ("num_nodes", u32), ("pad", u32), ("nodes", vp),
                ("node_middles", vp), ("light_orders", vp), ("light_mapping", vp), ("infinite_cdf", vp)]


class PropTree(C.Structure):
    _fields_ = [("num_nodes", u32), ("num_indices", u32), ("nodes", vp), ("indices", vp)]


class MeshSampler(C.Structure):
    _fields_ = [("bounds", Aabb), ("num_triangles", u32), ("num_nodes", u32), ("two_sided", u32), ("mesh", u32), ("nodes", vp),
                ("node_middles", vp), ("light_orders", vp), ("light_mapping", vp), ("triangle_mapping", vp), ("triangle_pdfs", vp),
                ("primitive_mapping", vp)]


class Scene(C.Structure):
    _fields_ = [("num_props", u32), ("num_parts", u32), ("num_materials", u32), ("num_lights", u32), ("num_infinite_props", u32),
                ("num_meshes", u32), ("props", vp), ("trafos", vp), ("aabbs", vp), ("material_ids", vp), ("light_ids", vp),
                ("materials", vp), ("lights", vp), ("light_aabbs", vp), ("light_cones", vp), ("light_tree", LightTree),
                ("solid_bvh", PropTree), ("unoccluding_bvh", PropTree), ("infinite_props", vp), ("meshes", vp),
                ("num_mesh_samplers", u32), ("mesh_samplers", vp), ("mesh_part_areas", vp), ("num_image_samplers", u32),
                ("image_samplers", vp), ("ggx_luts", vp)]


def view(ptr, dtype, count):
    """numpy copy of `count` records of `dtype` at address `ptr` (0 records when the pointer is null)."""
    dtype = np.dtype(dtype)
    if not ptr or 0 == count:
        return np.zeros(0, dtype)
    buf = (C.c_char * (dtype.itemsize * count)).from_address(ptr)
    return np.frombuffer(buf, dtype, count).copy()


def scene_at(address) -> Scene:
    return Scene.from_address(address)


def mesh_samplers(scene: Scene):
    if 0 == scene.num_mesh_samplers:
        return []
    return list((MeshSampler * scene.num_mesh_samplers).from_address(scene.mesh_samplers))
