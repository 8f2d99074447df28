"""ctypes loader for the CPU oracle (oracle/libzyg_oracle.so). Test infrastructure only."""

from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_PATH = os.path.join(ORACLE_DIR, "libzyg_oracle.so")

RAY_DTYPE = np.dtype([("origin", "<f4", 3), ("min_t", "<f4"), ("direction", "<f4", 3), ("max_t", "<f4")])
HIT_DTYPE = np.dtype([("t", "<f4"), ("u", "<f4"), ("v", "<f4"), ("primitive", "<u4")])

_lib = None


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(ORACLE_PATH):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "-s"])
    lib = C.CDLL(ORACLE_PATH)
    vp, u64 = C.c_void_p, C.c_uint64
    lib.zo_trace_closest.argtypes = [vp, vp, vp, vp, u64, vp, C.c_uint32, C.POINTER(u64), C.POINTER(u64)]
    lib.zo_trace_closest.restype = None
    lib.zo_trace_any.argtypes = [vp, vp, vp, vp, u64, vp, C.c_uint32]
    lib.zo_trace_any.restype = None
    lib.zo_brute_closest.argtypes = [vp, C.c_uint32, vp, vp, u64, vp, vp, C.c_uint32]
    lib.zo_brute_closest.restype = None
    lib.zo_pcg32_uints.argtypes = [u64, u64, C.c_uint32, vp]
    lib.zo_pcg32_uints.restype = None
    lib.zo_pcg32_floats.argtypes = [u64, u64, C.c_uint32, vp]
    lib.zo_pcg32_floats.restype = None
    _lib = lib
    return lib


def _p(a: np.ndarray) -> int:
    assert a.flags.c_contiguous
    return a.ctypes.data


def trace_closest(nodes, triangles, positions, rays, threads=0, count=False):
    lib = load()
    out = np.empty(rays.shape[0], HIT_DTYPE)
    vn, tt = C.c_uint64(), C.c_uint64()
    lib.zo_trace_closest(_p(nodes), _p(triangles), _p(positions), _p(rays), rays.shape[0], _p(out), threads,
                         C.byref(vn) if count else None, C.byref(tt) if count else None)
    return (out, vn.value, tt.value) if count else out


def trace_any(nodes, triangles, positions, rays, threads=0):
    lib = load()
    out = np.empty(rays.shape[0], np.uint32)
    lib.zo_trace_any(_p(nodes), _p(triangles), _p(positions), _p(rays), rays.shape[0], _p(out), threads)
    return out


def brute_closest(triangles, positions, rays, threads=0):
    lib = load()
    out = np.empty(rays.shape[0], HIT_DTYPE)
    ties = np.empty(rays.shape[0], np.uint32)
    lib.zo_brute_closest(_p(triangles), triangles.size // 3, _p(positions), _p(rays), rays.shape[0], _p(out), _p(ties),
                         threads)
    return out, ties


def pcg32_uints(state, sequence, n):
    out = np.empty(n, np.uint32)
    load().zo_pcg32_uints(state, sequence, n, _p(out))
    return out


def pcg32_floats(state, sequence, n):
    out = np.empty(n, np.float32)
    load().zo_pcg32_floats(state, sequence, n, _p(out))
    return out


# ---- forward surface-integration pass (oracle/render.cpp) ---------------------------------------------

def _bind_render(lib):
    if getattr(lib, "_render_bound", False):
        return lib
    vp, u32 = C.c_void_p, C.c_uint32
    lib.zo_render.argtypes = [vp, vp, vp, u32, u32, C.c_int, vp, u32]
    lib.zo_render.restype = None
    lib.zo_resolve.argtypes = [vp, vp, u32, vp]
    lib.zo_resolve.restype = None
    lib.zo_render_aov.argtypes = [vp, vp, vp, u32, u32, C.c_int, vp, vp, u32]
    lib.zo_render_aov.restype = None
    lib.zo_resolve_aov.argtypes = [u32, vp, u32, vp]
    lib.zo_resolve_aov.restype = None
    lib.zo_ggx_micro_directional_albedo.argtypes = [C.c_float, C.c_float, u32]
    lib.zo_ggx_micro_directional_albedo.restype = C.c_float
    lib.zo_ggx_f_s_ss.argtypes = [C.c_float, C.c_float, C.c_float, C.c_float, u32]
    lib.zo_ggx_f_s_ss.restype = C.c_float
    lib.zo_ggx_directional_albedo.argtypes = [vp, C.c_float, C.c_float, C.c_float, u32]
    lib.zo_ggx_directional_albedo.restype = C.c_float
    lib.zo_ggx_average_albedo.argtypes = [vp, C.c_float, C.c_float, u32]
    lib.zo_ggx_average_albedo.restype = C.c_float
    lib.zo_ggx_micro_average_albedo.argtypes = [vp, C.c_float, u32]
    lib.zo_ggx_micro_average_albedo.restype = C.c_float
    lib.zo_set_wavefront_light_order.argtypes = [C.c_int]
    lib.zo_set_wavefront_light_order.restype = None
    lib.zo_light_tree_random.argtypes = [vp, vp, vp, vp, C.c_int, C.c_float, C.c_float, vp]
    lib.zo_light_tree_random.restype = u32
    lib.zo_light_tree_pdf.argtypes = [vp, vp, vp, vp, C.c_int, C.c_float, u32]
    lib.zo_light_tree_pdf.restype = C.c_float
    for fn in (lib.zo_image_sample, lib.zo_image_pdf, lib.zo_image_texel):
        fn.argtypes = [vp, u32, u32, vp, vp]
        fn.restype = None
    lib.zo_sobol_stream.argtypes = [u32, u32, u32, u32, vp]
    lib.zo_sobol_stream.restype = None
    lib.zo_sobol_directions.argtypes = [vp]
    lib.zo_sobol_directions.restype = None
    lib._render_bound = True
    return lib


class ZoMesh(C.Structure):
    _fields_ = [("nodes", C.c_void_p), ("triangles", C.c_void_p), ("positions", C.c_void_p), ("normals", C.c_void_p),
                ("uvs", C.c_void_p), ("parts", C.c_void_p)]


def mesh_table(num_meshes):
    """ZoMesh[num_meshes] over the host arrays of the engine's compiled meshes (shape ids 7..), for zo_render."""
    from zyg_b200 import lib as zlib, su

    if 0 == num_meshes:
        return None
    L = zlib.load_library()
    table = (ZoMesh * num_meshes)()
    which = (zlib.MESH_BINARY_NODES, zlib.MESH_TRIANGLES, zlib.MESH_POSITIONS, zlib.MESH_NORMALS, zlib.MESH_UVS,
             zlib.MESH_PARTS)
    for i in range(num_meshes):
        handle = su._su().zyg_su_mesh(7 + i)
        assert handle, f"no mesh registered as shape {7 + i}"
        ptrs = [L.zyg_mesh_data(handle, w, None) for w in which]
        table[i] = ZoMesh(*ptrs)
    return table


def render(scene, view, width, height, iteration, num_samples, per_sample_iterations=True, threads=0, film=None,
           num_meshes=0, wavefront_light_order=False):
    """zo_render over the flattened scene (pointers from zyg_b200.su.compile_scene). Returns the film (H, W, 4).
    wavefront_light_order: take the sampler draws of sampleLights in the device's order (see zyg_oracle.h)."""
    lib = _bind_render(load())
    lib.zo_set_wavefront_light_order(1 if wavefront_light_order else 0)
    if film is None:
        film = np.zeros((height, width, 4), np.float32)
    table = mesh_table(num_meshes)
    lib.zo_render(scene, view, table, iteration, num_samples, 1 if per_sample_iterations else 0, _p(film), threads)
    return film


AOV_CLASSES = ("Albedo", "Depth", "MaterialId", "GeometricNormal", "ShadingNormal", "Roughness", "Emission", "Direct", "Indirect")


def render_aov(scene, view, width, height, iteration, num_samples, aov_slots, threads=0, num_meshes=0, wavefront_light_order=False):
    """zo_render_aov: the film plus one Pack4f layer per class of `aov_slots` (the bit mask the view was compiled with), each cleared to
    the class default first (aov.Buffer.clear). Returns (film, {class index: layer})."""
    lib = _bind_render(load())
    lib.zo_set_wavefront_light_order(1 if wavefront_light_order else 0)
    film = np.zeros((height, width, 4), np.float32)
    layers, table = {}, (C.c_void_p * len(AOV_CLASSES))()
    for c in range(len(AOV_CLASSES)):
        if aov_slots & (1 << c):
            layer = np.zeros((height, width, 4), np.float32)
            if 1 == c:
                layer[..., :3] = np.finfo(np.float32).max
            layers[c] = layer
            table[c] = layer.ctypes.data
    lib.zo_render_aov(scene, view, mesh_table(num_meshes), iteration, num_samples, 1, _p(film), table, threads)
    return film, layers


def render_alpha(scene, view, width, height, iteration, num_samples, threads=0, num_meshes=0, wavefront_light_order=False):
    """zo_render_layers with the Transparent buffer's alpha lane: returns (film, alpha sums (H, W))."""
    lib = _bind_render(load())
    lib.zo_render_layers.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]
    lib.zo_render_layers.restype = None
    lib.zo_set_wavefront_light_order(1 if wavefront_light_order else 0)
    film = np.zeros((height, width, 4), np.float32)
    alpha = np.zeros((height, width), np.float32)
    lib.zo_render_layers(scene, view, mesh_table(num_meshes), iteration, num_samples, 1, _p(film), None, _p(alpha), threads)
    return film, alpha


def resolve_transparent(view, film, alpha):
    lib = _bind_render(load())
    lib.zo_resolve_transparent.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]
    lib.zo_resolve_transparent.restype = None
    out = np.empty_like(film)
    lib.zo_resolve_transparent(view, _p(film), _p(alpha), film.shape[0] * film.shape[1], _p(out))
    return out


def denoise(view, film, normal_layer, albedo_layer, sigma):
    lib = _bind_render(load())
    lib.zo_denoise.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p]
    lib.zo_denoise.restype = None
    out = np.empty_like(film)
    lib.zo_denoise(view, _p(film), _p(normal_layer), _p(albedo_layer), sigma, _p(out))
    return out


def resolve_aov(aov_class, layer):
    out = np.empty_like(layer)
    _bind_render(load()).zo_resolve_aov(aov_class, _p(layer), layer.shape[0] * layer.shape[1], _p(out))
    return out


def image_sample(scene, index, r2):
    r2 = np.ascontiguousarray(r2, np.float32)
    out = np.empty((r2.shape[0], 3), np.float32)
    _bind_render(load()).zo_image_sample(scene, index, r2.shape[0], _p(r2), _p(out))
    return out


def image_pdf(scene, index, uv):
    uv = np.ascontiguousarray(uv, np.float32)
    out = np.empty(uv.shape[0], np.float32)
    _bind_render(load()).zo_image_pdf(scene, index, uv.shape[0], _p(uv), _p(out))
    return out


def image_texel(scene, index, uvr):
    uvr = np.ascontiguousarray(uvr, np.float32)
    out = np.empty((uvr.shape[0], 3), np.float32)
    _bind_render(load()).zo_image_texel(scene, index, uvr.shape[0], _p(uvr), _p(out))
    return out


def resolve(view, film):
    lib = _bind_render(load())
    out = np.empty_like(film)
    lib.zo_resolve(view, _p(film), film.shape[0] * film.shape[1], _p(out))
    return out


def ggx_micro_directional_albedo(alpha, n_dot_wo, num_samples=1024):
    return _bind_render(load()).zo_ggx_micro_directional_albedo(alpha, n_dot_wo, num_samples)


def ggx_directional_albedo(luts, alpha, f0, n_dot_wo, num_samples=1024):
    return _bind_render(load()).zo_ggx_directional_albedo(_p(luts), alpha, f0, n_dot_wo, num_samples)


def ggx_micro_average_albedo(luts, alpha, num_samples=1024):
    return _bind_render(load()).zo_ggx_micro_average_albedo(_p(luts), alpha, num_samples)


def ggx_average_albedo(luts, alpha, f0, num_samples=1024):
    return _bind_render(load()).zo_ggx_average_albedo(_p(luts), alpha, f0, num_samples)


def ggx_f_s_ss(alpha, f0, ior_t, n_dot_wo, num_samples=1024):
    return _bind_render(load()).zo_ggx_f_s_ss(alpha, f0, ior_t, n_dot_wo, num_samples)


def light_tree_random(scene, view, p, n, random, split_threshold, total_sphere=False):
    """Tree.randomLight: list of (light id, pdf)."""
    lib = _bind_render(load())
    p, n = np.asarray(p, np.float32), np.asarray(n, np.float32)
    picks = np.zeros(128, np.float32)
    num = lib.zo_light_tree_random(scene, view, _p(p), _p(n), int(total_sphere), random, split_threshold, _p(picks))
    return [(int(picks[2 * i]), float(picks[2 * i + 1])) for i in range(num)]


def light_tree_pdf(scene, view, p, n, split_threshold, light, total_sphere=False):
    lib = _bind_render(load())
    p, n = np.asarray(p, np.float32), np.asarray(n, np.float32)
    return float(lib.zo_light_tree_pdf(scene, view, _p(p), _p(n), int(total_sphere), split_threshold, light))


def sobol_stream(sample, seed, n, pad_every=0):
    out = np.empty(n, np.float32)
    _bind_render(load()).zo_sobol_stream(sample, seed, n, pad_every, _p(out))
    return out


def sobol_directions():
    out = np.empty((5, 32), np.uint32)
    _bind_render(load()).zo_sobol_directions(_p(out))
    return out


# ---- the oracle's own scene-compile builders (oracle/builders.cpp) -----------------------------------

def _bind_builders(lib):
    if getattr(lib, "_builders_bound", False):
        return lib
    vp, u32 = C.c_void_p, C.c_uint32
    lib.zo_build_free.argtypes = [vp]
    lib.zo_build_free.restype = None
    lib.zo_mesh_build.argtypes = [u32, vp, u32, vp, u32, vp, u32, vp, u32, vp, u32, u32]
    lib.zo_mesh_build.restype = vp
    lib.zo_prop_tree_build.argtypes = [vp, u32, vp, u32]
    lib.zo_prop_tree_build.restype = vp
    lib.zo_light_tree_build.argtypes = [u32, vp, vp, vp, vp]
    lib.zo_light_tree_build.restype = vp
    lib.zo_mesh_sampler_build.argtypes = [vp, u32, u32, u32, C.c_int]
    lib.zo_mesh_sampler_build.restype = vp
    for fn in (lib.zo_mesh_data, lib.zo_prop_tree_data, lib.zo_mesh_sampler_data):
        fn.argtypes = [vp, C.c_int, C.POINTER(C.c_uint64)]
        fn.restype = vp
    lib.zo_light_tree_data.argtypes = [vp, C.c_int, C.c_int, C.POINTER(C.c_uint64)]
    lib.zo_light_tree_data.restype = vp
    lib._builders_bound = True
    return lib


def _blob(fn, handle, *which):
    n = C.c_uint64()
    p = fn(handle, *which, C.byref(n))
    if not p or 0 == n.value:
        return b""
    return C.string_at(p, n.value)


MESH_NODE_DTYPE = np.dtype([("min", "<f4", 3), ("min_data", "<u4"), ("max", "<f4", 3), ("max_data", "<u4")])


class BuiltMesh:
    """A triangle tree built by the oracle's own builder (zo_mesh_build): same arrays, same numbering as zyg_mesh_data."""

    NODES, TRIANGLES, ORIGINAL, POSITIONS, NORMALS, UVS, PARTS = range(7)
    _dtypes = {0: MESH_NODE_DTYPE, 1: "<u4", 2: "<u4", 3: "<f4", 4: "<u2", 5: "<f4", 6: "<u2"}

    def __init__(self, positions, indices, normals=None, uvs=None, parts=None, threads=0):
        lib = _bind_builders(load())
        positions = np.ascontiguousarray(positions, np.float32)
        indices = np.ascontiguousarray(indices, np.uint32).reshape(-1)
        normals = None if normals is None else np.ascontiguousarray(normals, np.float32)
        uvs = None if uvs is None else np.ascontiguousarray(uvs, np.float32)
        parts = None if parts is None else np.ascontiguousarray(parts, np.uint32).reshape(-1)
        self._lib = lib
        self.handle = lib.zo_mesh_build(0 if parts is None else parts.size // 3, None if parts is None else _p(parts),
                                        indices.size // 3, _p(indices), positions.shape[0], _p(positions), 3,
                                        None if normals is None else _p(normals), 3, None if uvs is None else _p(uvs), 2, threads)
        self._cache = {}

    def raw(self, which) -> bytes:
        return _blob(self._lib.zo_mesh_data, self.handle, which)

    def data(self, which) -> np.ndarray:
        if which not in self._cache:
            self._cache[which] = np.frombuffer(self.raw(which), self._dtypes[which]).copy()
        return self._cache[which]

    def diagnostics(self):
        d = np.frombuffer(self.raw(100), np.uint32)
        return {"leaf_offset_mismatches": int(d[0]), "unsplittable": int(d[1]), "task_root_leaves": int(d[2])}

    def table(self):
        """ZoMesh[1] over this mesh's arrays, for zo_render / zo_mesh_sampler_build."""
        arrays = [self.data(w) for w in (self.NODES, self.TRIANGLES, self.POSITIONS, self.NORMALS, self.UVS, self.PARTS)]
        self._keep = arrays
        t = (ZoMesh * 1)()
        t[0] = ZoMesh(*[_p(a) for a in arrays])
        return t

    def __del__(self):
        try:
            self._lib.zo_build_free(self.handle)
        except Exception:
            pass


def build_prop_tree(indices, aabbs, threads=0):
    """PropBvhBuilder.build over prop ids `indices` and the (N, 32-byte) world boxes: (nodes bytes, indices bytes)."""
    lib = _bind_builders(load())
    indices = np.ascontiguousarray(indices, np.uint32)
    aabbs = np.ascontiguousarray(aabbs)
    h = lib.zo_prop_tree_build(_p(indices), indices.size, _p(aabbs), threads)
    out = _blob(lib.zo_prop_tree_data, h, 0), _blob(lib.zo_prop_tree_data, h, 1)
    lib.zo_build_free(h)
    return out


LIGHT_TREE_PARTS = ("nodes", "middles", "orders", "mapping", "infinite_cdf", "floats", "uints")


def build_light_tree(light_aabbs, light_cones, two_sided, finite):
    """Builder.build over the scene lights: dict of raw byte strings per LIGHT_TREE_PARTS."""
    lib = _bind_builders(load())
    light_aabbs = np.ascontiguousarray(light_aabbs)
    light_cones = np.ascontiguousarray(light_cones, np.float32)
    two_sided = np.ascontiguousarray(two_sided, np.uint8)
    finite = np.ascontiguousarray(finite, np.uint8)
    h = lib.zo_light_tree_build(two_sided.size, _p(light_aabbs), _p(light_cones), _p(two_sided), _p(finite))
    out = {name: _blob(lib.zo_light_tree_data, h, 0, i) for i, name in enumerate(LIGHT_TREE_PARTS)}
    lib.zo_build_free(h)
    return out


def build_mesh_sampler(table, num_tree_triangles, num_parts, part, two_sided):
    """Part.configure + Builder.buildPrimitive over a ZoMesh: dict with the sampler tables and its primitive tree."""
    lib = _bind_builders(load())
    h = lib.zo_mesh_sampler_build(table, num_tree_triangles, num_parts, part, int(two_sided))
    out = {name: _blob(lib.zo_mesh_sampler_data, h, i)
           for i, name in enumerate(("triangle_mapping", "triangle_pdfs", "primitive_mapping", "part_areas", "floats"))}
    out["tree"] = {name: _blob(lib.zo_light_tree_data, h, 1, i) for i, name in enumerate(LIGHT_TREE_PARTS)}
    lib.zo_build_free(h)
    return out
