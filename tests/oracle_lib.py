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
