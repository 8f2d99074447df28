"""ctypes bindings for ``libzyg_b200.so`` (C ABI declared in ``include/zygpu.h``)."""

from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ZYG_B200_LIB") or os.path.join(_HERE, "libzyg_b200.so")  # the override is for tuning builds

RAY_DTYPE = np.dtype([("origin", "<f4", 3), ("min_t", "<f4"), ("direction", "<f4", 3), ("max_t", "<f4")])
HIT_DTYPE = np.dtype([("t", "<f4"), ("u", "<f4"), ("v", "<f4"), ("primitive", "<u4")])
assert RAY_DTYPE.itemsize == 32 and HIT_DTYPE.itemsize == 16

RAY_MAX_T = np.float32(2.14748313e9)  # src/core/scene/ray_offset.zig:5

CLOSEST, ANY, CLOSEST_BINARY, ANY_BINARY = 0, 1, 2, 3

(MESH_BINARY_NODES, MESH_TRIANGLES, MESH_ORIGINAL, MESH_POSITIONS, MESH_NORMALS, MESH_UVS, MESH_PARTS,
 MESH_WIDE_NODES, MESH_WIDE_TRIS) = range(9)


class MeshInfo(C.Structure):
    _fields_ = [
        ("num_source_triangles", C.c_uint32),
        ("num_tree_triangles", C.c_uint32),
        ("num_vertices", C.c_uint32),
        ("num_binary_nodes", C.c_uint32),
        ("num_wide_nodes", C.c_uint32),
        ("wide_max_depth", C.c_uint32),
        ("num_degenerate_leaves", C.c_uint32),
        ("num_leaf_order_fixups", C.c_uint32),
        ("aabb_min", C.c_float * 3),
        ("aabb_max", C.c_float * 3),
    ]


class TraceCounters(C.Structure):
    _fields_ = [("nodes", C.c_uint64), ("triangles", C.c_uint64), ("rays", C.c_uint64), ("max_stack", C.c_uint64)]


_lib = None


def load_library() -> C.CDLL:
    """Load the shared library; raise loudly when it is missing (no CPU fallback exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C zyg_b200/csrc`). zyg_b200 has no CPU fallback."
        )
    lib = C.CDLL(LIB_PATH)
    u32p, f32p, vp = C.POINTER(C.c_uint32), C.POINTER(C.c_float), C.c_void_p

    lib.zygpu_last_error.restype = C.c_char_p
    lib.zyg_mesh_build.argtypes = [C.c_uint32, u32p, C.c_uint32, u32p, C.c_uint32, f32p, C.c_uint32, f32p, C.c_uint32,
                                   f32p, C.c_uint32, C.c_uint32, C.POINTER(vp)]
    lib.zyg_mesh_free.argtypes = [vp]
    lib.zyg_mesh_free.restype = None
    lib.zyg_mesh_data.argtypes = [vp, C.c_int, C.POINTER(C.c_uint64)]
    lib.zyg_mesh_data.restype = vp
    lib.zyg_mesh_info.argtypes = [vp, C.POINTER(MeshInfo)]
    lib.zygpu_create.argtypes = [C.c_int, C.POINTER(vp)]
    lib.zygpu_destroy.argtypes = [vp]
    lib.zygpu_destroy.restype = None
    lib.zygpu_upload_mesh.argtypes = [vp, vp]
    lib.zygpu_mesh_build.argtypes = [vp, C.c_uint32, u32p, C.c_uint32, u32p, C.c_uint32, f32p, C.c_uint32, f32p, C.c_uint32,
                                     f32p, C.c_uint32, C.POINTER(vp), C.POINTER(C.c_float)]
    lib.zygpu_mesh_refit.argtypes = [vp, vp, f32p, C.c_uint32, f32p, C.c_uint32, C.POINTER(C.c_float)]
    lib.zygpu_trace_batch.argtypes = [vp, C.c_int, C.c_int, vp, C.c_uint64, vp]
    lib.zygpu_trace_batch_device.argtypes = [vp, C.c_int, C.c_int, vp, C.c_uint64, vp, vp, C.POINTER(TraceCounters)]
    _lib = lib
    return lib


def _check(rc: int, what: str) -> int:
    if rc < 0:
        raise RuntimeError(f"{what} failed: {load_library().zygpu_last_error().decode()}")
    return rc


def _ptr(a: np.ndarray | None, ctype):
    return None if a is None else a.ctypes.data_as(C.POINTER(ctype))


class Mesh:
    """A compiled triangle mesh (``zyg_mesh``): reference-order binary BVH + device layout."""

    _DTYPES = {
        MESH_BINARY_NODES: np.dtype([("min", "<f4", 3), ("min_data", "<u4"), ("max", "<f4", 3), ("max_data", "<u4")]),
        MESH_TRIANGLES: np.dtype("<u4"),
        MESH_ORIGINAL: np.dtype("<u4"),
        MESH_POSITIONS: np.dtype("<f4"),
        MESH_NORMALS: np.dtype("<u2"),
        MESH_UVS: np.dtype("<f4"),
        MESH_PARTS: np.dtype("<u2"),
        MESH_WIDE_NODES: np.dtype("u1"),
        MESH_WIDE_TRIS: np.dtype([("a", "<f4", 3), ("primitive", "<u4"), ("e1", "<f4", 3), ("leaf_min_x", "<f4"),
                                  ("e2", "<f4", 3), ("leaf_min_y", "<f4"), ("leaf_min_z", "<f4"),
                                  ("leaf_max", "<f4", 3)]),
    }

    def __init__(self, positions, indices=None, normals=None, uvs=None, parts=None, num_threads: int = 0, device=None):
        """`device`: a Device -> the tree is built on the GPU (zygpu_mesh_build, an LBVH) instead of the host's reference-order
        SAH build; `build_ms` then holds the CUDA-event time of the build."""
        lib = load_library()
        self._keep = []

        def f32(a):
            if a is None:
                return None
            a = np.ascontiguousarray(a, dtype=np.float32)
            self._keep.append(a)
            return a

        positions = f32(positions).reshape(-1, 3)
        normals = f32(normals)
        uvs = f32(uvs)
        if indices is not None:
            indices = np.ascontiguousarray(indices, dtype=np.uint32).reshape(-1)
            num_triangles = indices.size // 3
        else:
            num_triangles = positions.shape[0] // 3
        if parts is not None:
            parts = np.ascontiguousarray(parts, dtype=np.uint32).reshape(-1)
        num_parts = 0 if parts is None else parts.size // 3

        handle = C.c_void_p()
        self.build_ms = None
        if device is not None:
            ms = C.c_float()
            _check(lib.zygpu_mesh_build(device.handle, num_parts, _ptr(parts, C.c_uint32), num_triangles, _ptr(indices, C.c_uint32),
                                        positions.shape[0], _ptr(positions, C.c_float), 3, _ptr(normals, C.c_float), 3,
                                        _ptr(uvs, C.c_float), 2, C.byref(handle), C.byref(ms)), "zygpu_mesh_build")
            self.build_ms = ms.value
        else:
            _check(lib.zyg_mesh_build(num_parts, _ptr(parts, C.c_uint32), num_triangles, _ptr(indices, C.c_uint32),
                                      positions.shape[0], _ptr(positions, C.c_float), 3, _ptr(normals, C.c_float), 3,
                                      _ptr(uvs, C.c_float), 2, num_threads, C.byref(handle)), "zyg_mesh_build")
        self.handle = handle
        self._lib = lib

    def __del__(self):
        if getattr(self, "handle", None):
            self._lib.zyg_mesh_free(self.handle)
            self.handle = None

    def info(self) -> MeshInfo:
        info = MeshInfo()
        _check(self._lib.zyg_mesh_info(self.handle, C.byref(info)), "zyg_mesh_info")
        return info

    def data(self, which: int) -> np.ndarray:
        """Copy of one of the compiled arrays as a numpy array."""
        n = C.c_uint64()
        p = self._lib.zyg_mesh_data(self.handle, which, C.byref(n))
        if not p:
            raise ValueError(f"no mesh array {which}")
        buf = (C.c_uint8 * n.value).from_address(p)
        return np.frombuffer(buf, dtype=self._DTYPES[which]).copy()


class Device:
    """One GPU (``zygpu_device``). Raises if CUDA or an sm_100 device is unavailable."""

    def __init__(self, ordinal: int = 0):
        lib = load_library()
        handle = C.c_void_p()
        _check(lib.zygpu_create(ordinal, C.byref(handle)), "zygpu_create")
        self.handle = handle
        self._lib = lib

    def close(self):
        if getattr(self, "handle", None):
            self._lib.zygpu_destroy(self.handle)
            self.handle = None

    __del__ = close

    def upload_mesh(self, mesh: Mesh) -> int:
        return _check(self._lib.zygpu_upload_mesh(self.handle, mesh.handle), "zygpu_upload_mesh")

    def refit_mesh(self, mesh: Mesh, positions, normals=None) -> float:
        """Moved vertices, same topology (``zygpu_mesh_refit``); returns the CUDA-event time in ms."""
        positions = np.ascontiguousarray(positions, np.float32).reshape(-1, 3)
        normals = None if normals is None else np.ascontiguousarray(normals, np.float32)
        ms = C.c_float()
        _check(self._lib.zygpu_mesh_refit(self.handle, mesh.handle, _ptr(positions, C.c_float), 3, _ptr(normals, C.c_float), 3,
                                          C.byref(ms)), "zygpu_mesh_refit")
        return ms.value

    def trace_batch(self, mesh_id: int, mode: int, rays: np.ndarray, out: np.ndarray | None = None) -> np.ndarray:
        """Host buffers in, host buffers out (``zygpu_trace_batch``)."""
        assert rays.dtype == RAY_DTYPE and rays.flags.c_contiguous
        n = rays.shape[0]
        if out is None:
            out = np.empty(n, dtype=np.uint32 if mode in (ANY, ANY_BINARY) else HIT_DTYPE)
        _check(self._lib.zygpu_trace_batch(self.handle, mesh_id, mode, rays.ctypes.data, n, out.ctypes.data),
               "zygpu_trace_batch")
        return out

    def trace_batch_ptr(self, mesh_id: int, mode: int, rays_ptr: int, n: int, out_ptr: int, host: bool,
                        stream: int = 0, counters: TraceCounters | None = None) -> None:
        """Raw-pointer variant: pinned host pointers (host=True) or device pointers on ``stream``."""
        if host:
            _check(self._lib.zygpu_trace_batch(self.handle, mesh_id, mode, rays_ptr, n, out_ptr), "zygpu_trace_batch")
        else:
            _check(self._lib.zygpu_trace_batch_device(self.handle, mesh_id, mode, rays_ptr, n, out_ptr, stream,
                                                      C.byref(counters) if counters is not None else None),
                   "zygpu_trace_batch_device")
