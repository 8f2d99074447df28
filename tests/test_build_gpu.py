"""Device BVH build and refit (builder_base.zig:65-283 + triangle_tree_builder.zig:33-207 moved to the GPU as an
LBVH). The tree differs from the reference's SAH tree by design, so the bar is traversal equality:

* the oracle walking the device-built binary tree equals the device's order-exact kernel byte for byte, and the wide kernel agrees
  with it (bit-equal t, same triangle up to equal-t ties) - the same contract as a host-built mesh (test_trace_gpu.check_closest);
* device-built and host-built trees report the same hits for the same rays in terms of the caller's triangle ids, except where
  the leaf-box gate decides (a hit whose ray misses the box of its leaf is dropped by whichever tree has that leaf):
  a bounded handful per million rays;
* a render of a scene whose meshes were built on the device matches the oracle's render of the same compiled scene per pixel.
"""

import numpy as np
import pytest

import oracle_lib as oracle
from test_trace_gpu import arrays, check_closest
from zyg_b200 import lib, scenes, su

pytestmark = pytest.mark.gpu
NULL = 0xFFFFFFFF


def sphere(quads=(200, 100), seed=None):
    kw = {} if seed is None else {"seed": seed}
    return scenes.displaced_sphere(*quads, **kw)


def rays_for_tests(n=1 << 16):
    return np.concatenate([scenes.primary_rays(192, 192), scenes.random_rays(n)])


def leaf_ranges(nodes):
    """(start, count) of every leaf reachable from the root of a 32-byte reference-layout tree."""
    out, stack = [], [0]
    while stack:
        n = stack.pop()
        if nodes["max_data"][n]:
            out.append((int(nodes["min_data"][n]), int(nodes["max_data"][n])))
        else:
            stack += [int(nodes["min_data"][n]), int(nodes["min_data"][n]) + 1]
    return out


def test_device_built_tree_structure(device):
    positions, normals, uvs, indices = sphere((64, 32))
    mesh = lib.Mesh(positions, indices, normals, uvs, device=device)
    info = mesh.info()
    n = indices.size // 3
    assert info.num_source_triangles == n == info.num_tree_triangles  # no spatial splits: every triangle once
    original = mesh.data(lib.MESH_ORIGINAL)
    assert np.array_equal(np.sort(original), np.arange(n))
    tris = mesh.data(lib.MESH_TRIANGLES).reshape(-1, 3)
    assert np.array_equal(tris, indices.reshape(-1, 3)[original])
    nodes = mesh.data(lib.MESH_BINARY_NODES)
    ranges = sorted(leaf_ranges(nodes))
    assert ranges[0][0] == 0 and all(1 <= c <= 3 for _, c in ranges)
    assert all(a + c == b for (a, c), (b, _) in zip(ranges, ranges[1:])) and ranges[-1][0] + ranges[-1][1] == n
    # boxes: root = mesh bounds, every leaf box = the box of its triangles
    p = positions.reshape(-1, 3)
    used = p[indices.reshape(-1)]
    assert np.array_equal(nodes["min"][0], used.min(0)) and np.array_equal(nodes["max"][0], used.max(0))
    wide = mesh.data(lib.MESH_WIDE_TRIS)
    assert np.array_equal(np.sort(wide["primitive"]), np.arange(n))
    assert mesh.build_ms is not None and mesh.build_ms > 0


def test_device_built_mesh_matches_oracle(device):
    positions, normals, uvs, indices = sphere()
    mesh = lib.Mesh(positions, indices, normals, uvs, device=device)
    check_closest(device, mesh, rays_for_tests(), want_hits=20000)
    nodes, tris, pos, _ = arrays(mesh)
    shadow = scenes.random_rays(1 << 16, shadow=True)
    mid = device.upload_mesh(mesh)
    want = oracle.trace_any(nodes, tris, pos, shadow)
    assert np.array_equal(device.trace_batch(mid, lib.ANY_BINARY, shadow), want)
    assert np.array_equal(device.trace_batch(mid, lib.ANY, shadow), want)


def compare_trees(device, mesh_a, mesh_b, rays):
    """Hits of two trees over the same triangles, by the caller's triangle id. Returns the number of rays that differ."""
    out = []
    for m in (mesh_a, mesh_b):
        hits = device.trace_batch(device.upload_mesh(m), lib.CLOSEST, rays)
        original = m.data(lib.MESH_ORIGINAL)
        tri = np.where(hits["primitive"] != NULL, original[np.minimum(hits["primitive"], original.size - 1)], NULL)
        out.append((hits["t"].view(np.uint32), tri))
    (ta, ia), (tb, ib) = out
    differ = ta != tb
    # same t but another triangle: an equal-t tie (shared edge), not a difference of the trees
    return int(differ.sum()), int(((ia != ib) & ~differ).sum())


def test_device_tree_agrees_with_host_tree(device, sphere_mesh):
    host, positions, indices = sphere_mesh
    _, normals, uvs, _ = sphere()
    built = lib.Mesh(positions, indices, normals, uvs, device=device)
    rays = rays_for_tests(1 << 18)
    differ, ties = compare_trees(device, host, built, rays)
    assert differ <= max(4, rays.shape[0] // 50000), f"{differ} of {rays.shape[0]} rays see another hit (leaf-box gate cases only)"
    assert ties < rays.shape[0] // 100


def test_million_triangle_build_is_fast(device):
    positions, normals, uvs, indices = sphere((1000, 500))
    mesh = lib.Mesh(positions, indices, normals, uvs, device=device)
    mesh = lib.Mesh(positions, indices, normals, uvs, device=device)  # second build: allocator warm
    assert mesh.info().num_source_triangles == 1000000
    assert mesh.build_ms < 100.0, f"device build of 1M triangles took {mesh.build_ms:.1f} ms"
    rays = scenes.primary_rays(256, 256)
    nodes, tris, pos, _ = arrays(mesh)
    ref = oracle.trace_closest(nodes, tris, pos, rays)
    got = device.trace_batch(device.upload_mesh(mesh), lib.CLOSEST, rays)
    assert np.array_equal(got["t"].view(np.uint32), ref["t"].view(np.uint32))


@pytest.mark.parametrize("builder", ["host", "device"])
def test_refit_after_vertices_moved(device, builder):
    positions, normals, uvs, indices = sphere((160, 80))
    mesh = lib.Mesh(positions, indices, normals, uvs, device=device if builder == "device" else None)
    p = positions.reshape(-1, 3)
    rng = np.random.default_rng(5)
    moved = (p * (1.0 + 0.08 * np.sin(7.0 * p[:, 1:2] + 2.0)) + 0.01 * rng.standard_normal(p.shape)).astype(np.float32)
    ms = device.refit_mesh(mesh, moved)
    assert ms > 0
    assert np.array_equal(mesh.data(lib.MESH_POSITIONS)[:-1].reshape(-1, 3), moved)
    # the refitted trees are valid trees over the moved triangles: oracle == order-exact kernel, wide agrees
    check_closest(device, mesh, rays_for_tests(), want_hits=10000)
    # and they see what a tree built from scratch over the moved vertices sees
    fresh = lib.Mesh(moved, indices, normals, uvs)
    rays = rays_for_tests(1 << 17)
    differ, _ = compare_trees(device, fresh, mesh, rays)
    assert differ <= max(4, rays.shape[0] // 50000)


def download_film(width, height):
    import ctypes as C

    L = lib.load_library()
    L.zygpu_download_film.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
    film = np.zeros((height, width, 4), np.float32)
    assert 0 == L.zygpu_download_film(su.device_handle(), film.ctypes.data, width * height)
    return film


def test_render_with_device_built_meshes_matches_oracle(monkeypatch):
    su.release()
    monkeypatch.setattr(su, "MESH_BUILDER", su.DEVICE_BUILDER)
    w, spp = 96, 4
    scenes.sphere_scene(w, w, spp=spp, quads=(100, 50))
    scene, view = su.compile_scene()
    ref = oracle.render(scene, view, w, w, 0, spp, num_meshes=1)
    su.render_frame(0)
    gpu = download_film(w, w)
    su.release()
    rel = np.abs(gpu[..., :3] - ref[..., :3]).sum(-1) / np.maximum(np.abs(ref[..., :3]).sum(-1), 1e-6)
    assert np.array_equal(gpu[..., 3], ref[..., 3])
    assert np.median(rel) < 2e-6 and (rel > 1e-3).mean() < 5e-3
