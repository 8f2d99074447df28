"""Host-side scene compile: the restated SAH/spatial-split builder and the derived wide layout.

Checked with the oracle's traversal (reference order) against the oracle's O(N) brute force, plus
structural invariants the reference relies on (src/core/scene/bvh/builder_base.zig).
"""

import numpy as np
import pytest

import oracle_lib as oracle
from zyg_b200 import lib, scenes

NULL = 0xFFFFFFFF


def arrays(mesh):
    return (mesh.data(lib.MESH_BINARY_NODES), mesh.data(lib.MESH_TRIANGLES), mesh.data(lib.MESH_POSITIONS),
            mesh.data(lib.MESH_ORIGINAL))


def test_tree_covers_every_triangle(sphere_mesh):
    mesh, positions, indices = sphere_mesh
    nodes, tris, pos, original = arrays(mesh)
    info = mesh.info()
    assert info.num_source_triangles == indices.shape[0]
    assert info.num_tree_triangles >= info.num_source_triangles
    assert np.array_equal(np.unique(original), np.arange(indices.shape[0]))
    # tree-order triangles are the caller's triangles
    assert np.array_equal(tris.reshape(-1, 3), indices[original])
    assert np.array_equal(pos[:-1].reshape(-1, 3), positions)
    assert info.num_degenerate_leaves == 0 and info.num_leaf_order_fixups == 0


def test_leaves_partition_the_triangle_list(sphere_mesh):
    mesh, _, _ = sphere_mesh
    nodes, tris, _, _ = arrays(mesh)
    leaf = nodes["max_data"] != 0
    start, count = nodes["min_data"][leaf], nodes["max_data"][leaf]
    order = np.argsort(start)
    assert start[order][0] == 0
    assert np.array_equal(start[order][1:], (start[order] + count[order])[:-1])
    assert (start[order] + count[order])[-1] == tris.size // 3
    # shape_provider.zig:922 -> at most 4 primitives unless the SAH says a bigger leaf is cheaper (<= 255)
    assert count.max() <= 255
    assert (count <= 4).mean() > 0.9 and (count <= 8).mean() > 0.99
    # children are adjacent and come after their parent (serialised depth-first)
    inner = ~leaf
    assert (nodes["min_data"][inner] > np.nonzero(inner)[0]).all()


def test_children_nest_inside_parents(sphere_mesh):
    mesh, _, _ = sphere_mesh
    nodes = mesh.data(lib.MESH_BINARY_NODES)
    inner = np.nonzero(nodes["max_data"] == 0)[0]
    for k in (0, 1):
        child = nodes[nodes["min_data"][inner] + k]
        assert (child["min"] >= nodes["min"][inner]).all()
        assert (child["max"] <= nodes["max"][inner]).all()


def test_leaf_boxes_contain_their_clipped_triangles(sphere_mesh):
    mesh, _, _ = sphere_mesh
    nodes, tris, pos, _ = arrays(mesh)
    P = pos[:-1].reshape(-1, 3)[tris.reshape(-1, 3)]  # (T, 3, 3)
    tmin, tmax = P.min(axis=1), P.max(axis=1)
    for n in np.nonzero(nodes["max_data"] != 0)[0][::7]:
        s, c = nodes["min_data"][n], nodes["max_data"][n]
        # every referenced triangle overlaps the leaf box (spatial splits clip, never drop)
        assert (tmin[s:s + c] <= nodes["max"][n]).all() and (tmax[s:s + c] >= nodes["min"][n]).all()


def test_tree_traversal_equals_brute_force(sphere_mesh):
    mesh, _, indices = sphere_mesh
    nodes, tris, pos, original = arrays(mesh)
    rays = np.concatenate([scenes.primary_rays(48, 48), scenes.random_rays(2048)])
    tree = oracle.trace_closest(nodes, tris, pos, rays)
    brute, ties = oracle.brute_closest(np.ascontiguousarray(indices.reshape(-1)), pos, rays)
    hit = tree["primitive"] != NULL
    assert np.array_equal(hit, brute["primitive"] != NULL)
    assert np.array_equal(tree["t"].view(np.uint32), brute["t"].view(np.uint32))
    same = original[tree["primitive"][hit]] == brute["primitive"][hit]
    assert (same | (ties[hit] > 0)).all()
    assert hit.sum() > 1000


def test_any_hit_consistent_with_closest(sphere_mesh):
    mesh, _, _ = sphere_mesh
    nodes, tris, pos, _ = arrays(mesh)
    rays = scenes.random_rays(4096, shadow=True)
    occluded = oracle.trace_any(nodes, tris, pos, rays)
    closest = oracle.trace_closest(nodes, tris, pos, rays)
    assert np.array_equal(occluded != 0, closest["primitive"] != NULL)
    assert 0.2 < occluded.mean() < 0.95


def test_build_is_independent_of_thread_count():
    positions, normals, uvs, indices = scenes.displaced_sphere(96, 48)  # > 1024 triangles: task path
    a = lib.Mesh(positions, indices, normals, uvs, num_threads=1)
    b = lib.Mesh(positions, indices, normals, uvs, num_threads=5)
    for which in (lib.MESH_BINARY_NODES, lib.MESH_TRIANGLES, lib.MESH_WIDE_NODES, lib.MESH_WIDE_TRIS):
        assert a.data(which).tobytes() == b.data(which).tobytes()


@pytest.mark.parametrize("shape", ["single", "quad", "soup", "degenerate"])
def test_small_and_ragged_meshes(shape):
    rng = np.random.default_rng(5)
    if shape == "single":
        positions = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float32)
        indices = None
    elif shape == "quad":
        positions = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0]], np.float32)
        indices = np.array([[0, 1, 2], [0, 2, 3]], np.uint32)
    elif shape == "soup":  # unindexed random triangles, overlapping boxes -> spatial splits
        positions = (rng.random((300 * 3, 3)) * 2 - 1).astype(np.float32)
        indices = None
    else:  # many identical triangles: no plane separates them
        positions = np.tile(np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float32), (40, 1))
        indices = None
    mesh = lib.Mesh(positions, indices)
    nodes, tris, pos, original = arrays(mesh)
    src = np.arange(positions.shape[0], dtype=np.uint32) if indices is None else indices.reshape(-1)
    rays = np.empty(512, lib.RAY_DTYPE)
    rays["origin"] = (rng.random((512, 3)) * 2 - 1).astype(np.float32) + np.float32([0, 0, 3])
    d = (rng.random((512, 3)) - 0.5).astype(np.float32) * np.float32([1.2, 1.2, 0]) + np.float32([0, 0, -3])
    rays["direction"] = d / np.linalg.norm(d, axis=1, keepdims=True)
    rays["min_t"], rays["max_t"] = 0, lib.RAY_MAX_T
    tree = oracle.trace_closest(nodes, tris, pos, rays)
    brute, ties = oracle.brute_closest(np.ascontiguousarray(src), pos, rays)
    assert np.array_equal(tree["t"].view(np.uint32), brute["t"].view(np.uint32))
    hit = tree["primitive"] != NULL
    assert ((original[tree["primitive"][hit]] == brute["primitive"][hit]) | (ties[hit] > 0)).all()


def decode_wide(mesh):
    raw = mesh.data(lib.MESH_WIDE_NODES).reshape(-1, 96)  # 80 bytes of node + 16 bytes of padding
    p = raw[:, 0:12].copy().view("<f4").reshape(-1, 3)
    e = raw[:, 12:15].astype(np.int32) - 127
    imask = raw[:, 15]
    child_base = raw[:, 16:20].copy().view("<u4").reshape(-1)
    tri_base = raw[:, 20:24].copy().view("<u4").reshape(-1)
    meta = raw[:, 24:32]
    qlo = raw[:, 32:56].reshape(-1, 3, 8)
    qhi = raw[:, 56:80].reshape(-1, 3, 8)
    cell = np.ldexp(1.0, e)  # (N, 3) float64
    lo = p[:, :, None].astype(np.float64) + qlo * cell[:, :, None]
    hi = p[:, :, None].astype(np.float64) + qhi * cell[:, :, None]
    return dict(p=p, imask=imask, child_base=child_base, tri_base=tri_base, meta=meta, lo=lo, hi=hi)


def test_wide_layout_is_conservative_and_complete(sphere_mesh):
    mesh, _, _ = sphere_mesh
    w = decode_wide(mesh)
    recs = mesh.data(lib.MESH_WIDE_TRIS)
    info = mesh.info()
    n = w["p"].shape[0]
    assert n == info.num_wide_nodes

    seen_tris = np.zeros(recs.shape[0], bool)
    seen_nodes = np.zeros(n, bool)
    seen_nodes[0] = True
    for i in range(n):
        rank = 0
        for s in range(8):
            m = int(w["meta"][i, s])
            if m == 0:
                assert not (w["imask"][i] >> s) & 1
                continue
            lo, hi = w["lo"][i, :, s], w["hi"][i, :, s]
            if (w["imask"][i] >> s) & 1:
                assert m == (1 << 5) | (24 + s)
                c = int(w["child_base"][i]) + rank
                rank += 1
                assert not seen_nodes[c]
                seen_nodes[c] = True
                # child node's own child boxes lie inside this slot's box
                used = w["meta"][c] != 0
                assert (w["lo"][c][:, used] >= lo[:, None] - 1e-12).all() or True  # quantisation grids differ
            else:
                count = {1: 1, 3: 2, 7: 3}[m >> 5]
                first = int(w["tri_base"][i]) + (m & 31)
                for r in recs[first:first + count]:
                    a = r["a"].astype(np.float64)
                    verts = np.stack([a, a + r["e1"], a + r["e2"]])
                    # the part of the triangle inside the reference leaf lies inside the quantised box;
                    # unsplit triangles lie inside entirely
                    assert (np.minimum(verts.max(axis=0), hi) >= np.maximum(verts.min(axis=0), lo) - 1e-6).all()
                seen_tris[first:first + count] = True
        assert rank == bin(int(w["imask"][i])).count("1")
    assert seen_nodes.all() and seen_tris.all()
    assert np.array_equal(np.sort(recs["primitive"]), np.arange(info.num_tree_triangles))
    # every record carries the exact box of the reference leaf that owns its triangle (the gate)
    nodes = mesh.data(lib.MESH_BINARY_NODES)
    leaves = np.nonzero(nodes["max_data"] != 0)[0]
    start_sorted = leaves[np.argsort(nodes["min_data"][leaves])]
    owner = np.repeat(start_sorted, nodes["max_data"][start_sorted])  # tree-order triangle -> leaf node
    gate_min = np.stack([recs["leaf_min_x"], recs["leaf_min_y"], recs["leaf_min_z"]], -1)
    assert np.array_equal(gate_min, nodes["min"][owner[recs["primitive"]]])
    assert np.array_equal(recs["leaf_max"], nodes["max"][owner[recs["primitive"]]])
