"""Two restatements of the reference's scene-compile builders have to agree byte for byte: the product's host
(zyg_b200/csrc/host/{bvh_builder,triangle_tree,light_tree_builder,mesh_sampler,scene_model}.cpp) and the oracle's own
(oracle/builders.cpp), both written from builder_base.zig / split_candidate.zig / triangle_tree_builder.zig /
prop_tree_builder.zig / light_tree_builder.zig / triangle_mesh.zig (Part.configure). An error in either shows here; before
this test the oracle consumed the product-built arrays and could not see one. CPU only."""

import ctypes as C

import numpy as np
import pytest

import oracle_lib as oracle
import scene_view as sv
from zyg_b200 import lib, scenes, su

PRODUCT = {oracle.BuiltMesh.NODES: lib.MESH_BINARY_NODES, oracle.BuiltMesh.TRIANGLES: lib.MESH_TRIANGLES,
           oracle.BuiltMesh.ORIGINAL: lib.MESH_ORIGINAL, oracle.BuiltMesh.POSITIONS: lib.MESH_POSITIONS,
           oracle.BuiltMesh.NORMALS: lib.MESH_NORMALS, oracle.BuiltMesh.UVS: lib.MESH_UVS, oracle.BuiltMesh.PARTS: lib.MESH_PARTS}
NAMES = ("nodes", "triangles", "original", "positions", "normals", "uvs", "parts")


@pytest.fixture()
def engine():
    su.release()
    yield
    su.release()


def assert_same_mesh(product: lib.Mesh, mine: oracle.BuiltMesh):
    for which, pw in PRODUCT.items():
        assert product.data(pw).tobytes() == mine.raw(which), f"{NAMES[which]} differ between the two builders"
    d = mine.diagnostics()
    # the two places where both restatements deviate from the literal reference (documented in DESIGN.md) never trigger
    assert 0 == d["leaf_offset_mismatches"] and 0 == d["task_root_leaves"] and 0 == d["unsplittable"]
    assert 0 == product.info().num_leaf_order_fixups


@pytest.mark.parametrize("quads", [(4, 2), (16, 8), (48, 24), (100, 50), (250, 125)])
def test_triangle_tree_is_identical(quads):
    """16 to 62 500 triangles: below the sweep threshold, above it (sliced planes + spatial splits), and above
    ParallelizeThreshold (sub-tree tasks appended to the main kernel, builder_base.zig:354-390)."""
    positions, normals, uvs, indices = scenes.displaced_sphere(*quads)
    assert_same_mesh(lib.Mesh(positions, indices, normals, uvs), oracle.BuiltMesh(positions, indices, normals, uvs))


def test_triangle_tree_with_parts_and_without_attributes():
    """Several parts (shape_provider.zig:863-898 fills triangles part by part), no normals / uvs (defaults), and a mesh whose
    long thin triangles make the spatial splits duplicate references."""
    positions, normals, uvs, indices = scenes.displaced_sphere(40, 20, seed=0x5EED0009)
    n = indices.shape[0]
    parts = np.array([[0, 3 * (n // 3), 0], [3 * (n // 3), 3 * (n // 2 - n // 3), 1], [3 * (n // 2), 3 * (n - n // 2), 2]], np.uint32)
    a = lib.Mesh(positions, indices, parts=parts)
    b = oracle.BuiltMesh(positions, indices, parts=parts)
    assert_same_mesh(a, b)
    assert set(np.unique(b.data(b.PARTS))) == {0, 1, 2}

    rng = np.random.default_rng(3)
    k = 3000  # random needles: long, thin, overlapping boxes
    base = rng.uniform(-1, 1, (k, 3))
    d = rng.normal(size=(k, 3))
    pos = np.concatenate([base, base + d, base + d + 0.01 * rng.normal(size=(k, 3))]).astype(np.float32)
    idx = np.stack([np.arange(k), np.arange(k) + k, np.arange(k) + 2 * k], -1).astype(np.uint32)
    a, b = lib.Mesh(pos, idx), oracle.BuiltMesh(pos, idx)
    assert_same_mesh(a, b)
    assert b.data(b.ORIGINAL).size > k, "spatial splits were expected to duplicate references"


def test_builder_result_does_not_depend_on_threads():
    positions, normals, uvs, indices = scenes.displaced_sphere(120, 60)
    one = oracle.BuiltMesh(positions, indices, normals, uvs, threads=1)
    many = oracle.BuiltMesh(positions, indices, normals, uvs, threads=8)
    for which in PRODUCT:
        assert one.raw(which) == many.raw(which)


def test_oracle_built_tree_traverses_like_brute_force():
    """The pin that does not involve the product at all: closest hits through the oracle-built tree equal the O(N) loop."""
    positions, normals, uvs, indices = scenes.displaced_sphere(64, 32)
    m = oracle.BuiltMesh(positions, indices, normals, uvs)
    rays = np.concatenate([scenes.primary_rays(48, 48), scenes.random_rays(4096)])
    hits = oracle.trace_closest(m.data(m.NODES), m.data(m.TRIANGLES), m.data(m.POSITIONS), rays)
    brute, _ties = oracle.brute_closest(m.data(m.TRIANGLES), m.data(m.POSITIONS), rays)
    assert np.array_equal(hits["t"].view(np.uint32), brute["t"].view(np.uint32))
    miss = hits["primitive"] == 0xFFFFFFFF
    assert np.array_equal(miss, brute["primitive"] == 0xFFFFFFFF) and 0 < miss.sum() < rays.size
    # leaves partition the references: every source triangle appears, duplicates only through spatial splits
    assert set(m.data(m.ORIGINAL)) == set(range(indices.shape[0]))


def product_scene():
    address, view = su.compile_scene()
    s = sv.scene_at(address)
    return address, view, s


def check_prop_trees(s):
    props = sv.view(s.props, sv.PROP_DTYPE, s.num_props)
    aabbs = sv.view(s.aabbs, sv.AABB_DTYPE, s.num_props)
    infinite = np.isin(props["shape"], (sv.SHAPE_CANOPY, sv.SHAPE_DISTANT, sv.SHAPE_DOME))
    unocc = (props["flags"] & sv.PROP_UNOCCLUDING) != 0
    checked = 0
    for tree, want_unocc in ((s.solid_bvh, False), (s.unoccluding_bvh, True)):
        product_indices = sv.view(tree.indices, "<u4", tree.num_indices)
        # Scene.classifyProp (scene.zig:322-340) hands the builder the classified props in creation order; which props are in
        # the tree at all (prototypes of instancers, the camera entity are not) is read off the product's index list
        members = np.zeros(s.num_props, bool)
        members[product_indices] = True
        ids = np.nonzero(members & ~infinite & (unocc == want_unocc))[0].astype(np.uint32)
        assert ids.size == np.unique(product_indices).size
        nodes, indices = oracle.build_prop_tree(ids, aabbs)
        assert nodes == sv.view(tree.nodes, sv.NODE_DTYPE, tree.num_nodes).tobytes()
        assert indices == product_indices.tobytes()
        checked += tree.num_nodes
    return checked


def check_light_tree(s):
    props = sv.view(s.props, sv.PROP_DTYPE, s.num_props)
    lights = sv.view(s.lights, sv.LIGHT_DTYPE, s.num_lights)
    light_aabbs = sv.view(s.light_aabbs, sv.AABB_DTYPE, s.num_lights)
    light_cones = sv.view(s.light_cones, "<f4", 4 * s.num_lights)
    finite = ~np.isin(props["shape"][lights["prop"]], (sv.SHAPE_CANOPY, sv.SHAPE_DISTANT, sv.SHAPE_DOME))
    mine = oracle.build_light_tree(light_aabbs, light_cones, lights["two_sided"] != 0, finite)
    t = s.light_tree
    assert mine["nodes"] == sv.view(t.nodes, sv.LIGHT_NODE_DTYPE, t.num_nodes).tobytes()
    assert mine["middles"] == sv.view(t.node_middles, "<u4", t.num_nodes).tobytes()
    assert mine["orders"] == sv.view(t.light_orders, "<u4", t.num_lights).tobytes()
    assert mine["mapping"] == sv.view(t.light_mapping, "<u4", t.num_lights).tobytes()
    if t.num_infinite_lights > 0:
        assert mine["infinite_cdf"] == sv.view(t.infinite_cdf, "<f4", t.num_infinite_lights + 1).tobytes()
    f = np.frombuffer(mine["floats"], np.float32)
    u = np.frombuffer(mine["uints"], np.uint32)
    if t.num_nodes > 0:
        assert f[:4].tobytes() == bytes(t.bounds.min) and f[4:8].tobytes() == bytes(t.bounds.max)
    assert f[8].tobytes() == np.float32(t.infinite_weight).tobytes() and f[9].tobytes() == np.float32(t.infinite_guard).tobytes()
    assert (int(u[0]), int(u[1]), int(u[2])) == (t.infinite_end, t.max_split_depth, t.num_infinite_lights)
    return t.num_nodes


def check_mesh_samplers(s, num_meshes):
    L = lib.load_library()
    table = oracle.mesh_table(num_meshes)
    part_areas = sv.view(s.mesh_part_areas, "<f4", s.num_parts)
    checked = 0
    for ms in sv.mesh_samplers(s):
        handle = su._su().zyg_su_mesh(7 + ms.mesh)
        n = C.c_uint64()
        L.zyg_mesh_data(handle, lib.MESH_PARTS, C.byref(n))
        tree_triangles = n.value // 2
        parts = np.frombuffer(C.string_at(L.zyg_mesh_data(handle, lib.MESH_PARTS, None), n.value), np.uint16)
        num_parts = int(parts.max()) + 1
        tm = sv.view(ms.triangle_mapping, "<u4", ms.num_triangles)
        part = int(parts[tm[0]])
        mine = oracle.build_mesh_sampler(C.byref(table[ms.mesh]), tree_triangles, num_parts, part, ms.two_sided)
        assert mine["triangle_mapping"] == tm.tobytes()
        assert mine["triangle_pdfs"] == sv.view(ms.triangle_pdfs, "<f4", ms.num_triangles).tobytes()
        assert mine["primitive_mapping"] == sv.view(ms.primitive_mapping, "<u4", tree_triangles).tobytes()
        tree = mine["tree"]
        assert tree["nodes"] == sv.view(ms.nodes, sv.LIGHT_NODE_DTYPE, ms.num_nodes).tobytes()
        assert tree["middles"] == sv.view(ms.node_middles, "<u4", ms.num_nodes).tobytes()
        assert tree["orders"] == sv.view(ms.light_orders, "<u4", ms.num_triangles).tobytes()
        assert tree["mapping"] == sv.view(ms.light_mapping, "<u4", ms.num_triangles).tobytes()
        f = np.frombuffer(tree["floats"], np.float32)
        assert f[:4].tobytes() == bytes(ms.bounds.min) and f[4:8].tobytes() == bytes(ms.bounds.max)
        # Part.area of the sampled part, as the scene carries it for the props that use this mesh
        areas = np.frombuffer(mine["part_areas"], np.float32)
        assert np.any(part_areas == areas[part])
        checked += 1
    return checked


SCENES = {
    "config1_cornell": (lambda: scenes.cornell_box(64, 64, spp=1), 0),
    "config3_instanced": (lambda: scenes.instanced_scene(64, 64, spp=1, grid=(40, 40), prototypes=6, quads=(24, 12), sun=60.0), 6),
    "config3_instancer_entity": (lambda: scenes.instanced_scene(64, 64, spp=1, grid=(12, 12), prototypes=3, quads=(12, 6),
                                                                instancer=su.transformation((0.5, 0.0, 0.25), (1.0, 1.0, 1.0), (0.0, 20.0, 0.0))), 3),
    "many_lights": (lambda: scenes.many_lights_scene(64, 64, spp=1, num_lights=300), 0),
    "config4_mesh_lights": (lambda: scenes.mesh_lights_scene(64, 64, spp=1, num_lights=200, geometry_quads=(40, 20), sun=15.0, sky=64,
                                                             unoccluding=True), 3),
    "sphere_lights": (lambda: scenes.sphere_lights_scene(64, 64, spp=1), 0),
    "sky": (lambda: scenes.sky_scene(64, 64, spp=1, sky_size=32), 0),
}


@pytest.mark.parametrize("name", list(SCENES))
def test_scene_trees_are_identical(engine, name):
    """Prop trees (solid + un-occluding), the scene light tree and every mesh-light sampler of the compiled scene, rebuilt by
    the oracle from the scene's primary records (world boxes, light boxes / cones / powers, mesh arrays)."""
    build, num_meshes = SCENES[name]
    r = build()
    num_meshes = r if isinstance(r, int) and r > 0 and num_meshes > 0 else num_meshes
    _address, _view, s = product_scene()
    assert check_prop_trees(s) > 0
    nodes = check_light_tree(s)
    assert nodes > 0 or "sky" == name
    samplers = check_mesh_samplers(s, num_meshes)
    if "config4_mesh_lights" == name:
        assert samplers >= 2 and nodes > 100
