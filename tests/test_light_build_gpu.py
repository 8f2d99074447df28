"""Device light-tree build (light_tree_builder.zig:281-428 as an LBVH with bottom-up statistics). The tree is
another valid tree than the reference's cost-driven one, so the checks are:

* structure: the serialised nodes are a tree over exactly the finite lights, children adjacent, `middle` splits a node's range,
  powers add up, every light's order is its tree position;
* parity through the oracle: the CPU restatement of Tree.randomLight / pdf and PrimitiveTree.randomLight / pdf walks whatever tree
  the compiled scene holds, so a frame rendered with device-built trees has to match the oracle's frame per pixel;
* same estimator: frames with host-built and device-built trees agree statistically (both unbiased)."""

import numpy as np
import pytest

import oracle_lib as oracle
import scene_view as sv
from zyg_b200 import scenes, su

pytestmark = pytest.mark.gpu


@pytest.fixture()
def device_trees(monkeypatch):
    su.release()
    monkeypatch.setattr(su, "LIGHT_TREE_BUILDER", 1)
    yield
    su.release()


def download_film(width, height):
    import ctypes as C

    from zyg_b200 import lib

    L = lib.load_library()
    L.zygpu_download_film.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
    film = np.zeros((height, width, 4), np.float32)
    assert 0 == L.zygpu_download_film(su.device_handle(), film.ctypes.data, width * height)
    return film


def check_tree(nodes, middles, mapping, orders, powers, first, count):
    """Walks the serialised tree; returns the number of leaves."""
    assert np.array_equal(orders[mapping[first:first + count]], np.arange(first, first + count))
    leaves, stack = 0, [(0, first, first + count)]
    seen = np.zeros(len(nodes), bool)
    while stack:
        n, begin, end = stack.pop()
        assert not seen[n]
        seen[n] = True
        node = nodes[n]
        assert node["num_lights"] == end - begin
        want = powers[mapping[begin:end]].sum(dtype=np.float64)
        assert abs(node["power"] - want) <= 1e-4 * max(want, 1e-20)
        if node["meta"] & 1:
            c, mid = int(node["meta"] >> 2), int(middles[n])
            assert begin < mid < end
            stack += [(c, begin, mid), (c + 1, mid, end)]
        else:
            assert int(node["meta"] >> 2) == begin
            leaves += 1
    assert seen.all()
    return leaves


def test_scene_tree_structure(device_trees):
    num = 400
    scenes.many_lights_scene(64, 64, spp=1, num_lights=num, split_threshold=0.5)
    scene, _ = su.compile_scene()
    s = sv.scene_at(scene)
    t = s.light_tree
    assert t.num_lights == num and t.num_nodes == 2 * num - 1  # single-light leaves
    nodes = sv.view(t.nodes, sv.LIGHT_NODE_DTYPE, t.num_nodes)
    middles = sv.view(t.node_middles, "<u4", t.num_nodes)
    mapping = sv.view(t.light_mapping, "<u4", t.num_lights)
    orders = sv.view(t.light_orders, "<u4", t.num_lights)
    powers = sv.view(s.light_aabbs, sv.AABB_DTYPE, t.num_lights)["min"][:, 3]
    assert check_tree(nodes, middles, mapping, orders, powers, t.num_infinite_lights, num - t.num_infinite_lights) == num
    assert 1 <= t.max_split_depth <= 10


@pytest.mark.parametrize("num_lights,split_threshold", [(400, 0.5), (64, 0.0)])
def test_many_lights_with_device_tree_match_oracle(device_trees, num_lights, split_threshold):
    w, spp = 96, 16
    scenes.many_lights_scene(w, w, spp=spp, num_lights=num_lights, split_threshold=split_threshold)
    scene, view = su.compile_scene()
    ref = oracle.render(scene, view, w, w, 0, spp, wavefront_light_order=True)
    su.render_frame(0)
    gpu = download_film(w, w)
    assert np.array_equal(gpu[..., 3], ref[..., 3])
    rel = np.abs(gpu[..., :3] - ref[..., :3]).sum(-1) / np.maximum(np.abs(ref[..., :3]).sum(-1), 1e-6)
    assert np.median(rel) < 1e-4
    assert (rel > 1e-2).mean() < 1e-2
    assert abs(gpu[..., :3].mean() - ref[..., :3].mean()) / ref[..., :3].mean() < 1e-4


def test_mesh_lights_with_device_trees_match_oracle(device_trees):
    """Scene tree and the per-part primitive trees (the 576-triangle emitter and the icosahedra) all come from the device."""
    w, spp = 96, 16
    n = scenes.mesh_lights_scene(w, w, spp=spp, split_threshold=0.5)
    scene, view = su.compile_scene()
    s = sv.scene_at(scene)
    for ms in sv.mesh_samplers(s):  # every part tree is a tree over the part's triangles with leaves of at most four
        nodes = sv.view(ms.nodes, sv.LIGHT_NODE_DTYPE, ms.num_nodes)
        middles = sv.view(ms.node_middles, "<u4", ms.num_nodes)
        mapping = sv.view(ms.light_mapping, "<u4", ms.num_triangles)
        orders = sv.view(ms.light_orders, "<u4", ms.num_triangles)
        pdfs = sv.view(ms.triangle_pdfs, "<f4", ms.num_triangles)
        check_tree(nodes, middles, mapping, orders, pdfs, 0, ms.num_triangles)
        leaf = (nodes["meta"] & 1) == 0
        assert nodes["num_lights"][leaf].max() <= 4
    ref = oracle.render(scene, view, w, w, 0, spp, num_meshes=n, wavefront_light_order=True)
    su.render_frame(0)
    gpu = download_film(w, w)
    assert np.array_equal(gpu[..., 3], ref[..., 3])
    rel = np.abs(gpu[..., :3] - ref[..., :3]).sum(-1) / np.maximum(np.abs(ref[..., :3]).sum(-1), 1e-6)
    assert np.median(rel) < 2e-5
    assert (rel > 1e-2).mean() < 1e-2
    assert abs(gpu[..., :3].mean() - ref[..., :3].mean()) / ref[..., :3].mean() < 1e-4


def test_host_and_device_trees_give_the_same_image(monkeypatch):
    w, spp = 64, 256
    films = []
    for builder in (0, 1):
        su.release()
        monkeypatch.setattr(su, "LIGHT_TREE_BUILDER", builder)
        scenes.many_lights_scene(w, w, spp=spp, num_lights=200, split_threshold=0.5)
        su.render_frame(0)
        f = download_film(w, w)
        films.append(f[..., :3] / f[..., 3:4])
    su.release()
    a, b = films
    assert abs(a.mean() - b.mean()) / a.mean() < 0.01
    # per pixel: the difference is noise of two independent estimates, far below the signal
    assert np.abs(a - b).mean() / a.mean() < 0.08


def test_device_build_of_many_lights_is_fast(device_trees):
    import ctypes as C

    from zyg_b200 import lib

    L = lib.load_library()
    L.zygpu_light_tree_build_ms.argtypes = [C.c_int]
    L.zygpu_light_tree_build_ms.restype = C.c_float
    scenes.many_lights_scene(32, 32, spp=1, num_lights=20000, split_threshold=0.5)
    su.compile_scene()
    times = []
    for _ in range(4):  # the kernels take under a millisecond; the rest is the allocator, which varies with what ran before
        L.zygpu_light_tree_build_ms(1)
        su.compile_scene()
        times.append(L.zygpu_light_tree_build_ms(1))
    assert 0 < min(times) < 50, f"device light-tree build of 20 000 lights took {min(times):.1f} ms at best ({times})"
