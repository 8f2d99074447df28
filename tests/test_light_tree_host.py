"""Host light-tree builder (zyg_b200/csrc/host/light_tree_builder.cpp, restating light_tree_builder.zig:281-376, 446-789)
checked through the oracle's Tree.randomLight / Tree.pdf (light_tree.zig:346-517). CPU only."""

import numpy as np
import pytest

import oracle_lib as oracle
from zyg_b200 import scenes, su


@pytest.fixture()
def engine():
    su.release()
    yield
    su.release()


def points(rng, k):
    p = np.stack([rng.uniform(-2.8, 2.8, k), rng.uniform(0.1, 2.9, k), rng.uniform(-2.8, 2.8, k)], -1).astype(np.float32)
    n = rng.normal(size=(k, 3))
    n = (n / np.linalg.norm(n, axis=1, keepdims=True)).astype(np.float32)
    return p, n


@pytest.mark.parametrize("num_lights", [2, 3, 4, 5, 37, 200])
def test_pdfs_form_a_sub_distribution_that_the_picks_realise(engine, num_lights):
    """With the split threshold at 0 the tree picks at most one light. Tree.pdf over all lights sums to at most 1 (mass that
    descends into a node whose lights all face away is dropped: such picks return nothing), and the share of random
    numbers that yield a pick is that sum, whatever the shape of the tree."""
    scenes.many_lights_scene(32, 32, spp=1, num_lights=num_lights, split_threshold=0.0)
    scene, view = su.compile_scene()
    rng = np.random.default_rng(num_lights)
    k = 400
    for p, n in zip(*points(rng, 8)):
        p = p - np.float32([0.0, 1.4, -2.8])  # camera-relative world, space.zig:94
        total = sum(oracle.light_tree_pdf(scene, view, p, n, 0.0, l) for l in range(num_lights))
        assert total <= 1.0 + 2e-4
        picked = sum(len(oracle.light_tree_random(scene, view, p, n, float(r), 0.0)) for r in (np.arange(k) + 0.5) / k)
        assert picked / k == pytest.approx(total, abs=2.0 / k + 1e-3)


@pytest.mark.parametrize("split_threshold", [0.0, 0.0625, 1.0])
def test_random_light_pdf_matches_pdf_query(engine, split_threshold):
    """The MIS invariant: the pdf returned with a pick equals Tree.pdf of that light from the same point, with and
    without adaptive splitting; a split returns every light at most once."""
    num_lights = 150
    scenes.many_lights_scene(32, 32, spp=1, num_lights=num_lights, split_threshold=0.5)
    scene, view = su.compile_scene()
    rng = np.random.default_rng(7)
    most = 0
    for p, n in zip(*points(rng, 40)):
        p = p - np.float32([0.0, 1.4, -2.8])
        for r in rng.random(5):
            picks = oracle.light_tree_random(scene, view, p, n, float(r), split_threshold)
            ids = [i for i, _ in picks]
            assert len(set(ids)) == len(ids)
            most = max(most, len(picks))
            for light, pdf in picks:
                assert 0 <= light < num_lights
                assert pdf == pytest.approx(oracle.light_tree_pdf(scene, view, p, n, split_threshold, light), rel=1e-5)
    assert most == 1 if 0.0 == split_threshold else most > 1


def test_picks_follow_the_pdf(engine):
    """Stratified random numbers pick each light with the frequency Tree.pdf states."""
    num_lights = 9
    scenes.many_lights_scene(32, 32, spp=1, num_lights=num_lights, split_threshold=0.0)
    scene, view = su.compile_scene()
    p, n = np.float32([0.3, -0.9, 2.5]), np.float32([0.0, 1.0, 0.0])
    k = 20000
    counts = np.zeros(num_lights)
    for r in (np.arange(k) + 0.5) / k:
        (light, _), = oracle.light_tree_random(scene, view, p, n, float(r), 0.0)
        counts[light] += 1
    pdfs = np.array([oracle.light_tree_pdf(scene, view, p, n, 0.0, l) for l in range(num_lights)])
    assert np.abs(counts / k - pdfs).max() < 2e-3


def test_wavefront_draw_order_is_statistically_equivalent(engine):
    """The device regroups the sampler draws of sampleLights (zyg_oracle.h: zo_set_wavefront_light_order). Both orders
    estimate the same image: means agree within the noise of the estimate."""
    w, spp = 48, 192
    scenes.many_lights_scene(w, w, spp=spp, num_lights=40)
    scene, view = su.compile_scene()
    a = oracle.render(scene, view, w, w, 0, spp)
    b = oracle.render(scene, view, w, w, 0, spp, wavefront_light_order=True)
    ia, ib = a[..., :3] / a[..., 3:4], b[..., :3] / b[..., 3:4]
    assert not np.array_equal(ia, ib)
    assert abs(ia.mean() - ib.mean()) / ia.mean() < 5e-3
    # per-pixel: differences are noise of two independent estimates, not a bias; compare coarse blocks
    blocks = lambda img: img.reshape(6, 8, 6, 8, 3).mean((1, 3))
    assert np.abs(blocks(ia) - blocks(ib)).max() / blocks(ia).mean() < 0.08
