"""AOV layers (Worker.commonAOV, worker.zig:209-242; Sensor.addSample's AOV half, sensor.zig:197-377; aov.Buffer,
rendering/sensor/aov/aov_buffer.zig) on the host and in the oracle. The reference holds no vectors for this path; the pins are identities
of the algorithm: Emission + Direct + Indirect is the beauty, first-hit geometry against what the scene description says."""

import ctypes as C

import numpy as np
import pytest

import oracle_lib as oracle
from test_su_api import View
from zyg_b200 import scenes, su

ALL = {name: True for name in oracle.AOV_CLASSES}


@pytest.fixture()
def engine():
    su.release()
    yield
    su.release()


def test_aovs_create_sets_and_clears_class_bits(engine):
    scenes.cornell_box(16, 16, spp=1)
    su.aovs_create({"Albedo": True, "ShadingNormal": True, "Indirect": True, "NotAClass": True})
    _, view = su.compile_scene()
    assert View.from_address(view).aov_slots == (1 << 0) | (1 << 4) | (1 << 8)
    su.aovs_create({"Albedo": False, "Depth": True})  # View.loadAOV sets or clears: the other bits stay (take.zig:106-129)
    _, view = su.compile_scene()
    assert View.from_address(view).aov_slots == (1 << 1) | (1 << 4) | (1 << 8)
    assert -1 == su._su().su_aovs_create(b"{ not json")


@pytest.mark.parametrize("filter_name", [None, "Mitchell"])
def test_light_classes_add_up_to_the_beauty(engine, filter_name):
    w, spp = 48, 4
    scenes.cornell_box(w, w, spp=spp, filter_name=filter_name)
    su.aovs_create(ALL)
    scene, view = su.compile_scene()
    film, layers = oracle.render_aov(scene, view, w, w, 0, spp, 0x1FF, threads=1)
    assert sorted(layers) == list(range(9))
    plain = oracle.render(scene, view, w, w, 0, spp, threads=1)
    assert np.array_equal(film, plain)  # recording AOVs does not touch the beauty
    total = layers[6][..., :3] + layers[7][..., :3] + layers[8][..., :3]
    assert np.allclose(total, film[..., :3], rtol=2e-5, atol=1e-6)
    for c in (0, 3, 4, 5, 6, 7, 8):  # filtered classes share the beauty's weights
        assert np.allclose(layers[c][..., 3], film[..., 3], rtol=1e-5)


def test_first_hit_classes_of_the_cornell_box(engine):
    w, spp = 64, 4
    scenes.cornell_box(w, w, spp=spp)
    su.aovs_create(ALL)
    scene, view = su.compile_scene()
    _, layers = oracle.render_aov(scene, view, w, w, 0, spp, 0x1FF)
    res = {c: oracle.resolve_aov(c, layers[c]) for c in layers}

    depth = res[1][..., 0]  # camera rays hit something between the front edge and the back wall, or leave past the box (floatMax)
    hit = depth < 1e30
    assert hit.mean() > 0.9 and np.all(depth[~hit] == np.finfo(np.float32).max)
    assert depth.min() > 2.8 and depth[hit].max() < 5.6
    assert depth[2, w // 2] > depth[w - 2, w // 2] * 0.9  # ceiling / floor near the front edge are about equally far
    assert abs(depth[w // 2 - 12, w // 2] - 4.9) < 0.1    # the back wall above the boxes: camera z = -3.9, wall z = 1

    ids = res[2][..., 0]
    # overwritePixel keeps the pixel's first sample (r = 0: every weight is 1): 0 where that sample left the box
    assert np.array_equal(ids, np.round(ids)) and (ids[hit] >= 1).mean() > 0.97 and np.all(ids[~hit] == 0)
    assert len(np.unique(ids[hit])) >= 4  # white, red, green, light (+ the boxes' material)

    for c in (3, 4):
        n = res[c][..., :3]
        norm = np.linalg.norm(n, axis=-1)  # averages of unit normals over a pixel's samples (shorter across an edge, 0 for a miss)
        assert norm.max() < 1.0 + 1e-5 and np.median(norm[hit]) > 0.999
    back = res[3][w // 2 - 12, w // 2, :3]
    assert np.allclose(back, [0.0, 0.0, -1.0], atol=1e-5)  # the back wall faces the camera
    assert res[3][w // 2, 1, 0] > 0.99 and res[3][w // 2, w - 2, 0] < -0.99  # left wall +x, right wall -x

    rough = res[5][..., 0]
    assert np.allclose(rough[w // 2 - 12, w // 2], 1.0, atol=1e-5)  # walls: roughness 1 -> alpha 1 -> sqrt(alpha) 1

    albedo = res[0][..., :3]
    left, right = albedo[w // 2, 1], albedo[w // 2, w - 2]
    assert left[0] > 2.0 * left[1] and right[1] > 2.0 * right[0]  # red wall on the left, green on the right
    assert albedo.min() > -0.05 and albedo.max() < 1.5


def test_inactive_class_is_refused(engine):
    su.init()
    assert -1 == su._su().su_resolve_frame(0)  # no device yet
