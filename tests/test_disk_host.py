"""The Disk shape as a prop (disk.zig:28-134): host classification and the oracle's restatement against the geometry it describes."""

import numpy as np
import pytest

import oracle_lib as oracle
from zyg_b200 import scenes, su


@pytest.fixture()
def engine():
    su.release()
    yield
    su.release()


def test_disk_prop_is_round_and_a_disk_light_is_refused(engine):
    w = 96
    su.init()
    su.perspective_camera_create(w, w)
    su.camera_set_fov(float(np.radians(40.0)))
    su.sampler_create(4)
    su.integrators_create({"surface": {"PTMIS": {"depth": {"surface": 1}}}})
    su.sensor_create({})
    su.aovs_create({"Depth": True, "GeometricNormal": True})
    m = su.material_create({"rendering": {"Substitute": {"color": [0.5, 0.5, 0.5], "roughness": 1.0, "two_sided": True}}})
    d = su.prop_create(su.DISK, [m])
    su.prop_set_transformation(d, su.transformation((0.0, 0.0, 5.0), (2.0, 2.0, 1.0)))  # radius 1, facing the camera at distance 5
    scene, view = su.compile_scene()
    _, layers = oracle.render_aov(scene, view, w, w, 0, 4, (1 << 1) | (1 << 3))
    depth = layers[1][..., 0]
    hit = depth < 1e30
    # the silhouette is a circle of radius (1 / 5) / tan(20 deg) of the half width
    r_px = (1.0 / 5.0) / np.tan(np.radians(20.0)) * (w / 2)
    yy, xx = np.mgrid[0:w, 0:w]
    rr = np.hypot(xx + 0.5 - w / 2, yy + 0.5 - w / 2)
    assert hit[rr < r_px - 1.5].all() and not hit[rr > r_px + 1.5].any()
    assert abs(hit.sum() - np.pi * r_px * r_px) < 0.04 * np.pi * r_px * r_px
    assert abs(depth[w // 2, w // 2] - 5.0) < 1e-3
    n = layers[3][w // 2, w // 2, :3] / layers[3][w // 2, w // 2, 3]
    assert np.allclose(np.abs(n), [0.0, 0.0, 1.0], atol=1e-6)

    light = su.material_create({"rendering": {"Light": {"emittance": {"value": 5.0}}}})
    lamp = su.prop_create(su.DISK, [light])
    su.light_create(lamp)
    with pytest.raises(su.SuError):
        su.compile_scene()


def test_disk_scene_renders_with_round_shadows(engine):
    w, spp = 64, 16
    scenes.disk_scene(w, w, spp=spp)
    scene, view = su.compile_scene()
    film = oracle.render(scene, view, w, w, 0, spp)
    img = film[..., :3] / film[..., 3:]
    assert np.isfinite(img).all() and img.min() >= 0 and 0.02 < img.mean() < 2.0
