"""The Disk shape as a prop and as a light (disk.zig:28-134, 171-332, 492-533): host classification and the oracle's restatement against
the geometry and the integrals it describes."""

import numpy as np
import pytest

import oracle_lib as oracle
from zyg_b200 import scenes, su


@pytest.fixture()
def engine():
    su.release()
    yield
    su.release()


def test_disk_prop_is_round_and_a_mapped_disk_light_is_refused(engine):
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
    su.compile_scene()  # a Disk light is in scope ...

    image = su.image_create(np.full((4, 4, 3), 0.5, np.float32))
    mapped = su.material_create({"rendering": {"Light": {"emittance": {"emission_map": {"id": image}, "value": 5.0}}}})
    lamp2 = su.prop_create(su.DISK, [mapped])
    su.light_create(lamp2)
    with pytest.raises(su.SuError):  # ... Disk.sampleMaterialTo is not
        su.compile_scene()


def test_disk_light_irradiance_matches_the_closed_form(engine):
    """A Lambertian floor under a one-sided Disk lamp of radius R at height h, seen at the point below its centre: the radiance is
    albedo * L * R^2 / (R^2 + h^2). One bounce, so only Disk.sampleTo / Disk.pdf / Disk.emission contribute; split_threshold 0 and
    1 exercise the one-sample and the all-samples branches of the light tree."""
    w, spp = 8, 256
    for num_samples in (1, 4):
        su.release()
        su.init()
        camera = su.perspective_camera_create(w, w)
        su.camera_set_fov(float(np.radians(2.0)))
        su.prop_set_transformation(camera, su.transformation(position=(0.0, 1.0, -3.0), rotation_deg=(-18.434948, 0.0, 0.0)))
        su.sampler_create(spp)
        su.integrators_create({"surface": {"PTMIS": {"depth": {"surface": 1}}}})
        su.sensor_create({})
        floor = su.material_create({"rendering": {"Substitute": {"color": [1.0, 1.0, 1.0], "roughness": 1.0}}})
        g = su.prop_create(su.RECTANGLE, [floor])
        su.prop_set_transformation(g, su.transformation((0.0, 0.0, 0.0), (40.0, 40.0, 1.0), (90.0, 0.0, 0.0)))
        light = su.material_create({"rendering": {"Light": {"emittance": {"value": 2.0, "num_samples": num_samples}}}})
        lamp = su.prop_create(su.DISK, [light], unoccluding=True)
        su.prop_set_transformation(lamp, su.transformation((0.0, 2.0, 0.0), (3.0, 3.0, 1.0), (-90.0, 0.0, 0.0)))
        su.light_create(lamp)
        scene, view = su.compile_scene()
        film = oracle.render(scene, view, w, w, 0, spp)
        img = film[..., :3] / film[..., 3:]
        expect = 2.0 * 1.5 ** 2 / (1.5 ** 2 + 2.0 ** 2)
        got = float(img[w // 2 - 1: w // 2 + 1, w // 2 - 1: w // 2 + 1, 1].mean())
        assert abs(got - expect) < 0.03 * expect, (num_samples, got, expect)


@pytest.mark.parametrize("disk_lights", [False, True])
def test_disk_scene_renders_with_round_shadows(engine, disk_lights):
    w, spp = 64, 16
    scenes.disk_scene(w, w, spp=spp, disk_lights=disk_lights)
    scene, view = su.compile_scene()
    film = oracle.render(scene, view, w, w, 0, spp)
    img = film[..., :3] / film[..., 3:]
    assert np.isfinite(img).all() and img.min() >= 0 and 0.02 < img.mean() < 2.0


def test_disk_lights_draw_orders_are_statistically_equivalent(engine):
    """Disk.sampleTo takes a 1D draw per sample; the device takes the draws of all picks before the draws of Light.evaluateTo
    (zyg_oracle.h: zo_set_wavefront_light_order). Both orders estimate the same image."""
    w, spp = 48, 128
    scenes.disk_scene(w, w, spp=spp, disk_lights=True)
    scene, view = su.compile_scene()
    a = oracle.render(scene, view, w, w, 0, spp)
    b = oracle.render(scene, view, w, w, 0, spp, wavefront_light_order=True)
    ia, ib = a[..., :3] / a[..., 3:4], b[..., :3] / b[..., 3:4]
    assert not np.array_equal(ia, ib)
    assert abs(ia.mean() - ib.mean()) / ia.mean() < 1e-2
    blocks = lambda img: img.reshape(6, 8, 6, 8, 3).mean((1, 3))
    assert np.abs(blocks(ia) - blocks(ib)).max() / blocks(ia).mean() < 0.15
