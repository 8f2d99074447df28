"""zyg's C API (su_*, include/zyg_su.h) on the host side: return conventions of src/capi/capi.zig, the scene model
and what Scene.compile flattens. No GPU needed."""

import ctypes as C

import numpy as np
import pytest

import oracle_lib as oracle
from zyg_b200 import scenes, su


class Trafo(C.Structure):
    _fields_ = [("r", C.c_float * 4 * 3), ("position", C.c_float * 4)]


class View(C.Structure):
    _fields_ = [("resolution", C.c_int32 * 2), ("crop", C.c_int32 * 4), ("left_top", C.c_float * 4), ("d_x", C.c_float * 4),
                ("d_y", C.c_float * 4), ("eye_offset", C.c_float * 4), ("camera_trafo", Trafo), ("aperture_radius", C.c_float),
                ("focus_distance", C.c_float), ("sampler", C.c_uint32), ("spp_total", C.c_uint32),
                ("max_depth_surface", C.c_uint32), ("max_depth_volume", C.c_uint32), ("split_threshold", C.c_float),
                ("regularize_roughness", C.c_float), ("caustics_path", C.c_uint32), ("specular_threshold", C.c_float),
                ("clamp", C.c_float * 3), ("filter_radius_int", C.c_int32), ("filter_range_end", C.c_float),
                ("filter_inverse_interval", C.c_float), ("filter", C.c_float * 30), ("exposure_factor", C.c_float), ("aov_slots", C.c_uint32), ("alpha_transparency", C.c_uint32)]


@pytest.fixture()
def engine():
    su.release()
    yield
    su.release()


def test_init_release_conventions(engine):
    lib = su._su()
    assert lib.su_release() == -1  # not initialised
    assert lib.su_init() == 0
    assert lib.su_init() == -1  # capi.zig:58-60
    assert lib.su_sampler_create(16) == -1  # capi.zig:215-221 returns -1 on every path
    assert lib.su_material_update(99, b'{"rendering":{"Substitute":{}}}') == -3
    assert lib.su_prop_create(99, 0, None) == -1
    assert lib.su_light_create(12345) == -1
    assert lib.su_release() == 0


def test_ids_follow_creation_order(engine):
    su.init()
    cam = su.perspective_camera_create(32, 16)
    assert cam == 0
    dims = (C.c_int32 * 2)()
    assert su._su().su_camera_sensor_dimensions(dims) == 0 and list(dims) == [32, 16]
    m0 = su.material_create({"rendering": {"Substitute": {"color": [0.5, 0.5, 0.5]}}})
    m1 = su.material_create({"rendering": {"Light": {"emittance": {"value": 3.0}}}})
    assert (m0, m1) == (1, 2)  # id 0 is the fallback Debug material (capi.zig:94-101)
    assert su._su().su_material_create(0, b'{"rendering":{"Unknown":{}}}') == -1
    assert su._su().su_material_create(0, b'{"no_rendering":1}') == -1
    assert su._su().su_material_create(0, b'{not json') == -1
    p = su.prop_create(su.RECTANGLE, [m0])
    assert p == 1
    positions = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float32)
    shape = su.triangle_mesh_create(positions, np.array([0, 1, 2], np.uint32))
    assert shape == 7  # first id after the seven built-in shapes (manager.zig:36-44)


def test_compile_is_camera_relative(engine):
    scenes.cornell_box(64, 64, spp=4)
    scene, view = su.compile_scene()
    v = View.from_address(view)
    assert list(v.resolution) == [64, 64] and list(v.crop) == [0, 0, 64, 64]
    assert list(v.camera_trafo.position)[:3] == [0.0, 0.0, 0.0]  # space.zig:94,103-110
    assert v.spp_total == 4 and v.max_depth_surface == 8 and v.max_depth_volume == 8  # take.zig:254-261
    assert v.split_threshold == pytest.approx(0.5 ** 4) and v.filter_radius_int == 0
    assert v.specular_threshold == pytest.approx(0.01314 ** 2)
    # Perspective.update: left_top = (-1, ratio, 1 / tan(fov / 2))
    assert v.left_top[0] == -1.0 and v.left_top[1] == 1.0
    assert v.left_top[2] == pytest.approx(1.0 / np.tan(np.radians(39.0) / 2), rel=1e-6)


def test_mitchell_filter_table_is_normalised(engine):
    scenes.cornell_box(32, 32, spp=1, filter_name="Mitchell")
    _, view = su.compile_scene()
    v = View.from_address(view)
    assert v.filter_radius_int == 2 and v.filter_range_end == 2.0
    f = np.array(list(v.filter), np.float64)
    assert f[0] > 0 and abs(f[-1]) < 1e-6
    x = np.linspace(0.0, 2.0, 30)
    assert 2.0 * np.trapezoid(f, x) == pytest.approx(1.0, abs=2e-3)  # Sensor.integral normalisation, sensor.zig:120-122


def test_default_sensor_is_mitchell(engine):
    su.init()
    su.perspective_camera_create(8, 8)
    su.integrators_create({"surface": {"PTMIS": {}}})
    _, view = su.compile_scene()
    v = View.from_address(view)
    assert v.filter_radius_int == 2  # take.zig:59-64
    assert v.max_depth_surface == 16 and v.max_depth_volume == 256  # Default_depth, take.zig:77
    su.release()
    su.init()
    su.perspective_camera_create(8, 8)
    with pytest.raises(su.SuError):
        su.compile_scene()  # no PTMIS integrator configured


def test_oracle_render_is_thread_count_independent(engine):
    scenes.cornell_box(32, 32, spp=4)
    scene, view = su.compile_scene()
    a = oracle.render(scene, view, 32, 32, 0, 4, threads=1)
    b = oracle.render(scene, view, 32, 32, 0, 4, threads=5)
    assert a.tobytes() == b.tobytes()
    assert np.all(a[..., 3] == 4.0)


def test_oracle_sample_ranges_add_up(engine):
    """Driver.render(iteration, num_samples): sample ranges of one pixel are independent given the absolute sample
    index (worker.zig:145-149), which is what the multi-GPU split relies on."""
    scenes.cornell_box(32, 32, spp=8)
    scene, view = su.compile_scene()
    whole = oracle.render(scene, view, 32, 32, 0, 8)
    parts = oracle.render(scene, view, 32, 32, 0, 3)
    parts = oracle.render(scene, view, 32, 32, 3, 5, film=parts)
    assert whole.tobytes() == parts.tobytes()


def test_oracle_cornell_mirror_symmetry(engine):
    """With both side walls white and no boxes the box is mirror symmetric in x: so is the converged image."""
    su.init()
    cam = su.perspective_camera_create(48, 48)
    su.camera_set_fov(float(np.radians(39.0)))
    su.prop_set_transformation(cam, su.transformation(position=(0.0, 1.0, -3.9)))
    su.sampler_create(64)
    su.integrators_create({"surface": {"PTMIS": {"depth": {"surface": 4}}}})
    su.sensor_create({})
    white = su.material_create({"rendering": {"Substitute": {"color": [0.7, 0.7, 0.7], "roughness": 1.0}}})
    light = su.material_create({"rendering": {"Light": {"emittance": {"value": 10.0}}}})
    for position, rotation in (((0, 0, 0), (90, 0, 0)), ((0, 2, 0), (-90, 0, 0)), ((0, 1, 1), (0, 180, 0)),
                               ((-1, 1, 0), (0, -90, 0)), ((1, 1, 0), (0, 90, 0))):
        p = su.prop_create(su.RECTANGLE, [white])
        su.prop_set_transformation(p, su.transformation(position, (2, 2, 1), rotation))
    lamp = su.prop_create(su.RECTANGLE, [light], unoccluding=True)
    su.prop_set_transformation(lamp, su.transformation((0, 1.98, 0), (0.5, 0.5, 1), (-90, 0, 0)))
    su.light_create(lamp)
    scene, view = su.compile_scene()
    film = oracle.render(scene, view, 48, 48, 0, 64)
    img = film[..., :3] / film[..., 3:4]
    lum = img.mean(-1)
    left, right = lum[:, :24], lum[:, 24:][:, ::-1]
    assert abs(left.mean() - right.mean()) / lum.mean() < 0.01
    # block averages agree too (noise at 64 spp averaged over 8x8 blocks)
    bl = left.reshape(6, 8, 3, 8).mean((1, 3))
    br = right.reshape(6, 8, 3, 8).mean((1, 3))
    assert np.abs(bl - br).max() / lum.mean() < 0.08


def test_oracle_occluding_and_unoccluding_light_agree(engine):
    """A light created through su_prop_create sits in the solid tree (capi.zig:425-455), one created as a scene-file
    Light entity is gathered by the un-occluding pass (scene_loader.zig:365-366, prop_tree.zig:302-356). The lamp hangs
    2 cm under the ceiling, so both must give the same picture up to noise."""
    imgs = []
    for unocc in (False, True):
        su.release()
        scenes.cornell_box(32, 32, spp=64, unoccluding_light=unocc)
        scene, view = su.compile_scene()
        film = oracle.render(scene, view, 32, 32, 0, 64)
        imgs.append(film[..., :3] / film[..., 3:4])
    a, b = imgs
    assert abs(a.mean() - b.mean()) / a.mean() < 0.02
