"""The Transparent sensor buffer (buffer_transparent.zig, sensor "alpha_transparency" take_loader.zig:194-196) and the
alpha PathtracerMIS hands it (Pool.transparency, vertex.zig:243-268; pathtracer_mis.zig:75-84, 163, 169-170) in the oracle. The reference
holds no vectors for it; the pins are what the model says about simple scenes."""

import numpy as np
import pytest

import oracle_lib as oracle
from test_su_api import View
from zyg_b200 import scenes, su


@pytest.fixture()
def engine():
    su.release()
    yield
    su.release()


def test_sensor_block_selects_the_buffer_class(engine):
    scenes.cornell_box(16, 16, spp=1)
    _, view = su.compile_scene()
    assert 0 == View.from_address(view).alpha_transparency
    su.sensor_create({"alpha_transparency": True})
    _, view = su.compile_scene()
    assert 1 == View.from_address(view).alpha_transparency


@pytest.mark.parametrize("filter_name", [None, "Mitchell"])
def test_opaque_surfaces_cover_and_empty_space_does_not(engine, filter_name):
    w, spp = 48, 8
    scenes.cornell_box(w, w, spp=spp)
    su.sensor_create({"alpha_transparency": True, "filter": {filter_name: {}}} if filter_name else {"alpha_transparency": True})
    scene, view = su.compile_scene()
    film, alpha = oracle.render_alpha(scene, view, w, w, 0, spp)
    plain = oracle.render(scene, view, w, w, 0, spp)
    assert np.allclose(film, plain, rtol=1e-5, atol=1e-6)  # the colour lanes are the Opaque buffer's (r > 0: atomics add in thread order)
    rgba = oracle.resolve_transparent(view, film, alpha)
    a = rgba[..., 3]
    slack = 0.1 if filter_name else 1e-5  # Mitchell has negative lobes: a filtered alpha may overshoot at an edge
    assert a.min() >= -slack and a.max() <= 1.0 + slack
    assert np.allclose(a[w // 2 - 8 : w // 2 + 8, w // 2 - 8 : w // 2 + 8], 1.0, atol=1e-5)  # walls and boxes: every path ends opaque
    assert np.allclose(a[0, 0], 0.0, atol=0.26 if filter_name else 1e-6)  # the corners look past the open box into nothing
    assert np.allclose(rgba[..., :3], oracle.resolve(view, film)[..., :3])


def test_sky_covers_by_its_brightness_and_glass_lets_it_through(engine):
    w, spp = 48, 16
    # a constant sky of radiance 0.25 and no geometry in the upper half of the image: a path that leaves keeps 1 - min(L, 1) of its
    # throughput, so alpha = average(min(L, 1)) (pathtracer_mis.zig:81-83, vertex.zig:260-261)
    scenes.sky_scene(w, w, spp=spp, uniform_sky=0.25, sun=None, objects=False)
    su.sensor_create({"alpha_transparency": True})
    scene, view = su.compile_scene()
    film, alpha = oracle.render_alpha(scene, view, w, w, 0, spp)
    a = oracle.resolve_transparent(view, film, alpha)[..., 3]
    sky_l = (film[2, w // 2, :3] / film[2, w // 2, 3]).mean()
    assert np.allclose(a[2, :], min(sky_l, 1.0), rtol=1e-4)
    assert np.allclose(a[w - 2, :], 1.0, atol=1e-5)  # the ground

    su.release()
    scenes.cornell_box(w, w, spp=spp, glass={"roughness": 0.0})
    su.sensor_create({"alpha_transparency": True})
    scene, view = su.compile_scene()
    film, alpha = oracle.render_alpha(scene, view, w, w, 0, spp)
    a = oracle.resolve_transparent(view, film, alpha)[..., 3]
    assert a.min() >= 0.0 and a.max() <= 1.0 + 1e-4 and np.isfinite(a).all()
    assert np.allclose(a[4:10, w // 2 - 4 : w // 2 + 4], 1.0, atol=1e-5)  # the back wall above the boxes
