"""Substitute roughness / metallic / normal maps on the host and in the oracle (substitute_material.zig:114-162,
material_helper.zig:16-79). The reference holds no vectors for this path, so the pins are identities:

* maps that hold one value everywhere give the film of the uniform-parameter material (a flat normal map leaves the shading normal
  alone up to the fp32 rounding of tangentToWorld + normalize);
* a byte image decodes like enc.unorm8ToFloat / snorm8ToFloat (encoding.zig:10-20);
* the adapted normal never lets the reflection of wo dip under the geometric surface (the point of sampleNormal's second half)."""

import numpy as np
import pytest

import oracle_lib as oracle
import scene_view as sv
from zyg_b200 import scenes, su


@pytest.fixture()
def engine():
    su.release()
    yield
    su.release()


def test_one_and_two_channel_images(engine):
    su.init()
    L = su._su()
    grey = np.arange(64, dtype=np.uint8).reshape(8, 8)
    assert 0 == su.image_create(grey)
    assert 1 == su.image_create(np.zeros((8, 8, 2), np.uint8))
    assert 2 == su.image_create(np.zeros((4, 4), np.float32))
    assert 3 == su.image_create(np.zeros((4, 4, 2), np.float32))
    px = np.zeros((4, 4, 4), np.uint8)
    assert -1 == L.su_image_create(0xFFFFFFFF, 0, 4, 4, 4, 1, 4, px.ctypes.data)  # Byte4: outside the scope


def test_map_of_the_wrong_channel_count_is_refused(engine):
    su.init()
    su.perspective_camera_create(16, 16)
    su.integrators_create({"surface": {"PTMIS": {}}})
    rgb = su.image_create(np.zeros((4, 4, 3), np.float32))
    m = su.material_create({"rendering": {"Substitute": {"roughness": {"id": rgb}}}})
    su.prop_create(su.RECTANGLE, [m])
    with pytest.raises(su.SuError):
        su.compile_scene()


def test_material_records_point_at_their_maps(engine):
    n = scenes.surface_maps_scene(32, 32, spp=1)
    scene, _ = su.compile_scene()
    s = sv.scene_at(scene)
    mats = sv.view(s.materials, np.dtype([("head", "<u4", 24), ("color_map", "<u4"), ("roughness_map", "<u4"), ("metallic_map", "<u4"),
                                          ("normal_map", "<u4"), ("coating", "<f4", 8)]), s.num_materials)
    null = 0xFFFFFFFF
    used = [(int(m["roughness_map"]) != null, int(m["metallic_map"]) != null, int(m["normal_map"]) != null) for m in mats]
    assert (True, False, True) in used and (True, True, False) in used and (False, True, True) in used
    assert n == 1 and s.num_image_samplers == 6


def _constant_images(monkeypatch):
    real_create = su.image_create

    def constant_image(pixels):  # every map becomes the constant surface_maps_scene passes for uniform=True; normal maps go flat
        px = np.asarray(pixels)
        if px.ndim == 3 and px.shape[2] == 2:
            return real_create(np.full(px.shape, 128, np.uint8) if px.dtype == np.uint8 else np.zeros(px.shape, np.float32))
        mean = float(px.mean()) / (255.0 if px.dtype == np.uint8 else 1.0)
        target = min([0.49, 0.53, 0.5], key=lambda c: abs(c - mean))
        return real_create(np.full(px.shape, target, np.float32))

    monkeypatch.setattr(su, "image_create", constant_image)


def test_constant_maps_equal_uniform_parameters(engine, monkeypatch):
    w, spp = 64, 8
    n = scenes.surface_maps_scene(w, w, spp=spp, uniform=True)
    scene, view = su.compile_scene()
    plain = oracle.render(scene, view, w, w, 0, spp, num_meshes=n)
    su.release()
    _constant_images(monkeypatch)

    # roughness and metallic maps that hold one value everywhere: the very same film
    n = scenes.surface_maps_scene(w, w, spp=spp, maps=("roughness", "metallic"))
    scene, view = su.compile_scene()
    mapped = oracle.render(scene, view, w, w, 0, spp, num_meshes=n)
    assert np.array_equal(mapped, plain)
    su.release()

    # a flat normal map keeps the shading normal but replaces the tangent frame by Frame.init(n) (substitute_material.zig:157-159):
    # the lobes are isotropic, so the same random numbers give directions rotated about n - another sample set of the same image
    n = scenes.surface_maps_scene(w, w, spp=64)
    scene, view = su.compile_scene()
    flat = oracle.render(scene, view, w, w, 0, 64, num_meshes=n)
    su.release()
    monkeypatch.undo()
    n = scenes.surface_maps_scene(w, w, spp=64, uniform=True)
    scene, view = su.compile_scene()
    plain = oracle.render(scene, view, w, w, 0, 64, num_meshes=n)
    a, b = flat[..., :3] / flat[..., 3:], plain[..., :3] / plain[..., 3:]
    assert abs(a.mean() - b.mean()) < 0.01 * b.mean()
    assert np.abs(a - b).mean() < 0.1 * b.mean()


def test_byte_maps_decode_like_the_reference(engine):
    """A unorm8 roughness map and the float map holding byte / 255 give the same film; same for snorm8 normals."""
    w, spp = 48, 4
    films = []
    real_create = su.image_create
    for as_float in (False, True):
        su.release()

        def create(pixels):
            px = np.asarray(pixels)
            if as_float and px.dtype == np.uint8 and (px.ndim == 2 or px.shape[2] == 1):
                return real_create((px.astype(np.float32) * np.float32(1.0 / 255.0)).astype(np.float32))
            if as_float and px.dtype == np.uint8 and px.shape[2] == 2:
                return real_create((px.astype(np.float32) * np.float32(1.0 / 128.0) - np.float32(1.0)).astype(np.float32))
            return real_create(px)

        su.image_create = create
        try:
            n = scenes.surface_maps_scene(w, w, spp=spp)
        finally:
            su.image_create = real_create
        scene, view = su.compile_scene()
        films.append(oracle.render(scene, view, w, w, 0, spp, num_meshes=n))
    assert np.array_equal(films[0], films[1])


def test_maps_change_the_image(engine):
    w, spp = 64, 16
    n = scenes.surface_maps_scene(w, w, spp=spp)
    scene, view = su.compile_scene()
    mapped = oracle.render(scene, view, w, w, 0, spp, num_meshes=n)
    su.release()
    n = scenes.surface_maps_scene(w, w, spp=spp, uniform=True)
    scene, view = su.compile_scene()
    plain = oracle.render(scene, view, w, w, 0, spp, num_meshes=n)
    a, b = mapped[..., :3] / mapped[..., 3:], plain[..., :3] / plain[..., 3:]
    assert np.isfinite(a).all() and (a >= 0).all()
    assert np.abs(a - b).mean() > 0.02 * b.mean()          # the bumps and the glossy stripes are visible
    assert abs(a.mean() - b.mean()) < 0.25 * b.mean()      # ... and energy stays in the same range
