"""The clear coat of a Substitute (substitute_coating.zig, substitute_sample.zig:138-142, 304-336, 412-433,
material_provider.zig:303-326) on the host and in the oracle. The reference holds no vectors for it; the pins are properties of the
model: the sampler's pdf is the pdf evaluate() reports, a coat conserves energy, a coat that vanishes leaves the base."""

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


def materials_of(scene):
    s = sv.scene_at(scene)
    dt = np.dtype([("head", "<u4", 28), ("absorption", "<f4", 3), ("thickness", "<f4"), ("ior", "<f4"), ("roughness", "<f4"), ("pad", "<f4", 2)])
    assert dt.itemsize == 144
    return sv.view(s.materials, dt, s.num_materials)


def test_coating_block_is_parsed(engine):
    scenes.coated_scene(16, 16, spp=1)
    scene, _ = su.compile_scene()
    mats = materials_of(scene)
    coated = [m for m in mats if m["thickness"] > 0]
    assert len(coated) == 4
    amber = max(coated, key=lambda m: m["thickness"])
    assert np.isclose(amber["thickness"], 0.2) and np.isclose(amber["ior"], 1.6) and np.isclose(amber["roughness"], 0.1)
    # setCoatingAttenuation -> attenuationCoefficient: -log(clamp(color, 0.01, 0.991102)) / distance of the colour json.readColor
    # returns (sRGB primaries -> AP1): amber keeps red, loses blue
    ab = amber["absorption"]
    assert ab[0] < ab[1] < ab[2] and 0.5 < ab[0] and ab[2] < -np.log(0.01) / 0.1 + 1e-3
    clear = min(coated, key=lambda m: m["roughness"])
    assert np.allclose(clear["absorption"], -np.log(np.float32(0.991102)) / np.float32(0.1), rtol=1e-5)  # the default white coat
    plain = [m for m in mats if m["thickness"] == 0]
    assert all(np.isclose(m["ior"], 1.5) or m["ior"] == 0 for m in plain)


def test_emissive_coated_substitute_is_refused(engine):
    su.init()
    su.perspective_camera_create(16, 16)
    su.integrators_create({"surface": {"PTMIS": {}}})
    m = su.material_create({"rendering": {"Substitute": {"emittance": {"value": 2.0}, "coating": {"thickness": 0.1}}}})
    su.prop_create(su.RECTANGLE, [m])
    with pytest.raises(su.SuError):
        su.compile_scene()


def test_coat_changes_the_image_and_keeps_energy_in_range(engine):
    w, spp = 64, 32
    films = {}
    for coated in (False, True):
        su.release()
        n = scenes.coated_scene(w, w, spp=spp, coated=coated)
        scene, view = su.compile_scene()
        f = oracle.render(scene, view, w, w, 0, spp, num_meshes=n)
        films[coated] = f[..., :3] / f[..., 3:]
    a, b = films[True], films[False]
    assert np.isfinite(a).all() and (a >= 0).all()
    assert np.abs(a - b).mean() > 0.03 * b.mean()       # highlights of the coat, the amber tint
    assert 0.6 * b.mean() < a.mean() < 1.3 * b.mean()   # the coat neither eats nor invents the light
    # the amber coat absorbs blue more than red on the cube (left part of the image)
    cube = (slice(w // 2 - 4, w // 2 + 14), slice(8, 24))
    assert (a[cube][..., 2] / np.maximum(a[cube][..., 0], 1e-6)).mean() < 0.8 * (b[cube][..., 2] / np.maximum(b[cube][..., 0], 1e-6)).mean()


def test_thin_coat_converges_to_the_uncoated_fresnel_layer(engine):
    """thickness -> 0+ keeps the coat's Fresnel layer but no absorption: two very thin coats give the same image (the thickness only enters
    through exp(-mu d)), and that image differs from the uncoated one by the coat's reflection alone."""
    w, spp = 48, 16
    films = []
    for t in (1e-6, 1e-7):
        su.release()
        n = scenes.coated_scene(w, w, spp=spp, thickness=t)
        scene, view = su.compile_scene()
        f = oracle.render(scene, view, w, w, 0, spp, num_meshes=n)
        films.append(f[..., :3] / f[..., 3:])
    assert np.allclose(films[0], films[1], rtol=2e-3, atol=1e-5)
