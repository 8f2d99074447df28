"""Image-mapped sky dome on the host and in the oracle (no GPU): su_image_create, the emission-map JSON, the Distribution2D
built by the scene compile (light_material.zig:54-119, distribution_1d.zig, distribution_2d.zig), Canopy sampling and the
light tree's infinite-light distribution (light_tree.zig:346-381, 449-462).

The reference holds no vectors for this path, so the pins are a numpy restatement of the inverse-CDF
lookup, sampling/pdf consistency, and two radiometric identities: a constant sky of radiance L shows L, and a diffuse ground
of albedo a under it shows a * L times the Substitute lobe's directional albedo."""

import numpy as np
import pytest

import oracle_lib as oracle
from zyg_b200 import scenes, su


@pytest.fixture()
def engine():
    su.release()
    yield
    su.release()


def test_image_create_formats(engine):
    su.init()
    rgb = np.random.default_rng(1).random((8, 16, 3), np.float32)
    a = su.image_create(rgb)
    b = su.image_create((rgb * 255).astype(np.uint8))
    assert (a, b) == (0, 1)  # ids count up from 0 like Cache.store
    L = su._su()
    px = np.zeros((4, 4, 4), np.float32)
    assert -1 == L.su_image_create(0xFFFFFFFF, 4, 4, 4, 4, 1, 16, px.ctypes.data)  # Float4: not an emission format here
    assert -1 == L.su_image_create(0xFFFFFFFF, 3, 3, 4, 4, 1, 6, px.ctypes.data)  # Float16
    assert 0 == L.su_image_update(a, 12, rgb.ctypes.data)
    assert -1 == L.su_image_update(17, 12, rgb.ctypes.data)


def test_missing_image_fails_compile(engine):
    su.init()
    su.perspective_camera_create(16, 16)
    su.integrators_create({"surface": {"PTMIS": {}}})
    m = su.material_create({"rendering": {"Light": {"emittance": {"emission_map": {"id": 5}, "value": 1.0}}}})
    sky = su.prop_create(su.CANOPY, [m])
    su.light_create(sky)
    with pytest.raises(su.SuError):
        su.compile_scene()


def _sky(size=64, **kw):
    scenes.sky_scene(32, 32, spp=4, sky_size=size, **kw)
    return su.compile_scene()


def test_distribution_sampling_matches_numpy_inverse_cdf(engine):
    """ImageImpl.sample over the compiled cdf rows == searchsorted on the same rows (the reference's LUT + linear search is
    only an accelerator), and the continuous offset / pdf follow distribution_1d.zig:62-75."""
    size = 64
    scene, _ = _sky(size)
    image = scenes.procedural_sky(size)
    # the host's luminance: hmax3 of the texel times Canopy.uvWeight, MIS-compensated (light_material.zig:248-272)
    c = (np.arange(size, dtype=np.float32) + np.float32(0.5)) * np.float32(1.0 / size)
    u, v = np.meshgrid(c, c)
    dx, dy = np.float32(2) * u - np.float32(1), np.float32(2) * v - np.float32(1)
    weight = ((dx * dx + dy * dy) <= 1).astype(np.float32)
    lum = image.max(-1) * weight
    avg = (image * weight[..., None]).sum((0, 1), dtype=np.float64) / weight.sum(dtype=np.float64)
    al = 0.6 * avg.max()
    lum = np.maximum(lum - np.float32(al), np.minimum(lum, np.float32(0.0025)))
    cond = np.cumsum(lum.astype(np.float64), 1) / lum.sum(1, dtype=np.float64)[:, None]
    marg = np.cumsum(lum.sum(1, dtype=np.float64)) / lum.sum(dtype=np.float64)

    rng = np.random.default_rng(7)
    r2 = rng.random((20000, 2)).astype(np.float32)
    got = oracle.image_sample(scene, 0, r2)
    row = np.minimum(np.searchsorted(marg, r2[:, 1].astype(np.float64), side="left"), size - 1)
    got_row = np.minimum((got[:, 1] * size).astype(np.int64), size - 1)
    # float32 cdf (fma-accumulated) vs float64 cumsum: a draw within an ulp of a cdf entry may land one texel over
    assert (got_row != row).mean() < 2e-3
    same = got_row == row
    col = np.minimum(np.array([np.searchsorted(cond[y], r, side="left") for y, r in zip(row, r2[:, 0].astype(np.float64))]), size - 1)
    got_col = np.minimum((got[:, 0] * size).astype(np.int64), size - 1)
    assert (got_col[same] != col[same]).mean() < 2e-3

    # pdf of the sampled point == pdf the sample carries (same texel, same products), pdf > 0 inside the disk only
    pdf = oracle.image_pdf(scene, 0, got[:, :2])
    inner = same & (got_col == col)
    close = np.isclose(pdf[inner], got[inner, 2], rtol=1e-5)
    assert close.mean() > 0.995  # a sample sitting exactly on a texel border may read the neighbour's pdf
    d = 2 * got[:, :2] - 1
    assert ((d * d).sum(1) <= 1.0 + 4.0 / size).all() and (got[:, 2] > 0).all()

    # ImageImpl.pdf = texel probability * total_weight: summed over the texels it gives total_weight (= texels inside the disk)
    centres = np.stack([u.ravel(), v.ravel()], 1)
    total = oracle.image_pdf(scene, 0, centres).astype(np.float64).sum()
    assert abs(total / weight.sum(dtype=np.float64) - 1.0) < 1e-4


def test_stochastic_bilinear_lookup(engine):
    """LinearStochastic2D (texture_sampler.zig:126-170): the expectation over the stochastic draw is the bilinear filter,
    texel centres return the texel, clamped addressing holds the border."""
    size = 16
    scene, _ = _sky(size)
    image = scenes.procedural_sky(size)
    c = (np.arange(size, dtype=np.float32) + 0.5) / size
    u, v = np.meshgrid(c, c)
    centre = np.stack([u.ravel(), v.ravel(), np.full(size * size, 0.37, np.float32)], 1)
    assert np.array_equal(oracle.image_texel(scene, 0, centre).reshape(size, size, 3), image)

    # between four texels: mean over r of the picked texel = bilinear weights
    x, y = 5, 9
    fu, fv = 0.3, 0.6
    uu, vv = (x + 0.5 + fu) / size, (y + 0.5 + fv) / size
    r = (np.arange(4096, dtype=np.float32) + 0.5) / 4096
    got = oracle.image_texel(scene, 0, np.stack([np.full_like(r, uu), np.full_like(r, vv), r], 1)).astype(np.float64).mean(0)
    want = ((1 - fu) * (1 - fv) * image[y, x] + fu * (1 - fv) * image[y, x + 1] + (1 - fu) * fv * image[y + 1, x]
            + fu * fv * image[y + 1, x + 1])
    assert np.allclose(got, want, rtol=2e-3, atol=1e-5)

    edge = oracle.image_texel(scene, 0, np.array([[1.7, -0.4, 0.9], [-3.0, 0.5, 0.1]], np.float32))
    assert np.array_equal(edge[0], image[0, size - 1])
    assert np.array_equal(edge[1], image[size // 2 - 1, 0]) or np.array_equal(edge[1], image[size // 2, 0])


def test_constant_sky_furnace(engine):
    """A constant sky of radiance L over an infinite diffuse ground: escaping camera rays show exactly L; the ground shows
    a * L * (directional albedo of the roughness-1 Substitute lobe, within a few per cent of 1)."""
    w, spp, L, a = 48, 128, 2.0, 0.5
    scenes.sky_scene(w, w, spp=spp, uniform_sky=L, sun=None, objects=False, ground_albedo=a, max_depth=4, sky_size=32)
    scene, view = su.compile_scene()
    film = oracle.render(scene, view, w, w, 0, spp)
    img = film[..., :3] / film[..., 3:4]
    assert np.array_equal(img[:8], np.full((8, w, 3), L, np.float32))
    ground = img[-12:].astype(np.float64)
    assert abs(ground.mean() / (a * L) - 1.0) < 0.02
    assert ground.std() / ground.mean() < 0.03


def test_infinite_light_distribution_agrees_with_split(engine):
    """Sky + sun = two infinite lights. With a split threshold the tree returns both (pdf 1 each); without, it picks one by
    power through Tree.infinite_light_distribution. Both are unbiased estimators of the same image."""
    w, spp = 64, 256
    means = []
    for st in (0.5, 0.0):
        su.release()
        scenes.sky_scene(w, w, spp=spp, sun=8.0, split_threshold=st, max_depth=4, sky_size=64)
        scene, view = su.compile_scene()
        picks = oracle.light_tree_random(scene, view, (0.0, 0.1, 0.0), (0.0, 1.0, 0.0), 0.3, 0.5 ** 4 if st > 0 else 0.0)
        if st > 0:
            assert [p[1] for p in picks] == [1.0, 1.0] and sorted(int(p[0]) for p in picks) == [0, 1]
        else:
            assert 1 == len(picks)
            total = sum(oracle.light_tree_pdf(scene, view, (0.0, 0.1, 0.0), (0.0, 1.0, 0.0), 0.0, l) for l in (0, 1))
            assert abs(total - 1.0) < 1e-6
        film = oracle.render(scene, view, w, w, 0, spp)
        means.append((film[..., :3] / film[..., 3:4]).astype(np.float64).mean((0, 1)))
    assert np.allclose(means[0], means[1], rtol=0.01)


def test_sphere_light_irradiance(engine):
    """Sphere.sampleTo / pdf / emission pinned against the closed form: a diffuse sphere emitter of radiance L and radius r
    whose centre is d above a diffuse plane gives the point below it the irradiance pi * L * (r / d)^2, i.e. the radiance
    a * L * (r / d)^2 (times the Substitute lobe's directional albedo, within 2 % of 1)."""
    L, a, r, d = 5.0, 0.5, 0.5, 2.0
    for unoccluding in (False, True):
        su.release()
        su.init()
        camera = su.perspective_camera_create(32, 32)
        su.camera_set_fov(float(np.radians(2.0)))
        su.prop_set_transformation(camera, su.transformation(position=(0.0, 1.0, -3.0), rotation_deg=(-18.434949, 0.0, 0.0)))
        su.sampler_create(256)
        su.integrators_create({"surface": {"PTMIS": {"depth": {"surface": 1}}}})
        su.sensor_create({})
        ground = su.material_create({"rendering": {"Substitute": {"color": [a] * 3, "roughness": 1.0}}})
        g = su.prop_create(su.RECTANGLE, [ground])
        su.prop_set_transformation(g, su.transformation((0.0, 0.0, 0.0), (100.0, 100.0, 1.0), (90.0, 0.0, 0.0)))
        lamp_material = su.material_create({"rendering": {"Light": {"emittance": {"value": L}}}})
        lamp = su.prop_create(su.SPHERE, [lamp_material], unoccluding=unoccluding)
        su.prop_set_transformation(lamp, su.transformation((0.0, d, 0.0), (2 * r, 2 * r, 2 * r)))
        su.light_create(lamp)
        scene, view = su.compile_scene()
        film = oracle.render(scene, view, 32, 32, 0, 256)
        img = (film[..., :3] / film[..., 3:4]).astype(np.float64)
        # sRGB (1,1,1) -> AP1 keeps grey: the film holds AP1 radiance; centre pixels look at the origin
        got = img[12:20, 12:20].mean()
        assert abs(got / (a * L * (r / d) ** 2) - 1.0) < 0.03, (unoccluding, got)


def test_constant_colour_map_equals_uniform_colour(engine, monkeypatch):
    """Substitute colour maps (substitute_material.zig:120): images that hold one colour everywhere give the film of the
    uniform-colour materials (the grey passes sRGB -> AP1 within an ulp), with either texture filter."""
    films = {}
    for key, nearest in (("linear", False), ("nearest", True)):
        su.release()
        monkeypatch.setattr(scenes, "checker_image", lambda size=64, cells=8, a=None, b=None, dtype=np.float32:
                            np.full((size, size, 3), 0.5, np.float32))
        n = scenes.textured_scene(64, 64, spp=8, nearest=nearest)
        scene, view = su.compile_scene()
        films[key] = oracle.render(scene, view, 64, 64, 0, 8, num_meshes=n)
    monkeypatch.undo()
    su.release()
    n = scenes.textured_scene(64, 64, spp=8, uniform=(0.5, 0.5, 0.5))
    scene, view = su.compile_scene()
    plain = oracle.render(scene, view, 64, 64, 0, 8, num_meshes=n)
    assert np.array_equal(films["linear"], films["nearest"])
    assert np.allclose(films["linear"], plain, rtol=1e-4, atol=1e-6)


def test_image_update_reaches_the_next_frame(engine):
    """su_image_update (capi.zig:300-340) overwrites the library's copy of the pixels; the next compile rebuilds the
    distribution from them."""
    w, spp = 32, 8
    su.init()
    camera = su.perspective_camera_create(w, w)
    su.prop_set_transformation(camera, su.transformation(rotation_deg=(60.0, 0.0, 0.0)))  # look up
    su.sampler_create(spp)
    su.integrators_create({"surface": {"PTMIS": {"depth": {"surface": 2}}}})
    su.sensor_create({})
    pixels = np.full((16, 16, 3), 2.0, np.float32)
    scenes.add_sky(pixels)
    scene, view = su.compile_scene()
    a = oracle.render(scene, view, w, w, 0, spp)
    assert np.array_equal(a[..., :3], np.full((w, w, 3), 2.0 * spp, np.float32))
    su.image_update(0, np.full((16, 16, 3), 0.25, np.float32))
    scene, view = su.compile_scene()
    b = oracle.render(scene, view, w, w, 0, spp)
    assert np.array_equal(b[..., :3], np.full((w, w, 3), 0.25 * spp, np.float32))


def test_image_mapped_rectangle_light_agrees_with_uniform_light(engine):
    """Rectangle.sampleMaterialTo / materialPdf with a constant emission image is an ordinary area light sampled through the
    image's (then uniform) Distribution2D instead of the spherical rectangle: both estimators converge to the same image."""
    w, spp = 48, 256
    means = []
    for image in (np.full((8, 8, 3), 1.0, np.float32), False):
        su.release()
        scenes.image_light_scene(w, w, spp=spp, image=image)
        scene, view = su.compile_scene()
        film = oracle.render(scene, view, w, w, 0, spp)
        img = (film[..., :3] / film[..., 3:4]).astype(np.float64)
        means.append(img.reshape(6, 8, 6, 8, 3).mean((1, 3)))
    assert np.allclose(means[0], means[1], rtol=0.04)
    assert abs(means[0].mean() / means[1].mean() - 1.0) < 0.01
