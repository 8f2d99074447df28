"""Parity of the device's wavefront PathtracerMIS with the CPU oracle on BASELINE config 1 (Cornell box), through
zyg's C API (su_*). The oracle is the checker only.

Bar (BASELINE.json north_star): renders statistically indistinguishable from the CPU path, RMSE at the tested spp
within 1 % of the CPU path's RMSE. Because the device keeps every path on the CPU path's sampler dimensions
(device/render.cuh), the comparison can be much tighter than statistical: per-pixel agreement to fp32 rounding of the
libm-dependent functions (sin / cos / acos differ by an ulp between glibc and CUDA), except for the few paths where
such an ulp flips a discrete decision (Russian roulette, a hit at a silhouette edge)."""

import ctypes as C
import os

import numpy as np
import pytest

import oracle_lib as oracle
from zyg_b200 import lib, scenes, su

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def download_film(width, height):
    L = lib.load_library()
    L.zygpu_download_film.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
    film = np.zeros((height, width, 4), np.float32)
    assert 0 == L.zygpu_download_film(su.device_handle(), film.ctypes.data, width * height), L.zygpu_last_error()
    return film


def rel_error(a, b):
    d = np.abs(a[..., :3] - b[..., :3]).sum(-1)
    return d / np.maximum(np.abs(b[..., :3]).sum(-1), 1e-6)


@pytest.fixture()
def engine():
    su.release()
    yield
    su.release()


@pytest.mark.parametrize("filter_name", [None, "Mitchell", "Blackman"])
def test_cornell_matches_oracle_per_pixel(engine, filter_name):
    w, spp = 128, 16
    scenes.cornell_box(w, w, spp=spp, filter_name=filter_name)
    scene, view = su.compile_scene()
    ref = oracle.render(scene, view, w, w, 0, spp)
    su.render_frame(0)
    gpu = download_film(w, w)

    if filter_name is None:
        assert np.array_equal(gpu[..., 3], ref[..., 3])  # weights: exactly spp
    else:
        assert np.abs(gpu[..., 3] - ref[..., 3]).max() < 1e-3 * spp
    rel = rel_error(gpu, ref)
    assert np.median(rel) < 2e-6
    assert (rel > 1e-3).mean() < 2e-3, "more than 0.2 % of the pixels carry a path that diverged"
    assert abs(gpu[..., :3].mean() - ref[..., :3].mean()) / ref[..., :3].mean() < 1e-4


def test_cornell_full_config(engine):
    """BASELINE configs[0]: 512 x 512, 64 spp, max 8 bounces."""
    w, spp = 512, 64
    scenes.cornell_box(w, w, spp=spp)
    scene, view = su.compile_scene()
    ref = oracle.render(scene, view, w, w, 0, spp)
    su.render_frame(0)
    gpu = download_film(w, w)
    assert np.array_equal(gpu[..., 3], ref[..., 3])
    rel = rel_error(gpu, ref)
    assert np.median(rel) < 2e-6 and (rel > 1e-3).mean() < 1e-3
    assert abs(gpu[..., :3].mean() - ref[..., :3].mean()) / ref[..., :3].mean() < 1e-5

    rgba = su.resolve_frame_to_buffer(w, w)
    want = oracle.resolve(view, gpu)
    assert np.allclose(rgba, want, rtol=1e-6, atol=1e-7)


def test_rmse_against_high_spp_reference(engine):
    """north_star: "per-pixel mean relative error against a 16k-spp reference must converge at the reference's own rate, with
    RMSE at the tested spp within 1 % of the reference's RMSE". RMSE(device) within 1 % of RMSE(CPU path) at 16, 64 and 256 spp
    against a 16 384-spp CPU reference, the error falls at the CPU path's rate, and so does the mean relative error."""
    w = 64
    scenes.cornell_box(w, w, spp=16384 + 256)
    scene, view = su.compile_scene()
    truth = oracle.render(scene, view, w, w, 256, 16384)  # samples disjoint from the ones under test
    truth = truth[..., :3] / truth[..., 3:4]

    def rmse(film):
        img = film[..., :3] / film[..., 3:4]
        return float(np.sqrt(np.mean((img - truth) ** 2)))

    def mre(film):  # per-pixel mean relative error
        img = film[..., :3] / film[..., 3:4]
        return float(np.mean(np.abs(img - truth).sum(-1) / np.maximum(truth.sum(-1), 1e-4)))

    errs, rels = {}, {}
    for spp in (16, 64, 256):
        ref = oracle.render(scene, view, w, w, 0, spp)
        su.render_frame_range(0, 0, spp)
        gpu = download_film(w, w)
        errs[spp] = (rmse(gpu), rmse(ref))
        rels[spp] = (mre(gpu), mre(ref))
        assert abs(errs[spp][0] - errs[spp][1]) / errs[spp][1] < 0.01
        assert abs(rels[spp][0] - rels[spp][1]) / rels[spp][1] < 0.01
    for table in (errs, rels):
        slope_gpu = np.log(table[16][0] / table[256][0]) / np.log(16.0)
        slope_ref = np.log(table[16][1] / table[256][1]) / np.log(16.0)
        assert abs(slope_gpu - slope_ref) < 0.02 and slope_gpu > 0.35


def test_sample_ranges_accumulate_bit_exactly(engine):
    """su_start_frame + su_render_iterations (capi.zig:581-611) adds sample ranges into the film; because every
    sample is seeded from its absolute index the result is bit-identical to one su_render_frame."""
    w, spp = 96, 12
    scenes.cornell_box(w, w, spp=spp, filter_name="Mitchell")
    su.render_frame(0)
    whole = download_film(w, w)
    su.start_frame(0)
    for n in (1, 4, 7):
        su.render_iterations(n)
    su._ok(lib.load_library().zygpu_synchronize(su.device_handle()), "zygpu_synchronize")
    parts = download_film(w, w)
    assert whole.tobytes() == parts.tobytes()


def test_sample_range_split_sums_to_whole(engine):
    """The multi-GPU schedule: disjoint sample ranges rendered into separate films and summed equal
    the single-device film up to fp32 summation order."""
    w, spp = 96, 16
    scenes.cornell_box(w, w, spp=spp)
    su.render_frame(0)
    whole = download_film(w, w)
    total = np.zeros_like(whole)
    for g in range(4):
        su.render_frame_range(0, g * 4, 4)
        total += download_film(w, w)
    assert np.array_equal(total[..., 3], whole[..., 3])
    assert np.allclose(total, whole, rtol=2e-6, atol=1e-6)


def test_small_pass_size_gives_identical_film(engine, monkeypatch):
    """The number of samples traced per pass is a scheduling choice; it must not change the film."""
    w, spp = 64, 8
    scenes.cornell_box(w, w, spp=spp)
    su.render_frame(0)
    a = download_film(w, w)
    monkeypatch.setenv("ZYGPU_PATHS_PER_PASS", str(w * w * 3))
    su.render_frame(0)
    b = download_film(w, w)
    assert a.tobytes() == b.tobytes()


def test_depth_zero_and_random_sampler(engine):
    """Edge cases: max depth 1 (direct light only) and the Random sampler (sampler.zig:17-74) agree with the oracle."""
    w, spp = 64, 8
    scenes.cornell_box(w, w, spp=spp, max_depth=1)
    scene, view = su.compile_scene()
    ref = oracle.render(scene, view, w, w, 0, spp)
    su.render_frame(0)
    gpu = download_film(w, w)
    assert (rel_error(gpu, ref) > 1e-3).mean() < 2e-3


def test_mesh_scene_matches_oracle(engine):
    """A triangle-mesh prop (40 000-triangle displaced sphere) inside analytic props: prop tree -> 8-wide mesh BVH ->
    Mesh.fragment (triangle_mesh.zig:310-335) on the device against the reference-order binary tree on the CPU."""
    w, spp = 96, 8
    n = scenes.sphere_scene(w, w, spp=spp, quads=(200, 100))
    scene, view = su.compile_scene()
    ref = oracle.render(scene, view, w, w, 0, spp, num_meshes=n)
    su.render_frame(0)
    gpu = download_film(w, w)
    assert np.array_equal(gpu[..., 3], ref[..., 3])
    rel = rel_error(gpu, ref)
    assert np.median(rel) < 5e-6
    assert (rel > 1e-3).mean() < 5e-3
    assert abs(gpu[..., :3].mean() - ref[..., :3].mean()) / ref[..., :3].mean() < 2e-4


def test_metal_mesh_matches_oracle(engine):
    """Rough metal (Substitute, metallic 1): the GGX lobe + multi-scatter compensation path."""
    w, spp = 64, 8
    n = scenes.sphere_scene(w, w, spp=spp, quads=(100, 50), metallic=1.0, roughness=0.3)
    scene, view = su.compile_scene()
    ref = oracle.render(scene, view, w, w, 0, spp, num_meshes=n)
    su.render_frame(0)
    gpu = download_film(w, w)
    rel = rel_error(gpu, ref)
    assert np.median(rel) < 1e-5 and (rel > 1e-3).mean() < 1e-2


def test_instanced_scene_matches_oracle(engine):
    """su_prop_create_instance (capi.zig:457-469): 256 instances of 6 meshes with their own transformations, the three
    material kinds of BASELINE config 3 (diffuse, rough metal, glass with interior absorption) under a Rectangle light and
    a Distant sun; the prop tree has ~130 nodes, rays collect several mesh candidates, glass meshes split paths."""
    w, spp = 96, 8
    n = scenes.instanced_scene(w, w, spp=spp, grid=(16, 16), prototypes=6, quads=(40, 20), sun=60.0)
    scene, view = su.compile_scene()
    ref = oracle.render(scene, view, w, w, 0, spp, num_meshes=n, wavefront_light_order=True)
    su.render_frame(0)
    gpu = download_film(w, w)
    assert np.array_equal(gpu[..., 3], ref[..., 3])
    rel = rel_error(gpu, ref)
    assert np.median(rel) < 1e-5
    assert (rel > 1e-3).mean() < 1e-2
    assert abs(gpu[..., :3].mean() - ref[..., :3].mean()) / ref[..., :3].mean() < 5e-4


@pytest.mark.parametrize("roughness", [0.0, 0.25])
def test_glass_matches_oracle(engine, roughness):
    """Glass (glass_sample.zig): reflection + refraction split while max_splits allows (up to 4 vertices per camera
    sample, processed in the reference's pool order so they draw the same sampler dimensions), the medium stack with
    two nested dielectrics of different priority, Beer-Lambert absorption inside (volume_integrator.zig:51-66)."""
    w, spp = 128, 16
    scenes.cornell_box(w, w, spp=spp, glass={"roughness": roughness})
    scene, view = su.compile_scene()
    ref = oracle.render(scene, view, w, w, 0, spp)
    su.render_frame(0)
    gpu = download_film(w, w)
    assert np.array_equal(gpu[..., 3], ref[..., 3])
    rel = rel_error(gpu, ref)
    assert np.median(rel) < 5e-6
    assert (rel > 1e-3).mean() < 5e-3
    assert abs(gpu[..., :3].mean() - ref[..., :3].mean()) / ref[..., :3].mean() < 2e-4


@pytest.mark.parametrize("num_lights,split_threshold", [(64, 0.5), (400, 0.5), (9, 0.0)])
def test_many_lights_match_oracle(engine, num_lights, split_threshold):
    """The light tree the host builds (light_tree_builder.zig:281-376) over many Rectangle lights: stochastic descent,
    adaptive splitting with several picks per vertex, Tree.pdf for the MIS weight of emitter hits. A vertex with several
    light samples takes its sampler draws in the wavefront's order (zyg_oracle.h: zo_set_wavefront_light_order), so the
    oracle is asked for that order; tests/test_light_tree_host.py shows the two orders are statistically equivalent.
    Tolerance: the solid angle of a small, far spherical rectangle is a difference of acos sums (rectangle.zig:244-262),
    which amplifies the last ulp of libm's acos into ~1e-4 of the light pdf."""
    w, spp = 96, 16
    scenes.many_lights_scene(w, w, spp=spp, num_lights=num_lights, split_threshold=split_threshold)
    scene, view = su.compile_scene()
    ref = oracle.render(scene, view, w, w, 0, spp, wavefront_light_order=True)
    su.render_frame(0)
    gpu = download_film(w, w)
    assert np.array_equal(gpu[..., 3], ref[..., 3])
    rel = rel_error(gpu, ref)
    assert np.median(rel) < 1e-4
    assert (rel > 1e-2).mean() < 1e-2
    assert abs(gpu[..., :3].mean() - ref[..., :3].mean()) / ref[..., :3].mean() < 1e-4


def test_distant_light_matches_oracle(engine):
    """A Distant light (distant.zig:22-146) next to a Rectangle light: the infinite-light branch of the light tree
    (light_tree.zig:353-371), shadow rays that run to RayMaxT (shape.zig:401-403), emission met on escape through
    Scene.infinite_props (pathtracer_mis.zig:313-338) weighted with Distant.pdf."""
    w, spp = 96, 16
    n = scenes.sphere_scene(w, w, spp=spp, quads=(100, 50), sun=150.0)
    scene, view = su.compile_scene()
    ref = oracle.render(scene, view, w, w, 0, spp, num_meshes=n, wavefront_light_order=True)
    su.render_frame(0)
    gpu = download_film(w, w)
    assert np.array_equal(gpu[..., 3], ref[..., 3])
    rel = rel_error(gpu, ref)
    assert np.median(rel) < 5e-6 and (rel > 1e-3).mean() < 5e-3
    assert abs(gpu[..., :3].mean() - ref[..., :3].mean()) / ref[..., :3].mean() < 1e-4
    # the sun matters: the same scene without it is darker
    su.release()
    scenes.sphere_scene(w, w, spp=spp, quads=(100, 50))
    scene, view = su.compile_scene()
    dark = oracle.render(scene, view, w, w, 0, 4, num_meshes=n)
    assert ref[..., :3].sum() / ref[..., 3].sum() > 1.2 * dark[..., :3].sum() / dark[..., 3].sum()


@pytest.mark.parametrize("split_threshold,unoccluding", [(0.5, False), (0.0, False), (0.5, True)])
def test_mesh_lights_match_oracle(engine, split_threshold, unoccluding):
    """Emissive triangle meshes (config 4 style): 24 icosahedra + one 576-triangle emitter. Part.configure and the
    per-part PrimitiveTree on the host (triangle_mesh.zig:57-149, light_tree_builder.zig:378-428), Mesh.sampleTo with
    Arvo's spherical-triangle sampling near and area sampling far (triangle_mesh.zig:402-608), Mesh.pdf through the
    primitive mapping for emitter hits (:662-703). Un-occluding mesh emitters (what a scene file's Light entities are by
    default) are gathered by the all-hits traversal of TriangleTree.emission (triangle_tree.zig:405-477)."""
    w, spp = 96, 16
    n = scenes.mesh_lights_scene(w, w, spp=spp, split_threshold=split_threshold, unoccluding=unoccluding)
    scene, view = su.compile_scene()
    ref = oracle.render(scene, view, w, w, 0, spp, num_meshes=n, wavefront_light_order=True)
    su.render_frame(0)
    gpu = download_film(w, w)
    assert np.array_equal(gpu[..., 3], ref[..., 3])
    rel = rel_error(gpu, ref)
    assert np.median(rel) < 2e-5
    assert (rel > 1e-2).mean() < 1e-2
    assert abs(gpu[..., :3].mean() - ref[..., :3].mean()) / ref[..., :3].mean() < 1e-4


@pytest.mark.parametrize("split_threshold,sun", [(0.5, 8.0), (0.0, 8.0), (0.5, None)])
def test_sky_dome_matches_oracle(engine, split_threshold, sun):
    """Image-mapped Canopy sky (a PropImage light: su_image_create + emission_map, Distribution2D importance sampling,
    stochastic bilinear lookups; canopy.zig:27-131, shape_sampler.zig:128-152, texture_sampler.zig:126-170) with and without
    a Distant sun. Two infinite lights with split threshold 0 go through Tree.infinite_light_distribution
    (light_tree.zig:366-374, 456-461). The device finds the cdf entry by bisection, the oracle by the reference's lookup table
    and linear walk."""
    w, spp = 128, 16
    scenes.sky_scene(w, w, spp=spp, sun=sun, split_threshold=split_threshold, sky_size=256)
    scene, view = su.compile_scene()
    ref = oracle.render(scene, view, w, w, 0, spp, wavefront_light_order=True)
    su.render_frame(0)
    gpu = download_film(w, w)
    assert np.array_equal(gpu[..., 3], ref[..., 3])
    rel = rel_error(gpu, ref)
    # acos / atan2 / sin / cos of the equidistant mapping differ by an ulp between glibc and CUDA: a sample next to a texel
    # border can read the neighbouring texel
    assert np.median(rel) < 5e-6
    assert (rel > 1e-2).mean() < 1e-2
    assert abs(gpu[..., :3].mean() - ref[..., :3].mean()) / ref[..., :3].mean() < 2e-4


def test_constant_sky_furnace_on_device(engine):
    """Size-independent property at full resolution: a constant sky of radiance L shows exactly L where the camera sees it
    and a * L (times the diffuse lobe's albedo) on an infinite ground."""
    w, h, spp, L, a = 1920, 1080, 4, 2.0, 0.5
    scenes.sky_scene(w, h, spp=spp, uniform_sky=L, sun=None, objects=False, ground_albedo=a, max_depth=4, sky_size=64)
    su.render_frame(0)
    gpu = download_film(w, h)
    img = gpu[..., :3] / gpu[..., 3:4]
    assert np.array_equal(img[:200], np.full((200, w, 3), L, np.float32))
    ground = img[-300:].astype(np.float64)
    assert abs(ground.mean() / (a * L) - 1.0) < 0.02


def test_mesh_lights_with_sky_match_oracle(engine):
    """Config 4 in small: emissive meshes + Distant sun + image-mapped sky, light selection and sampling in the deferred
    persistent light kernels."""
    w, spp = 96, 8
    n = scenes.mesh_lights_scene(w, w, spp=spp, sun=15.0, sky=128)
    scene, view = su.compile_scene()
    ref = oracle.render(scene, view, w, w, 0, spp, num_meshes=n, wavefront_light_order=True)
    su.render_frame(0)
    gpu = download_film(w, w)
    assert np.array_equal(gpu[..., 3], ref[..., 3])
    rel = rel_error(gpu, ref)
    assert np.median(rel) < 2e-5
    assert (rel > 1e-2).mean() < 1e-2
    assert abs(gpu[..., :3].mean() - ref[..., :3].mean()) / ref[..., :3].mean() < 2e-4


@pytest.mark.parametrize("split_threshold,num_samples,deferred", [(0.5, 1, "0"), (0.5, 4, "0"), (0.0, 1, "0"), (0.5, 4, "1")])
def test_sphere_lights_match_oracle(engine, monkeypatch, split_threshold, num_samples, deferred):
    """Sphere lights (cone sampling of the visible cap, sphere.zig:323-393; pdf :472-487) as occluding props and as
    un-occluding emitters gathered by Sphere.emission (:271-279), sampled inside shade_a and in the deferred light kernels."""
    monkeypatch.setenv("ZYGPU_DEFERRED_LIGHTS", deferred)
    w, spp = 128, 16
    scenes.sphere_lights_scene(w, w, spp=spp, split_threshold=split_threshold, num_samples=num_samples)
    scene, view = su.compile_scene()
    ref = oracle.render(scene, view, w, w, 0, spp, wavefront_light_order=True)
    su.render_frame(0)
    gpu = download_film(w, w)
    assert np.array_equal(gpu[..., 3], ref[..., 3])
    rel = rel_error(gpu, ref)
    assert np.median(rel) < 5e-6
    assert (rel > 1e-2).mean() < 5e-3
    assert abs(gpu[..., :3].mean() - ref[..., :3].mean()) / ref[..., :3].mean() < 2e-4


def test_instancer_matches_oracle(engine):
    """A scene-file style Instancer entity (zyg_su_instancer_create) under a rotated / scaled / translated transformation:
    flattened on the host into the two-level layout, traced and shaded like prop instances."""
    w, spp = 128, 8
    outer = su.transformation((0.4, 0.3, -0.2), (1.25, 1.25, 1.25), (0.0, 35.0, 0.0))
    n = scenes.instanced_scene(w, w, spp=spp, grid=(16, 16), prototypes=3, quads=(40, 20), instancer=outer, sun=30.0)
    scene, view = su.compile_scene()
    ref = oracle.render(scene, view, w, w, 0, spp, num_meshes=n, wavefront_light_order=True)
    su.render_frame(0)
    gpu = download_film(w, w)
    assert np.array_equal(gpu[..., 3], ref[..., 3])
    rel = rel_error(gpu, ref)
    assert np.median(rel) < 5e-6
    assert (rel > 1e-2).mean() < 1e-2
    assert abs(gpu[..., :3].mean() - ref[..., :3].mean()) / ref[..., :3].mean() < 5e-4


@pytest.mark.parametrize("nearest", [False, True])
def test_colour_maps_match_oracle(engine, nearest):
    """Substitute colour maps on a Rectangle (Repeat, scaled), a Cube (sRGB bytes, Clamp) and a triangle mesh (its uvs):
    stochastic-bilinear / nearest texel lookups with the vertex's stochastic_r (texture_sampler.zig:99-170), the same texel in
    shade_a and shade_b."""
    w, spp = 128, 16
    n = scenes.textured_scene(w, w, spp=spp, nearest=nearest)
    scene, view = su.compile_scene()
    ref = oracle.render(scene, view, w, w, 0, spp, num_meshes=n)
    su.render_frame(0)
    gpu = download_film(w, w)
    assert np.array_equal(gpu[..., 3], ref[..., 3])
    rel = rel_error(gpu, ref)
    assert np.median(rel) < 5e-6
    assert (rel > 1e-2).mean() < 5e-3
    assert abs(gpu[..., :3].mean() - ref[..., :3].mean()) / ref[..., :3].mean() < 2e-4


def test_empty_and_sky_only_scenes(engine):
    """Edge cases of the scene compile and the stages: no props at all (every path escapes into nothing: a black film with
    full weights), and a sky dome over no geometry (every pixel shows the sky's radiance, no shadow rays, empty prop trees)."""
    w, spp = 64, 4
    su.init()
    su.perspective_camera_create(w, w)
    su.sampler_create(spp)
    su.integrators_create({"surface": {"PTMIS": {"depth": {"surface": 4}}}})
    su.sensor_create({})
    su.render_frame(0)
    gpu = download_film(w, w)
    assert np.array_equal(gpu[..., 3], np.full((w, w), spp, np.float32)) and not gpu[..., :3].any()

    su.release()
    su.init()
    camera = su.perspective_camera_create(w, w)
    su.prop_set_transformation(camera, su.transformation(rotation_deg=(60.0, 0.0, 0.0)))  # look up
    su.sampler_create(spp)
    su.integrators_create({"surface": {"PTMIS": {"depth": {"surface": 4}}}})
    su.sensor_create({})
    scenes.add_sky(np.full((32, 32, 3), 1.5, np.float32))
    scene, view = su.compile_scene()
    ref = oracle.render(scene, view, w, w, 0, spp)
    su.render_frame(0)
    gpu = download_film(w, w)
    assert np.array_equal(gpu, ref)
    assert np.array_equal(gpu[..., :3], np.full((w, w, 3), 1.5 * spp, np.float32))


@pytest.mark.parametrize("num_samples,two_sided,deferred", [(1, False, "0"), (3, True, "0"), (2, False, "1")])
def test_image_mapped_rectangle_light_matches_oracle(engine, monkeypatch, num_samples, two_sided, deferred):
    """A Rectangle light with an emission image (PropImage class on a finite shape): Rectangle.sampleMaterialTo / materialPdf,
    the sample's uv carried in the shadow record for Light.evaluateTo, inline and deferred light sampling."""
    monkeypatch.setenv("ZYGPU_DEFERRED_LIGHTS", deferred)
    w, spp = 128, 16
    scenes.image_light_scene(w, w, spp=spp, num_samples=num_samples, two_sided=two_sided)
    scene, view = su.compile_scene()
    ref = oracle.render(scene, view, w, w, 0, spp, wavefront_light_order=True)
    su.render_frame(0)
    gpu = download_film(w, w)
    assert np.array_equal(gpu[..., 3], ref[..., 3])
    rel = rel_error(gpu, ref)
    assert np.median(rel) < 5e-6
    assert (rel > 1e-2).mean() < 5e-3
    assert abs(gpu[..., :3].mean() - ref[..., :3].mean()) / ref[..., :3].mean() < 2e-4


@pytest.mark.parametrize("builder,kwargs,spp", [
    ("instanced_scene", {"grid": (100, 100), "prototypes": 20, "quads": (500, 250), "sun": 60.0}, 4),
    ("mesh_lights_scene", {"num_lights": 1000, "geometry_quads": (400, 250), "sun": 15.0, "sky": 1024, "max_depth": 8}, 2),
])
def test_full_size_configs_by_properties(engine, builder, kwargs, spp):
    """BASELINE configs 3 and 4 at their full size (5 M triangles / 10 k instances; 1 k mesh lights + sky + sun; 1920 x 1080),
    where the oracle would take minutes: size-independent properties instead. Sample ranges added into the film reproduce the
    whole frame bit for bit (every sample is seeded from its absolute index, equal-t ties are schedule-independent), the weights are
    exactly the sample count, the film is finite and lit, and the two halves of the sample range agree statistically."""
    w, h = 1920, 1080
    getattr(scenes, builder)(w, h, spp=spp, **kwargs)
    su.render_frame(0)
    whole = download_film(w, h)
    assert np.array_equal(whole[..., 3], np.full((h, w), spp, np.float32))
    assert np.isfinite(whole).all() and whole[..., :3].min() >= 0.0 and whole[..., :3].mean() > 0.0

    su.start_frame(0)
    su.render_iterations(spp // 2)
    su._ok(lib.load_library().zygpu_synchronize(su.device_handle()), "zygpu_synchronize")
    first = download_film(w, h)
    su.render_iterations(spp - spp // 2)
    su._ok(lib.load_library().zygpu_synchronize(su.device_handle()), "zygpu_synchronize")
    parts = download_film(w, h)
    # Equal-t ties (a ray through an edge shared by two triangles) are resolved by primitive id, not by the order in which the
    # lock-step schedule happens to test them, so mesh scenes accumulate bit-identically too
    assert parts.tobytes() == whole.tobytes()

    second = whole[..., :3] - first[..., :3]
    a, b = first[..., :3].astype(np.float64).mean(), second.astype(np.float64).mean()
    assert abs(a - b) / (0.5 * (a + b)) < 0.05


def _compare(w, h, spp, num_meshes=0, wavefront=False, median=5e-6, tail=5e-3, mean=2e-4):
    scene, view = su.compile_scene()
    ref = oracle.render(scene, view, w, h, 0, spp, num_meshes=num_meshes, wavefront_light_order=wavefront)
    su.render_frame(0)
    gpu = download_film(w, h)
    assert np.array_equal(gpu[..., 3], ref[..., 3])
    rel = rel_error(gpu, ref)
    lit = ref[..., 3] > 0
    assert np.median(rel[lit]) < median
    assert (rel[lit] > 1e-3).mean() < tail
    assert abs(gpu[..., :3].mean() - ref[..., :3].mean()) / ref[..., :3].mean() < mean
    return gpu, ref


def test_thin_lens_matches_oracle(engine):
    """Perspective.generateVertex with a lens (camera_perspective.zig:124-150): Aperture.sample maps lens_uv to the disk
    (aperture.zig:46-53), the ray starts on the lens and passes through the point of the pinhole ray at the focus distance.
    The defocused film differs from the pinhole film and agrees with the oracle per pixel."""
    w, spp = 128, 16
    scenes.cornell_box(w, w, spp=spp)
    su.render_frame(0)
    pinhole = download_film(w, w)
    su.camera_set_lens(0.08, 3.9)  # focused on the centre of the box
    gpu, ref = _compare(w, w, spp)
    blurred = rel_error(gpu, pinhole)
    assert (blurred > 1e-2).mean() > 0.2, "the lens left most of the film unchanged"
    # in focus (the plane z = 0 through the light) sharp features stay: the film is not simply a blur of everything
    assert np.isfinite(gpu).all()


def test_sensor_clamp_matches_oracle(engine):
    """Sensor.clamp per component class before summation (sensor.zig:177-181, 615-624): emission, direct and indirect are
    scaled down separately where their largest channel exceeds the limit."""
    w, spp = 128, 16
    scenes.cornell_box(w, w, spp=spp)
    su.render_frame(0)
    free = download_film(w, w)
    su.sensor_create({"clamp": {"emission": 4.0, "direct": 0.4, "indirect": 0.1}})
    gpu, ref = _compare(w, w, spp)
    assert gpu[..., :3].sum() < 0.9 * free[..., :3].sum(), "the clamp removed no energy"
    lamp = free[..., :3].max(-1) > 16.0 * spp  # pixels that only see the lamp: emission 17 clamped to 4
    assert lamp.any() and np.allclose(gpu[lamp][:, :3].max(-1) / spp, 4.0, rtol=0.2)


def test_visibility_flags_match_oracle(engine):
    """Prop visibility by depth (prop.zig:38-48, 78-91): the tall box is hidden from the camera but still seen in reflections
    and still casts shadows; the short box is hidden everywhere. su_prop_set_visibility, capi.zig:535-546."""
    w, spp = 128, 16
    scenes.cornell_box(w, w, spp=spp, roughness=1.0)
    su.render_frame(0)
    both = download_film(w, w)
    tall, short = 6, 7  # entity ids: camera 0, walls 1-5, the two boxes, the lamp
    su.prop_set_visibility(tall, False, True)
    su.prop_set_visibility(short, False, False)
    gpu, ref = _compare(w, w, spp)
    changed = rel_error(gpu, both) > 1e-2
    assert changed.mean() > 0.05
    # where the tall box was seen directly the back wall shows now, still shadowed by the box's shadow-ray visibility
    su.prop_set_visibility(tall, False, False)
    su.render_frame(0)
    gone = download_film(w, w)
    assert (rel_error(gone, gpu) > 1e-2).mean() > 0.01, "shadow / reflection visibility of the hidden box had no effect"


def test_ptmis_without_light_sampling_node_uses_raw_default(engine):
    """loadLightSampling (take.zig:263-271): without a "light_sampling" node the split threshold is the raw 0.5, not 0.5^4 —
    the reference's own capi-test/test.py passes PTMIS that way. Many lights make the threshold visible in the film."""
    w, spp = 96, 8
    scenes.many_lights_scene(w, w, spp=spp, num_lights=64)
    su.integrators_create({"surface": {"PTMIS": {"depth": {"surface": 6}}}})
    gpu, ref = _compare(w, w, spp, wavefront=True, median=1e-4, tail=5e-2, mean=2e-4)
    su.integrators_create({"surface": {"PTMIS": {"depth": {"surface": 6}, "light_sampling": {"split_threshold": 0.5}}}})
    su.render_frame(0)
    explicit = download_film(w, w)
    assert (rel_error(explicit, gpu) > 1e-3).mean() > 0.05, "0.5 and 0.5^4 gave the same film"


def test_overflowed_pass_is_rerun_not_dropped(engine, monkeypatch):
    """A vertex that produces more light samples than a slot reserves used to drop them (a darker film, rc 0 in the progressive
    API). The pass is now discarded and run again with a larger reservation: the film equals the one of a generous reservation."""
    w, spp = 64, 4
    L = lib.load_library()

    class Stats(C.Structure):
        _fields_ = [(n, C.c_uint64) for n in ("camera_samples", "closest_rays", "shadow_rays", "kernel_launches", "passes", "overflow_retries")]

    L.zygpu_render_stats.argtypes = [C.c_void_p, C.POINTER(Stats)]
    films = {}
    for cap in ("2", "512"):
        su.release()
        monkeypatch.setenv("ZYGPU_MAX_LIGHT_SAMPLES", cap)
        scenes.mesh_lights_scene(w, w, spp=spp, num_lights=24)
        su.start_frame(0)
        su.render_iterations(spp)
        films[cap] = download_film(w, w)
        st = Stats()
        assert 0 == L.zygpu_render_stats(su.device_handle(), C.byref(st))
        assert (st.overflow_retries > 0) == ("2" == cap)
    assert films["2"].tobytes() == films["512"].tobytes() or np.allclose(films["2"], films["512"], rtol=1e-5, atol=1e-6)


def _crop_compare(w, h, crop, spp, num_meshes, median, tail):
    """Device vs oracle on the window `crop` of a w x h frame; pixel ids and seeds run over the full resolution."""
    su.camera_set_crop(*crop)
    su.sampler_create(spp)
    scene, view = su.compile_scene()
    ref = oracle.render(scene, view, w, h, 0, spp, num_meshes=num_meshes, wavefront_light_order=True)
    su.render_frame(0)
    gpu = download_film(w, h)
    x0, y0, x1, y1 = crop
    inside = np.zeros((h, w), bool)
    inside[y0:y1, x0:x1] = True
    assert not gpu[~inside].any() and not ref[~inside].any(), "samples landed outside the crop"
    g, r = gpu[y0:y1, x0:x1], ref[y0:y1, x0:x1]
    assert np.array_equal(g[..., 3], r[..., 3]) and (r[..., 3] == spp).all()
    rel = rel_error(g, r)
    assert np.median(rel) < median
    assert (rel > 1e-2).mean() < tail
    assert abs(g[..., :3].mean() - r[..., :3].mean()) / r[..., :3].mean() < 1e-3
    return g, r


def test_config3_and_config5_full_scene_crops_match_oracle(engine):
    """BASELINE configs 3 and 5 at their full scene size — 20 prototypes x 250 k = 5 M triangles, 10 000 instances, diffuse /
    rough metal / glass, Rectangle light + Distant sun — compared with the oracle per pixel on windows of the 1920 x 1080 and
    3840 x 2160 frames (the oracle needs minutes for a whole frame). A traversal bug that needs the full TLAS depth or many
    mesh candidates per ray shows here; the miniature scenes of the other tests cannot see it."""
    kwargs = {"grid": (100, 100), "prototypes": 20, "quads": (500, 250), "sun": 60.0}
    n = scenes.instanced_scene(1920, 1080, spp=2, **kwargs)
    # a window across the horizon of the instance field (grazing rays cross many instances) and one near the camera
    _crop_compare(1920, 1080, (640, 300, 1280, 480), 2, n, 1e-5, 2e-2)
    _crop_compare(1920, 1080, (200, 800, 520, 980), 2, n, 1e-5, 2e-2)
    su.perspective_camera_create(3840, 2160)  # config 5: the same scene at 4K
    su.camera_set_fov(float(np.radians(50.0)))
    _crop_compare(3840, 2160, (1700, 900, 2212, 1188), 1, n, 1e-5, 2e-2)


def test_config4_full_scene_crop_matches_oracle(engine):
    """BASELINE config 4 at full size: 1000 emissive icosahedra + a 576-triangle emitter + 200 k triangles of diffuse geometry,
    sky image (1024^2) + sun, light-tree sampling with split threshold 0.5, on a window of the 1920 x 1080 frame."""
    kwargs = {"num_lights": 1000, "geometry_quads": (400, 250), "sun": 15.0, "sky": 1024, "max_depth": 8}
    n = scenes.mesh_lights_scene(1920, 1080, spp=2, **kwargs)
    _crop_compare(1920, 1080, (700, 400, 1220, 680), 2, n, 5e-5, 3e-2)


def test_export_frame_writes_the_resolved_image(engine, tmp_path, monkeypatch):
    """su_exporters_create + su_export_frame (capi.zig:189-200, 569-579; driver.zig:224-253): one file per exporter named
    image_<camera:02>_<frame:06>.<ext>; the float EXR holds exactly what su_resolve_frame_to_buffer returns."""
    from test_image_writer_host import read_exr, read_png, read_rgbe

    w, h = 96, 64
    scenes.cornell_box(w, h, spp=4)
    monkeypatch.chdir(tmp_path)
    su.exporters_create({"Image": {"format": "EXR", "bitdepth": 32}})
    su.render_frame(7)
    su.export_frame()
    resolved = su.resolve_frame_to_buffer(w, h)
    exr, names, window, _ = read_exr("image_00_000007.exr")
    assert names == ["B", "G", "R"] and window == (0, 0, w - 1, h - 1)
    assert np.array_equal(exr, resolved[..., [2, 1, 0]])

    su.exporters_create({"Image": {"format": "PNG"}})
    su.export_frame()
    png = read_png("image_00_000007.png")
    assert png.shape == (h, w, 3) and png.mean() > 5
    srgb = np.zeros((h, w, 3), np.uint8)
    su._su().su_resolve_frame(0xFFFFFFFF)
    su._su().su_copy_framebuffer(0, 3, w, h, srgb.ctypes.data)
    assert np.array_equal(png, srgb)

    su.exporters_create({"Image": {"format": "RGBE"}})
    su.export_frame()
    assert read_rgbe("image_00_000007.hdr").shape == (h, w, 4)


@pytest.mark.parametrize("nearest", [False, True])
def test_surface_maps_match_oracle(engine, nearest):
    """Substitute roughness, metallic and normal maps (substitute_material.zig:122-123, 157-159; hlp.sampleNormal,
    material_helper.zig:16-79) from float and byte (unorm / snorm) images on a Rectangle, a Cube and a triangle mesh: every map of a
    vertex is looked up with its one stochastic_r in shade_a and again in shade_b."""
    w, spp = 128, 16
    n = scenes.surface_maps_scene(w, w, spp=spp, nearest=nearest)
    scene, view = su.compile_scene()
    ref = oracle.render(scene, view, w, w, 0, spp, num_meshes=n)
    su.render_frame(0)
    gpu = download_film(w, w)
    assert np.array_equal(gpu[..., 3], ref[..., 3])
    rel = rel_error(gpu, ref)
    assert np.median(rel) < 5e-6
    assert (rel > 1e-2).mean() < 5e-3
    assert abs(gpu[..., :3].mean() - ref[..., :3].mean()) / ref[..., :3].mean() < 2e-4


def download_aov_layer(aov_class, width, height):
    L = lib.load_library()
    L.zygpu_resolve_aov.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_int]
    layer = np.zeros((height, width, 4), np.float32)
    assert 0 == L.zygpu_resolve_aov(su.device_handle(), aov_class, layer.ctypes.data, width * height, 1), L.zygpu_last_error()
    return layer


@pytest.mark.parametrize("scene_name,filter_name", [("surface_maps", None), ("surface_maps", "Mitchell"), ("glass", None), ("mesh", "Blackman")])
def test_aov_layers_match_oracle(engine, scene_name, filter_name):
    """All nine AOV classes (Worker.commonAOV, worker.zig:209-242; Sensor.addSample, sensor.zig:197-377; aov.Buffer,
    aov_buffer.zig) next to the beauty: the unresolved layers against the oracle's, then su_resolve_frame_to_buffer against
    aov.Buffer.resolve. Glass: the albedo is the last primary-ray vertex's (split paths, pool order)."""
    w, spp = 96, 8
    n = 0
    if "surface_maps" == scene_name:
        n = scenes.surface_maps_scene(w, w, spp=spp, filter_name=filter_name)
    elif "glass" == scene_name:
        scenes.cornell_box(w, w, spp=spp, glass={"roughness": 0.0}, filter_name=filter_name)
    else:
        n = scenes.sphere_scene(w, w, spp=spp, quads=(96, 48), filter_name=filter_name)
    su.aovs_create({name: True for name in oracle.AOV_CLASSES})
    scene, view = su.compile_scene()
    ref_film, ref = oracle.render_aov(scene, view, w, w, 0, spp, 0x1FF, num_meshes=n or 0)
    su.render_frame(0)
    gpu_film = download_film(w, w)
    assert np.median(rel_error(gpu_film, ref_film)) < 5e-6

    for c in range(9):
        gpu = download_aov_layer(c, w, w)
        name = oracle.AOV_CLASSES[c]
        if c in (1, 2):  # Depth (min over the pixel's samples) and MaterialId (first sample of the largest centre weight): stored values
            same = gpu[..., 0] == ref[c][..., 0]
            assert same.mean() > 0.998, f"{name}: {100 * (1 - same.mean()):.3f} % of the pixels differ"
            continue
        assert np.allclose(gpu[..., 3], ref[c][..., 3], rtol=2e-6, atol=0), name
        d = np.abs(gpu[..., :3] - ref[c][..., :3]).sum(-1)
        scale = np.maximum(np.abs(ref[c][..., :3]).sum(-1), 1e-3)
        assert np.median(d / scale) < 5e-6, name
        assert (d / scale > 1e-2).mean() < 5e-3, name

    # the C API: resolved classes, -2 for a class that is not recorded
    for c in (0, 1, 3, 5, 7):
        got = su.resolve_frame_to_buffer(w, w, c)
        want = oracle.resolve_aov(c, download_aov_layer(c, w, w))
        assert np.allclose(got, want, rtol=1e-6, atol=1e-7, equal_nan=True), oracle.AOV_CLASSES[c]
    su.aovs_create({"Roughness": False})
    su.render_frame(0)
    assert -2 == su._su().su_resolve_frame_to_buffer(5, w, w, np.zeros((w, w, 4), np.float32).ctypes.data)
    assert 0 == su._su().su_resolve_frame(0) and -2 == su._su().su_resolve_frame(5) and 0 == su._su().su_resolve_frame(9)


@pytest.mark.parametrize("normal_map", [True, False])
def test_coated_substitutes_match_oracle(engine, normal_map):
    """The clear coat (substitute_coating.zig; evaluate substitute_sample.zig:138-142, coatingSample :304-336, coatingReflect /
    coatingBaseSample :412-433) over glossy, diffuse, metallic and normal-mapped anisotropic bases, on Rectangle, Cube, Sphere and a mesh:
    the coat's two extra sampler draws per vertex, its Fresnel split, the absorption along both passes through the layer."""
    w, spp = 128, 16
    n = scenes.coated_scene(w, w, spp=spp, normal_map=normal_map)
    scene, view = su.compile_scene()
    ref = oracle.render(scene, view, w, w, 0, spp, num_meshes=n)
    su.render_frame(0)
    gpu = download_film(w, w)
    assert np.array_equal(gpu[..., 3], ref[..., 3])
    rel = rel_error(gpu, ref)
    assert np.median(rel) < 5e-6
    assert (rel > 1e-2).mean() < 5e-3
    assert abs(gpu[..., :3].mean() - ref[..., :3].mean()) / ref[..., :3].mean() < 2e-4


@pytest.mark.parametrize("sigma", [1.0, 0.6])
def test_denoise_matches_oracle(engine, sigma):
    """`it --denoise` as a post kernel (src/it/denoise.zig) over the device's film and its ShadingNormal / Albedo layers, against the
    oracle's restatement fed with the very same buffers; -2 without those AOV classes."""
    w, spp = 96, 4
    n = scenes.surface_maps_scene(w, w, spp=spp)
    scene, view = su.compile_scene()
    su.render_frame(0)
    assert -2 == su._su().zyg_su_denoise_frame_to_buffer(C.c_float(sigma), w, w, np.zeros((w, w, 4), np.float32).ctypes.data)
    su.aovs_create({"Albedo": True, "ShadingNormal": True})
    scene, view = su.compile_scene()
    su.render_frame(0)
    got = su.denoise_frame_to_buffer(sigma, w, w)
    want = oracle.denoise(view, download_film(w, w), download_aov_layer(4, w, w), download_aov_layer(0, w, w), sigma)
    assert np.allclose(got, want, rtol=2e-4, atol=2e-6)
    plain = su.resolve_frame_to_buffer(w, w)
    assert np.abs(got[..., :3] - plain[..., :3]).mean() > 1e-3  # it did something


@pytest.mark.parametrize("scene_name,filter_name", [("glass", None), ("glass", "Mitchell"), ("sky", None), ("mesh", None)])
def test_transparent_film_matches_oracle(engine, scene_name, filter_name, tmp_path, monkeypatch):
    """Sensor "alpha_transparency" (buffer_transparent.zig): the alpha of Pool.transparency (vertex.zig:243-268) — see-through paths
    through glass (Transmission keeps state.transparent), a sky that covers by its brightness, opaque hits — resolved next to the colour,
    and written into the exported PNG."""
    w, spp = 96, 8
    n = 0
    if "glass" == scene_name:
        scenes.cornell_box(w, w, spp=spp, glass={"roughness": 0.0})
    elif "sky" == scene_name:
        scenes.sky_scene(w, w, spp=spp, sky_size=64, sun=None)
    else:
        n = scenes.sphere_scene(w, w, spp=spp, quads=(96, 48))
    su.sensor_create({"alpha_transparency": True, "filter": {filter_name: {}}} if filter_name else {"alpha_transparency": True})
    scene, view = su.compile_scene()
    ref_film, ref_alpha = oracle.render_alpha(scene, view, w, w, 0, spp, num_meshes=n or 0)
    want = oracle.resolve_transparent(view, ref_film, ref_alpha)
    su.render_frame(0)
    gpu_film = download_film(w, w)
    assert np.median(rel_error(gpu_film, ref_film)) < 5e-6
    got = su.resolve_frame_to_buffer(w, w)
    d = np.abs(got[..., 3] - want[..., 3])
    assert np.median(d) < 1e-6 and (d > 1e-3).mean() < 5e-3, f"alpha: median {np.median(d):.2e}, {100 * (d > 1e-3).mean():.3f} % off"
    assert 0.02 < got[..., 3].mean() <= 1.0 + 1e-4 and (("sky" != scene_name) or got[..., 3].mean() < 0.99)

    monkeypatch.chdir(tmp_path)
    su.exporters_create({"Image": {"format": "PNG"}})
    su.export_frame()
    from test_image_writer_host import read_png
    rgba = read_png("image_00_000000.png")
    assert rgba.shape == (w, w, 4)
    assert np.abs(rgba[..., 3].astype(np.float32) / 255.0 - np.clip(got[..., 3], 0.0, 1.0)).max() < 1.5 / 255.0


def test_trace_kernel_variants_are_bit_identical(tmp_path):
    """The fused traversal kernel with one ray per lane, its ray-pool variant and the ray sort in front of either walk the rays in
    different orders and hand different rays to a warp: equal-t ties are resolved by ids and the box gates use the ray's initial max_t, so
    the films are the same bytes (the tuning variables are read once per process: one process per variant)."""
    import subprocess
    import sys

    script = (
        "import sys, numpy as np, ctypes as C; sys.path.insert(0, %r)\n"
        "from zyg_b200 import lib, scenes, su\n"
        "w = 160\n"
        "scenes.instanced_scene(w, w, spp=4, grid=(24, 24), prototypes=4, quads=(48, 24), sun=60.0)\n"
        "su.render_frame(0)\n"
        "L = lib.load_library(); L.zygpu_download_film.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]\n"
        "film = np.zeros((w, w, 4), np.float32)\n"
        "assert 0 == L.zygpu_download_film(su.device_handle(), film.ctypes.data, w * w)\n"
        "np.save(sys.argv[1], film)\n" % ROOT)
    films = {}
    for name, env in (("lock_step", {"ZYGPU_SCENE_POOL": "0"}), ("pool", {"ZYGPU_SCENE_POOL": "1"}),
                      ("pool_sorted", {"ZYGPU_SCENE_POOL": "1", "ZYGPU_RAY_SORT": "3", "ZYGPU_RAY_SORT_FROM": "0"}),
                      ("two_kernels", {"ZYGPU_SCENE_TRACE": "1"})):
        out = str(tmp_path / (name + ".npy"))
        subprocess.run([sys.executable, "-c", script, out], check=True, env=dict(os.environ, **env), timeout=600)
        films[name] = np.load(out)
    assert films["lock_step"][..., :3].sum() > 0
    for name in ("pool", "pool_sorted", "two_kernels"):
        assert films[name].tobytes() == films["lock_step"].tobytes(), name


def test_4k_instanced_frame_is_reproducible(engine):
    """BASELINE config 5's scene and resolution, 4 spp, three times: the films are the same bytes. Two of the 33 M paths used to flip from
    run to run (a sliver triangle at a pole of a lat-long prototype met exactly reports a hit at t = 0; whether its leaf was reached
    depended on how far max_t had come down, i.e. on the schedule) and failed the NCCL film check of the bench at 4K."""
    w, h, spp = 3840, 2160, 4
    scenes.instanced_scene(w, h, spp=spp, grid=(100, 100), prototypes=20, quads=(500, 250), sun=60.0)
    L = lib.load_library()
    L.zygpu_render.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32]
    L.zygpu_clear_film.argtypes = [C.c_void_p]
    su.start_frame(0)
    dev = su.device_handle()
    films = []
    for _ in range(3):
        assert 0 == L.zygpu_clear_film(dev)
        assert 0 == L.zygpu_render(dev, 0, spp)
        films.append(download_film(w, h))
    assert films[0][..., 3].min() == spp
    assert films[1].tobytes() == films[0].tobytes() and films[2].tobytes() == films[0].tobytes()


def test_export_frame_writes_the_aov_layers(engine, tmp_path, monkeypatch):
    """Driver.exportFrame (driver.zig:224-253): after the beauty, one file per recorded AOV class and exporter, named and encoded by
    its class (image_sequence.zig:36-78, aov_value.zig:32-40)."""
    from test_image_writer_host import read_exr, read_png

    w, h, spp = 48, 40, 4
    scenes.cornell_box(w, h, spp=spp)
    su.aovs_create({"Albedo": True, "Depth": True, "MaterialId": True, "ShadingNormal": True})
    su.exporters_create({"Image": {"format": "EXR", "bitdepth": 32}})
    monkeypatch.chdir(tmp_path)
    su.render_frame(3)
    su.export_frame()
    import os as _os
    assert sorted(_os.listdir(".")) == ["image_00_000003.exr", "image_00_000003_albedo.exr", "image_00_000003_depth.exr",
                                        "image_00_000003_mat.exr", "image_00_000003_n.exr"]
    depth, names, _, _ = read_exr("image_00_000003_depth.exr")
    assert names == ["Y"] and np.array_equal(depth[..., 0], su.resolve_frame_to_buffer(w, h, su.AOV_DEPTH)[..., 0])
    ids, names, _, _ = read_exr("image_00_000003_mat.exr")
    assert names == ["Y"] and np.array_equal(ids[..., 0], su.resolve_frame_to_buffer(w, h, su.AOV_MATERIAL_ID)[..., 0])
    normal, names, _, _ = read_exr("image_00_000003_n.exr")
    assert names == ["B", "G", "R"] and np.array_equal(normal[..., ::-1], su.resolve_frame_to_buffer(w, h, su.AOV_SHADING_NORMAL)[..., :3], equal_nan=True)
    su.exporters_create({"Image": {"format": "PNG"}})
    su.export_frame()
    assert read_png("image_00_000003_depth.png").shape == (h, w, 1) and read_png("image_00_000003_mat.png").shape == (h, w, 3)


@pytest.mark.parametrize("filter_name", [None, "Mitchell"])
def test_disk_props_match_oracle(engine, filter_name):
    """Disk.intersect / intersectP / fragment (disk.zig:28-134) as closest-hit and shadow geometry: a glossy, a metallic and an emissive
    disk that is met by the paths (no next-event estimation towards it)."""
    w, spp = 128, 16
    scenes.disk_scene(w, w, spp=spp, filter_name=filter_name)
    scene, view = su.compile_scene()
    ref = oracle.render(scene, view, w, w, 0, spp)
    su.render_frame(0)
    gpu = download_film(w, w)
    if filter_name is None:
        assert np.array_equal(gpu[..., 3], ref[..., 3])
    rel = rel_error(gpu, ref)
    assert np.median(rel) < 5e-6
    assert (rel > 1e-2).mean() < 5e-3
    assert abs(gpu[..., :3].mean() - ref[..., :3].mean()) / ref[..., :3].mean() < 2e-4


@pytest.mark.parametrize("deferred, split_threshold, num_samples", [(False, 0.5, 1), (False, 0.0, 4), (True, 0.5, 3)])
def test_disk_lights_match_oracle(engine, monkeypatch, deferred, split_threshold, num_samples):
    """Disk.sampleTo (equi-angular sampling of the chord through the shading point's foot, disk.zig:181-332), Disk.pdf (:492-533) for
    paths that meet the lit disk and Disk.emission (:171-179) for the un-occluding lamp; in the shade kernel and in the light kernels."""
    monkeypatch.setenv("ZYGPU_DEFERRED_LIGHTS", "1" if deferred else "0")
    w, spp = 128, 16
    scenes.disk_scene(w, w, spp=spp, disk_lights=True, split_threshold=split_threshold, num_samples=num_samples)
    scene, view = su.compile_scene()
    ref = oracle.render(scene, view, w, w, 0, spp, wavefront_light_order=True)  # two lights: several picks per vertex
    su.render_frame(0)
    gpu = download_film(w, w)
    assert np.array_equal(gpu[..., 3], ref[..., 3])
    rel = rel_error(gpu, ref)
    assert np.median(rel) < 5e-6
    assert (rel > 1e-2).mean() < 5e-3
    assert abs(gpu[..., :3].mean() - ref[..., :3].mean()) / ref[..., :3].mean() < 2e-4
