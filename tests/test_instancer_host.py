"""The scene-file "Instancer" entity (src/util/scene_loader.zig:401-508, src/core/scene/prop/instancer.zig) behind
zyg_su_instancer_create: prototypes leave the scene's prop tree, instances are flattened into props with the composed
transformation (ComposedTransformation.transform, composed_transformation.zig:55-68). Checked with the oracle (no GPU)."""

import numpy as np
import pytest

import oracle_lib as oracle
from zyg_b200 import scenes, su

KW = dict(grid=(10, 10), prototypes=3, quads=(20, 10))


@pytest.fixture()
def engine():
    su.release()
    yield
    su.release()


def render(w, spp, **kw):
    su.release()
    n = scenes.instanced_scene(w, w, spp=spp, **KW, **kw)
    scene, view = su.compile_scene()
    film = oracle.render(scene, view, w, w, 0, spp, num_meshes=n)
    su.release()
    return film


def test_identity_instancer_equals_prop_instances(engine):
    """With an identity instancer transformation the composition is exact: the film is bit-identical to the same scene
    built from su_prop_create_instance props."""
    a = render(64, 4)
    b = render(64, 4, instancer=su.transformation())
    assert np.array_equal(a, b)


def test_instancer_transformation_composes(engine):
    """Instances under a rotated, scaled, translated instancer == props placed with M_instance * M_instancer (row
    vectors), up to the rounding of composing in fp32."""
    outer = su.transformation((0.4, 0.3, -0.2), (1.25, 1.25, 1.25), (0.0, 35.0, 0.0))
    a = render(64, 8, instancer=outer)

    # the same placement through su_prop_create_instance: patch the scene builder's matrices
    original = su.prop_set_transformation
    instances = []

    def capture(prop, matrix):
        instances.append(prop)
        original(prop, matrix)

    su.release()
    su.prop_set_transformation = capture
    try:
        n = scenes.instanced_scene(64, 64, spp=8, **KW)
    finally:
        su.prop_set_transformation = original
    # entities 1 .. grid*grid+... : only the su_prop_create_instance props (they come right after the camera and prototypes)
    first = 1 + KW["prototypes"]
    count = KW["grid"][0] * KW["grid"][1]
    rng_scene = scenes.PCG32(0, np.array([3], np.uint64))
    k = 0
    for gy in range(KW["grid"][1]):
        for gx in range(KW["grid"][0]):
            r = [float(rng_scene.float()[0]) for _ in range(5)]
            scale = 0.3 + 0.3 * r[1]
            x = (gx + 0.5 + 0.6 * (r[2] - 0.5)) - 0.5 * KW["grid"][0]
            z = (gy + 0.5 + 0.6 * (r[3] - 0.5)) - 0.5 * KW["grid"][1]
            m = su.transformation((x, 1.05 * scale, z), (scale, scale, scale), (0.0, 360.0 * r[4], 0.0)).astype(np.float64)
            su.prop_set_transformation(first + k, (m @ outer.astype(np.float64)).astype(np.float32))
            k += 1
    assert k == count
    scene, view = su.compile_scene()
    b = oracle.render(scene, view, 64, 64, 0, 8, num_meshes=n)
    rel = np.abs(a[..., :3] - b[..., :3]).sum(-1) / np.maximum(np.abs(b[..., :3]).sum(-1), 1e-6)
    assert np.median(rel) < 1e-4 and (rel > 1e-2).mean() < 0.02
    assert abs(a[..., :3].mean() - b[..., :3].mean()) / b[..., :3].mean() < 2e-3


def test_prototypes_leave_the_scene(engine):
    su.init()
    su.perspective_camera_create(16, 16)
    su.integrators_create({"surface": {"PTMIS": {}}})
    m = su.material_create({"rendering": {"Substitute": {"color": [0.5, 0.5, 0.5]}}})
    proto = su.prop_create(su.SPHERE, [m])
    assert su._su().zyg_su_instancer_create(0, None, 0, None, None) == -1
    bad = np.array([99], np.uint32)
    assert su._su().zyg_su_instancer_create(1, bad.ctypes.data, 0, None, None) == -1
    inst = su.instancer_create([proto], [0, 0, 5], np.stack([su.transformation((i, 0, 0)) for i in range(3)]))  # index 5 -> prototype 0
    assert inst > proto
    scene, _ = su.compile_scene()
    import ctypes as C

    num_props = C.cast(scene, C.POINTER(C.c_uint32))[0]
    assert num_props == 1 + 1 + 1 + 3  # camera, prototype, instancer entity, three instances
