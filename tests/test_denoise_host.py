"""The `it` tool's denoise operator (src/it/denoise.zig) restated in the oracle. No reference vectors exist; pins:
a noise-free image passes through, blending stops at normal / albedo edges, a 4-spp frame moves towards the converged one."""

import numpy as np
import pytest

import oracle_lib as oracle
from zyg_b200 import scenes, su


@pytest.fixture()
def engine():
    su.release()
    yield
    su.release()


def _view(w):
    scenes.cornell_box(w, w, spp=1)
    _, view = su.compile_scene()
    return view


def _to_ap1(srgb):  # inverse of aces.AP1tosRGB, for comparing the operator's output with its input
    m = np.array([[1.70505155, -0.62179068, -0.08325840], [-0.13025714, 1.14080289, -0.01054853], [-0.02400328, -0.12896877, 1.15297171]])
    return srgb[..., :3] @ np.linalg.inv(m).T


def test_flat_image_passes_through_and_edges_stop_the_blend(engine):
    w = 32
    view = _view(w)
    rng = np.random.default_rng(3)
    film = np.ones((w, w, 4), np.float32)
    film[..., :3] = 0.4
    normal = np.zeros((w, w, 4), np.float32)
    normal[..., 2] = 1.0
    normal[..., 3] = 1.0
    albedo = np.full((w, w, 4), 0.5, np.float32)
    albedo[..., 3] = 1.0
    out = oracle.denoise(view, film, normal, albedo, 1.0)
    assert np.allclose(_to_ap1(out), 0.4, atol=1e-5) and np.all(out[..., 3] == 1.0)  # no noise: noise_estimate 0, every tap returns the pixel

    # noise on the left half, whose normals differ from the right half's: the right half must not receive any of it
    noisy = film.copy()
    noisy[:, : w // 2, :3] *= rng.uniform(0.2, 1.8, (w, w // 2, 1)).astype(np.float32)
    noisy[:, w // 2 :, :3] = 0.9
    normal[:, : w // 2, :3] = [1.0, 0.0, 0.0]
    out = _to_ap1(oracle.denoise(view, noisy, normal, albedo, 1.0))
    left_in, left_out = noisy[:, 2 : w // 2 - 4, 0], out[:, 2 : w // 2 - 4, 0]
    assert left_out.std() < 0.6 * left_in.std()                           # the noisy half is smoothed ...
    assert abs(left_out.mean() - left_in.mean()) < 0.05 * left_in.mean()  # ... around the same mean
    assert np.allclose(out[:, w // 2 + 1 :, :], 0.9, atol=1e-5)           # ... and does not leak across the normal edge
    # same with an albedo edge instead of a normal edge
    normal[..., :3] = [0.0, 0.0, 1.0]
    albedo[:, : w // 2, :3] = [2.0, 0.0, 0.0]  # distance >= 1: (1 - dist_albedo) = 0
    out = _to_ap1(oracle.denoise(view, noisy, normal, albedo, 1.0))
    assert np.allclose(out[:, w // 2 + 1 :, :], 0.9, atol=1e-5)


def test_low_sample_frame_moves_towards_the_converged_one(engine):
    w = 64
    scenes.cornell_box(w, w, spp=512)
    su.aovs_create({"Albedo": True, "ShadingNormal": True})
    scene, view = su.compile_scene()
    slots = (1 << 0) | (1 << 4)
    ref, _ = oracle.render_aov(scene, view, w, w, 0, 512, slots)
    film, layers = oracle.render_aov(scene, view, w, w, 512, 4, slots)  # other samples than the reference's
    truth = oracle.resolve(view, ref)[..., :3]
    noisy = oracle.resolve(view, film)[..., :3]
    clean = oracle.denoise(view, film, layers[4], layers[0], 1.0)[..., :3]
    assert np.isfinite(clean).all()
    # away from the lamp: it is an un-occluding emitter, so its pixels carry the ceiling's normal and albedo and the operator (like the
    # reference's) smears its edge
    bright = truth.max(-1) > 2.0
    near = np.zeros_like(bright)
    for dy in range(-4, 5):
        for dx in range(-4, 5):
            near |= np.roll(np.roll(bright, dy, 0), dx, 1)
    mse = lambda a: float(((a - truth)[~near] ** 2).mean())
    assert mse(clean) < 0.7 * mse(noisy)
    assert abs(clean.mean() - noisy.mean()) < 0.03 * noisy.mean()  # a blend of the neighbourhood: energy stays
