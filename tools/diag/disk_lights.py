"""Disk-light parity statistics: device film vs oracle film of scenes.disk_scene(disk_lights=True). Run on a GPU box from the repo root."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "..", "tests"))
import oracle_lib as oracle  # noqa: E402
import ctypes as C  # noqa: E402
from zyg_b200 import lib, scenes, su  # noqa: E402

for deferred, st, ns, depth in [(0, 0.5, 1, 6), (0, 0.5, 1, 1), (0, 0.0, 4, 1), (1, 0.5, 3, 6)]:
    os.environ["ZYGPU_DEFERRED_LIGHTS"] = str(deferred)
    su.release()
    w, spp = 128, 16
    scenes.disk_scene(w, w, spp=spp, disk_lights=True, split_threshold=st, num_samples=ns, max_depth=depth)
    scene, view = su.compile_scene()
    ref = oracle.render(scene, view, w, w, 0, spp, wavefront_light_order=True)
    su.render_frame(0)
    L = lib.load_library()
    L.zygpu_download_film.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
    gpu = np.zeros((w, w, 4), np.float32)
    assert 0 == L.zygpu_download_film(su.device_handle(), gpu.ctypes.data, w * w)
    rel = np.abs(gpu[..., :3] - ref[..., :3]).max(-1) / np.maximum(np.abs(ref[..., :3]).max(-1), 1e-6)
    print(deferred, st, ns, depth, "weights equal", np.array_equal(gpu[..., 3], ref[..., 3]), "median", np.median(rel), "p90", np.percentile(rel, 90),
          "p99", np.percentile(rel, 99), "max", rel.max(), "mean ratio", gpu[..., :3].mean() / ref[..., :3].mean(), flush=True)
