"""Renders a scene of zyg_b200.scenes on the device a few times and prints the frame time (diagnostics / profiling driver).
usage: SCENE=name KW='{"json": "kwargs"}' tools/render_scene.py width height spp [reps]"""
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from zyg_b200 import lib, scenes, su  # noqa: E402

w, h, spp = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
name, kw = os.environ.get("SCENE", "cornell_box"), json.loads(os.environ.get("KW", "{}"))
su.release()
getattr(scenes, name)(w, h, spp=spp, **kw)
L = lib.load_library()
L.zygpu_render.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32]
L.zygpu_synchronize.argtypes = [C.c_void_p]
L.zygpu_clear_film.argtypes = [C.c_void_p]
su.render_frame_range(0, 0, 1)
dev = su.device_handle()
best = 1e9
for _ in range(reps):
    L.zygpu_clear_film(dev)
    L.zygpu_synchronize(dev)
    t = time.perf_counter()
    assert 0 == L.zygpu_render(dev, 0, spp)
    assert 0 == L.zygpu_synchronize(dev), L.zygpu_last_error()
    best = min(best, time.perf_counter() - t)
print(f"{name} {kw} {w}x{h} x {spp} spp: {best * 1e3:.1f} ms, {w * h * spp / best / 1e6:.1f} Msamples/s", flush=True)
su.release()
