"""Run-to-run determinism of config 5: render(0, 4) several times, compare the films byte for byte."""
import sys, os, numpy as np, ctypes as C
sys.path.insert(0, os.getcwd())
from zyg_b200 import lib, scenes, su
w, h = 3840, 2160
scenes.instanced_scene(w, h, spp=8, grid=(100, 100), prototypes=20, quads=(500, 250), sun=60.0)
L = lib.load_library(); L.zygpu_download_film.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
L.zygpu_render.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32]; L.zygpu_clear_film.argtypes = [C.c_void_p]
su.start_frame(0)
dev = su.device_handle()
def render(first, count):
    L.zygpu_clear_film(dev)
    assert 0 == L.zygpu_render(dev, first, count)
    film = np.zeros((h, w, 4), np.float32)
    assert 0 == L.zygpu_download_film(dev, film.ctypes.data, w * h)
    return film
ref = render(0, 4)
for k in range(int(sys.argv[1]) if len(sys.argv) > 1 else 5):
    f = render(0, 4)
    d = np.abs(f - ref).max(-1)
    ys, xs = np.nonzero(d > 0)
    print("repeat", k, "differing pixels", len(ys), list(zip(ys.tolist(), xs.tolist()))[:4], flush=True)
