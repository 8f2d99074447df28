"""Sample-range independence on config 5: render(4, 4) before and after other passes, and [0, 4) + [4, 8) against [0, 8)."""
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
b1 = render(4, 4)
c = render(0, 8)
b2 = render(4, 4)
a = render(0, 4)
print("render(4,4) before / after other passes identical:", b1.tobytes() == b2.tobytes(), float(np.abs(b1 - b2).max()), flush=True)
s = a + b1
err = np.abs(s - c) / np.maximum(np.abs(c), 1e-3)
print("[0,4)+[4,8) vs [0,8): max rel", float(err.max()), "pixels above 2e-6:", int((err.max(-1) > 2e-6).sum()), flush=True)
ys, xs = np.nonzero(err.max(-1) > 2e-6)
for y, x in list(zip(ys.tolist(), xs.tolist()))[:5]:
    print(y, x, a[y, x], b1[y, x], c[y, x])
