import sys, os, numpy as np, ctypes as C
sys.path.insert(0, os.getcwd())
from zyg_b200 import lib, scenes, su
w, h, spp = 3840, 2160, int(sys.argv[2])
scenes.instanced_scene(w, h, spp=spp, grid=(100, 100), prototypes=20, quads=(500, 250), sun=60.0)
L = lib.load_library(); L.zygpu_download_film.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
L.zygpu_render.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32]; L.zygpu_clear_film.argtypes = [C.c_void_p]
su.render_frame_range(0, 0, 1)
dev = su.device_handle()
films = []
for rep in range(2):
    L.zygpu_clear_film(dev)
    assert 0 == L.zygpu_render(dev, 0, spp)
    film = np.zeros((h, w, 4), np.float32)
    assert 0 == L.zygpu_download_film(dev, film.ctypes.data, w * h)
    films.append(film)
print(sys.argv[1], "run-to-run identical:", films[0].tobytes() == films[1].tobytes(), "max abs diff", float(np.abs(films[0] - films[1]).max()), flush=True)
np.save(sys.argv[1], films[0])
