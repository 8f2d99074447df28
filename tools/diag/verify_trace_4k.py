"""ZYGPU_VERIFY_TRACE=1: config 5, samples [0, 4): every closest-hit stage is re-run with one thread per ray and compared."""
import sys, os, numpy as np, ctypes as C
sys.path.insert(0, os.getcwd())
from zyg_b200 import lib, scenes, su
w, h = 3840, 2160
scenes.instanced_scene(w, h, spp=8, grid=(100, 100), prototypes=20, quads=(500, 250), sun=60.0)
L = lib.load_library()
L.zygpu_render.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32]; L.zygpu_clear_film.argtypes = [C.c_void_p]; L.zygpu_synchronize.argtypes = [C.c_void_p]
su.start_frame(0)
dev = su.device_handle()
if os.environ.get("ZYGPU_DEBUG_TRACE_ITEM"):
    L.zygpu_set_counting.argtypes = [C.c_void_p, C.c_int]
    L.zygpu_set_counting(dev, 1)  # the instrumented kernel is the one that prints
assert 0 == L.zygpu_render(dev, 0, 4)
assert 0 == L.zygpu_synchronize(dev)
print("done", flush=True)
