import sys, time, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from zyg_b200 import scenes, su, lib
L = lib.load_library()
L.zygpu_clear_film.argtypes = [C.c_void_p]
L.zygpu_render.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32]
L.zygpu_synchronize.argtypes = [C.c_void_p]
su.release()
scenes.mesh_lights_scene(1920, 1080, spp=4, num_lights=1000, geometry_quads=(400, 250), sun=15.0, sky=1024, max_depth=8)
su.start_frame(0)
dev = su.device_handle()
for i in range(10):
    t0 = time.time(); su.compile_scene(); t1 = time.time(); su.start_frame(0); t2 = time.time()
    L.zygpu_render(dev, 0, 4); t3 = time.time(); L.zygpu_synchronize(dev); t4 = time.time()
    print(f"compile {t1-t0:.3f} start_frame(compile+upload) {t2-t1:.3f} render(enqueue) {t3-t2:.3f} sync {t4-t3:.3f}", flush=True)
