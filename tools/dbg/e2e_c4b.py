import sys, time, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch
torch.cuda.set_device(0)
from zyg_b200 import scenes, su, lib, multi
L = lib.load_library()
L.zygpu_clear_film.argtypes = [C.c_void_p]
L.zygpu_render.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32]
L.zygpu_synchronize.argtypes = [C.c_void_p]
su.release()
scenes.mesh_lights_scene(1920, 1080, spp=4, num_lights=1000, geometry_quads=(400, 250), sun=15.0, sky=1024, max_depth=8)
su._ok(su._su().zyg_su_set_device(0), "set")
su.start_frame(0)
dev = su.device_handle()
film = multi.device_film_tensor(1920, 1080)
for i in range(3):
    t = time.time(); L.zygpu_clear_film(dev); L.zygpu_render(dev, 0, 4); L.zygpu_synchronize(dev); print("step", time.time() - t, flush=True)
for i in range(3):
    t = time.time(); su.render_frame_range(0, 0, 4); t1 = time.time(); rgba = su.resolve_frame_to_buffer(1920, 1080); print("e2e", t1 - t, time.time() - t1, flush=True)
