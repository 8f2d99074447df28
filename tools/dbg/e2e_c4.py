import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from zyg_b200 import scenes, su
for sky in (None, 1024):
    su.release()
    scenes.mesh_lights_scene(1920, 1080, spp=4, num_lights=1000, geometry_quads=(400, 250), sun=15.0, sky=sky, max_depth=8)
    for i in range(3):
        t = time.time(); su.compile_scene(); t1 = time.time()
        su.start_frame(0); t2 = time.time()
        su.render_frame_range(0, 0, 4); t3 = time.time()
        rgba = su.resolve_frame_to_buffer(1920, 1080); t4 = time.time()
        print(f"sky={sky} compile {t1-t:.3f} start_frame {t2-t1:.3f} render_frame_range {t3-t2:.3f} resolve {t4-t3:.3f}", flush=True)
