"""Debug: whole frame vs separately rendered sample ranges, and process-to-process reproducibility."""
import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from zyg_b200 import lib, scenes, su

tag = sys.argv[1]
w, h, spp = 960, 540, 8
scenes.instanced_scene(w, h, spp=spp, grid=(60, 60), prototypes=6, quads=(120, 60), sun=60.0)
L = lib.load_library()
L.zygpu_download_film.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
def film():
    f = np.zeros((h, w, 4), np.float32)
    assert 0 == L.zygpu_download_film(su.device_handle(), f.ctypes.data, w * h)
    return f
su.render_frame_range(0, 0, 8); whole = film()
su.render_frame_range(0, 0, 4); a = film()
su.render_frame_range(0, 4, 4); b = film()
su.render_frame_range(0, 0, 8); whole2 = film()
print(tag, "whole == whole2:", whole.tobytes() == whole2.tobytes())
d = np.abs((a + b) - whole) / np.maximum(np.abs(whole), 1e-3)
print(tag, "split vs whole: max rel", d.max(), "pixels > 1e-4:", int((d.max(-1) > 1e-4).sum()))
np.save(f"gpurun_out/split_{tag}.npy", whole)
