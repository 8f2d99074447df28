"""Finds (pixel, iteration) pairs where device and oracle diverge, rendering one sample at a time."""
import sys, os, ctypes as C
sys.path.insert(0, os.getcwd()); sys.path.insert(0, 'tests'); sys.path.insert(0, 'tools')
import numpy as np
import oracle_lib as oracle
from zyg_b200 import scenes, su, lib
w = int(sys.argv[1]); spp = int(sys.argv[2])
g = {"roughness": float(os.environ.get("ROUGH", "0"))}
for k in os.environ.get("GLASS", "").split(","):
    if k: g[k] = True
su.release()
scenes.cornell_box(w, w, spp=spp, glass=g, max_depth=int(os.environ.get("DEPTH", "8")))
scene, view = su.compile_scene()
L = lib.load_library()
L.zygpu_download_film.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
found = []
for it in range(spp):
    ref = oracle.render(scene, view, w, w, it, 1)
    su.render_frame_range(0, it, 1)
    gpu = np.zeros((w, w, 4), np.float32)
    assert 0 == L.zygpu_download_film(su.device_handle(), gpu.ctypes.data, w * w)
    d = np.abs(gpu[..., :3] - ref[..., :3]).sum(-1); rel = d / np.maximum(np.abs(ref[..., :3]).sum(-1), 1e-6)
    bad = np.argwhere(rel > 1e-3)
    for y, x in bad[:3]:
        found.append((int(x), int(y), it, float(rel[y, x])))
    print('iteration', it, 'bad', len(bad), flush=True)
    if len(found) >= 3: break
print(found)
if found:
    x, y, it, _ = found[0]
    os.environ["ZO_DEBUG_PIXEL"] = f"{x},{y}"
    sys.stdout.flush()
    oracle.render(scene, view, w, w, it, 1, threads=1)
    os.environ["ZYGPU_DEBUG_SLOT"] = str(y * w + x)
    su.render_frame_range(0, it, 1)
    su._ok(L.zygpu_synchronize(su.device_handle()), "sync")
su.release()
