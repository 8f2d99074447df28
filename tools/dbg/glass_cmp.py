import sys, os, time, ctypes as C
sys.path.insert(0, os.getcwd()); sys.path.insert(0, 'tests'); sys.path.insert(0, 'tools')
import numpy as np
import oracle_lib as oracle
from zyg_b200 import scenes, su, lib
from png import write_png
w = int(sys.argv[1]); spp = int(sys.argv[2]); rough = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0
su.release()
g = {"roughness": rough}
for k in os.environ.get("GLASS", "").split(","):
    if k: g[k] = True
scenes.cornell_box(w, w, spp=spp, glass=g, max_depth=int(os.environ.get("DEPTH", "8")))
scene, view = su.compile_scene()
ref = oracle.render(scene, view, w, w, 0, spp)
t = time.time(); su.render_frame(0); print('gpu frame', time.time() - t)
L = lib.load_library()
L.zygpu_download_film.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
gpu = np.zeros((w, w, 4), np.float32)
assert 0 == L.zygpu_download_film(su.device_handle(), gpu.ctypes.data, w * w), L.zygpu_last_error()
d = np.abs(gpu[..., :3] - ref[..., :3]).sum(-1); rel = d / np.maximum(np.abs(ref[..., :3]).sum(-1), 1e-6)
print('rough', rough, 'weights equal', np.array_equal(gpu[..., 3], ref[..., 3]), 'median rel', np.median(rel), 'frac>1e-3', (rel > 1e-3).mean(),
      'mean gpu/ref', gpu[..., :3].mean(), ref[..., :3].mean(), 'nan', np.isnan(gpu).sum())
write_png(f'gpurun_out/glass_gpu_{rough}.png', su.resolve_frame_to_buffer(w, w))
bad = np.argwhere(rel > 1e-3)
print('bad pixels sample', bad[:10].tolist())
su.release()
