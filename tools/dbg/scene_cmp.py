"""Device vs oracle on a named scene of zyg_b200.scenes: SCENE=name KW='{"json":"kwargs"}' scene_cmp.py w spp [cpu]"""
import sys, os, time, json, ctypes as C
sys.path.insert(0, os.getcwd()); sys.path.insert(0, 'tests'); sys.path.insert(0, 'tools')
import numpy as np
import oracle_lib as oracle
from zyg_b200 import scenes, su, lib
from png import write_png
w = int(sys.argv[1]); spp = int(sys.argv[2]); cpu_only = len(sys.argv) > 3
name = os.environ.get("SCENE", "cornell_box"); kw = json.loads(os.environ.get("KW", "{}"))
su.release()
r = getattr(scenes, name)(w, w, spp=spp, **kw)
num_meshes = r if isinstance(r, int) and name in ("sphere_scene", "instanced_scene", "mesh_lights_scene") else 0
scene, view = su.compile_scene()
t = time.time(); ref = oracle.render(scene, view, w, w, 0, spp, num_meshes=num_meshes, wavefront_light_order=bool(int(os.environ.get('WF', '1')))); print('oracle', round(time.time() - t, 2), 's mean', ref[..., :3].mean(), 'nan', np.isnan(ref).sum())
os.makedirs('gpurun_out', exist_ok=True)
write_png(f'gpurun_out/{name}_ref.png', oracle.resolve(view, ref))
if not cpu_only:
    t = time.time(); su.render_frame(0); print('gpu frame', round(time.time() - t, 3))
    L = lib.load_library()
    L.zygpu_download_film.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
    gpu = np.zeros((w, w, 4), np.float32)
    assert 0 == L.zygpu_download_film(su.device_handle(), gpu.ctypes.data, w * w), L.zygpu_last_error()
    d = np.abs(gpu[..., :3] - ref[..., :3]).sum(-1); rel = d / np.maximum(np.abs(ref[..., :3]).sum(-1), 1e-6)
    print(name, kw, 'weights equal', np.array_equal(gpu[..., 3], ref[..., 3]), 'median rel', np.median(rel), 'frac>1e-3', (rel > 1e-3).mean(),
          'mean gpu/ref', gpu[..., :3].mean(), ref[..., :3].mean(), 'nan', np.isnan(gpu).sum())
    write_png(f'gpurun_out/{name}_gpu.png', su.resolve_frame_to_buffer(w, w))
su.release()
