import sys, os, time
sys.path.insert(0, os.getcwd()); sys.path.insert(0, 'tests'); sys.path.insert(0, 'tools')
import numpy as np
import oracle_lib as oracle
from zyg_b200 import scenes, su
from png import write_png
w = int(sys.argv[1]); spp = int(sys.argv[2]); rough = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0
su.release()
scenes.cornell_box(w, w, spp=spp, glass={"roughness": rough})
scene, view = su.compile_scene()
t = time.time()
film = oracle.render(scene, view, w, w, 0, spp)
print('oracle', time.time() - t, 's; mean', film[..., :3].mean(), 'max', film[..., :3].max(), 'nan', np.isnan(film).sum())
rgba = oracle.resolve(view, film)
os.makedirs('gpurun_out', exist_ok=True)
write_png(f'gpurun_out/glass_ref_{rough}.png', rgba)
su.release()
