# launch list of one device build of the 1 M-triangle mesh (warm second build), plus the wall / event split
cat > /tmp/build_once.py <<'P'
import sys, time
sys.path.insert(0, ".")
from zyg_b200 import lib, scenes
positions, normals, uvs, indices = scenes.displaced_sphere(1000, 500)
dev = lib.Device(0)
for k in range(3):
    t0 = time.perf_counter()
    m = lib.Mesh(positions, indices, normals, uvs, device=dev)
    print(f"build {k}: wall {(time.perf_counter() - t0) * 1e3:.1f} ms, device events {m.build_ms:.2f} ms", flush=True)
P
python /tmp/build_once.py > gpurun_out/${TAG}_build.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches_build1m.csv python /tmp/build_once.py > /dev/null 2>&1
cat gpurun_out/${TAG}_build.log
