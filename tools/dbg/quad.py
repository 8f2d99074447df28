import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, 'tests')
import numpy as np
import oracle_lib as oracle
from zyg_b200 import lib
rng = np.random.default_rng(3)
positions = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0]], np.float32)
indices = np.array([[0, 1, 2], [0, 2, 3]], np.uint32)
mesh = lib.Mesh(positions, indices)
n = 1 << 13
rays = np.empty(n, lib.RAY_DTYPE)
rays["origin"] = (rng.random((n, 3)) * 2 - 1).astype(np.float32) * np.float32([1, 1, 0]) + np.float32([0, 0, 3])
target = (rng.random((n, 3)) * 2 - 1).astype(np.float32) * np.float32([1, 1, 0])
target[: n // 4] = np.round(target[: n // 4] * 16) / 16
d = target - rays["origin"]
rays["direction"] = d / np.linalg.norm(d, axis=1, keepdims=True)
rays["min_t"], rays["max_t"] = 0, lib.RAY_MAX_T
nodes, tris, pos = (mesh.data(w) for w in (lib.MESH_BINARY_NODES, lib.MESH_TRIANGLES, lib.MESH_POSITIONS))
print(nodes, tris, pos)
dev = lib.Device(0)
m0 = lib.Mesh(np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float32)); dev.upload_mesh(m0)
mid = dev.upload_mesh(mesh)
ref = oracle.trace_closest(nodes, tris, pos, rays)
exact = dev.trace_batch(mid, lib.CLOSEST_BINARY, rays)
bad = np.nonzero((exact.view(np.uint32).reshape(-1,4) != ref.view(np.uint32).reshape(-1,4)).any(1))[0]
print(len(bad), bad[:20])
for i in bad[:10]:
    print(i, rays[i], 'ref', ref[i], ref[i:i+1].view(np.uint32), 'dev', exact[i], exact[i:i+1].view(np.uint32))
