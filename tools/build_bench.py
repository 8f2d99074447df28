"""Device vs host BVH build: build time, tree size and the traversal rate each tree gives on the same rays.
usage: tools/build_bench.py [nu nv]   (default 1000 500 = 1M triangles)"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402  (device buffers and events only)

from zyg_b200 import lib, scenes  # noqa: E402

nu, nv = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1000, 500)
positions, normals, uvs, indices = scenes.displaced_sphere(nu, nv)
dev = lib.Device(0)

t0 = time.perf_counter()
host = lib.Mesh(positions, indices, normals, uvs)
host_s = time.perf_counter() - t0
lib.Mesh(positions, indices, normals, uvs, device=dev)  # warm-up (allocator, module load)
t0 = time.perf_counter()
built = lib.Mesh(positions, indices, normals, uvs, device=dev)
call_s = time.perf_counter() - t0

rays = {"primary": scenes.primary_rays(2048, 2048), "incoherent": scenes.random_rays(1 << 22),
        "shadow": scenes.random_rays(1 << 22, shadow=True)}
for name, mesh, secs in (("host SAH + spatial splits", host, host_s), ("device LBVH", built, call_s)):
    info = mesh.info()
    line = (f"{name}: {info.num_source_triangles} triangles -> {info.num_tree_triangles} references, {info.num_wide_nodes} wide nodes, "
            f"depth {info.wide_max_depth}; build {secs * 1e3:.1f} ms wall" + (f" ({mesh.build_ms:.2f} ms on the device)" if mesh.build_ms else ""))
    mid = dev.upload_mesh(mesh)
    for kind, r in rays.items():
        mode = lib.ANY if kind == "shadow" else lib.CLOSEST
        d_r = torch.from_numpy(r.view(np.float32).reshape(-1, 8)).cuda()
        d_o = torch.empty((r.shape[0], 1 if mode == lib.ANY else 4), dtype=torch.int32, device="cuda")
        best = 1e9
        for _ in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            dev.trace_batch_ptr(mid, mode, d_r.data_ptr(), r.shape[0], d_o.data_ptr(), host=False, stream=torch.cuda.current_stream().cuda_stream)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        line += f" | {kind} {r.shape[0] / best / 1e3:.0f} Mrays/s"
    print(line, flush=True)
ms = dev.refit_mesh(built, positions * np.float32(1.01))
print(f"refit of the device-built mesh: {ms:.2f} ms on the device")
ms = dev.refit_mesh(host, positions * np.float32(1.01))
print(f"refit of the host-built mesh: {ms:.2f} ms on the device")
