"""Debug (torchrun, 2 ranks): is rank 1's half of the samples bit-identical to rank 0 rendering the same half, and does the
NCCL-reduced film equal the sum?"""
import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
from zyg_b200 import lib, multi, scenes, su

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
big = len(sys.argv) > 1 and sys.argv[1] == "big"
w, h, spp = (3840, 2160, 8) if big else (960, 540, 8)
kw = dict(grid=(100, 100), prototypes=20, quads=(500, 250), sun=60.0) if big else dict(grid=(60, 60), prototypes=6, quads=(120, 60), sun=60.0)
scenes.instanced_scene(w, h, spp=spp, **kw)
su._ok(su._su().zyg_su_set_device(local), "set_device")
L = lib.load_library()
L.zygpu_download_film.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
L.zygpu_clear_film.argtypes = [C.c_void_p]
L.zygpu_render.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32]
def film():
    f = np.zeros((h, w, 4), np.float32)
    assert 0 == L.zygpu_download_film(su.device_handle(), f.ctypes.data, w * h)
    return f
first, count = multi.sample_range(rank, world, spp)
su.render_frame_range(0, first, count)
mine = film()
t = torch.from_numpy(mine).cuda()
halves = [torch.empty_like(t) for _ in range(world)]
dist.all_gather(halves, t)
torch.cuda.synchronize()
if rank == 0:
    su.render_frame_range(0, 4, 4)
    other_local = film()
    other_remote = halves[1].cpu().numpy()
    diff = (other_local != other_remote).any(-1)
    print("rank 1's half == rank 0 rendering the same half:", other_local.tobytes() == other_remote.tobytes(), "pixels differing:", int(diff.sum()), flush=True)
    if diff.any():
        ys, xs = np.nonzero(diff)
        print("  first differing pixels:", list(zip(ys[:5].tolist(), xs[:5].tolist())), other_local[ys[0], xs[0]], other_remote[ys[0], xs[0]])
# the reduce path
dev = su.device_handle()
su.start_frame(0)
for _ in range(3):
    su._ok(L.zygpu_clear_film(dev), "clear")
    su._ok(L.zygpu_render(dev, first, count), "render")
    multi.reduce_film(0)
multi.synchronize()
dist.barrier()
if rank == 0:
    reduced = film()
    total = mine + halves[1].cpu().numpy()
    d = np.abs(reduced - total) / np.maximum(np.abs(total), 1e-3)
    print("reduced vs sum of halves: max rel", float(d.max()), "pixels > 1e-5:", int((d.max(-1) > 1e-5).sum()), flush=True)
dist.barrier()
su.release()
dist.destroy_process_group()
