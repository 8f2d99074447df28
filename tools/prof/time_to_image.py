"""Time to image of BASELINE config 5 (the instanced 5M-triangle scene at 3840x2160, 4096 spp) on N GPUs: every rank renders its
sample range of the resident scene, the films are reduced to rank 0 (zygpu_reduce_film), rank 0 resolves and copies the RGBA frame to
the host. Launch under torch.distributed.run; usage: time_to_image.py [spp] (default 4096). Prints one JSON line on rank 0."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from zyg_b200 import multi, scenes, su  # noqa: E402

spp = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
w, h = 3840, 2160
su.release()
t0 = time.perf_counter()
scenes.instanced_scene(w, h, spp=spp, grid=(100, 100), prototypes=20, quads=(500, 250), sun=60.0)
su._ok(su._su().zyg_su_set_device(local), "zyg_su_set_device")
su.start_frame(0)  # compile + upload
build_s = time.perf_counter() - t0
multi.render_frame_distributed(w, h, min(spp, 8 * world), rank, world)  # warm-up: path buffers, NCCL communicator
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
t0 = time.perf_counter()
multi.render_frame_distributed(w, h, spp, rank, world)
rgba = su.resolve_frame_to_buffer(w, h) if 0 == rank else None
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
seconds = time.perf_counter() - t0
if 0 == rank:
    print(json.dumps({"workload": "config5: instanced 5M-triangle scene, 3840x2160", "spp": spp, "n_gpus": world,
                      "time_to_image_s": seconds, "path_samples_per_s": w * h * spp / seconds,
                      "scene_build_compile_upload_s": build_s, "mean_rgb": float(np.asarray(rgba)[..., :3].mean()),
                      "includes": "per-rank render of spp / N samples, ncclReduce of the 132.7 MB film, resolve + D2H of the RGBA frame on rank 0"}))
if world > 1:
    dist.destroy_process_group()
