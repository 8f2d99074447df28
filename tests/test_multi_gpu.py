"""Sample-range split over two B200s with an NCCL film reduce. Skipped on a single-GPU box."""

import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, spp, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    from zyg_b200 import multi, scenes, su

    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        w = 96
        scenes.cornell_box(w, w, spp=spp, filter_name="Mitchell")
        su._ok(su._su().zyg_su_set_device(rank), "zyg_su_set_device")
        film = multi.render_frame_distributed(w, w, spp, rank, world)
        torch.cuda.synchronize()
        if 0 == rank:
            np.save(os.path.join(out_dir, "reduced.npy"), film.cpu().numpy())
            su.render_frame(0)
            torch.cuda.synchronize()
            np.save(os.path.join(out_dir, "whole.npy"), multi.device_film_tensor(w, w).cpu().numpy())
        dist.barrier()
        su.release()
    finally:
        dist.destroy_process_group()


def test_two_gpu_split_equals_single_gpu(tmp_path):
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, 16, str(tmp_path)), nprocs=2, join=True)
    reduced = np.load(tmp_path / "reduced.npy")
    whole = np.load(tmp_path / "whole.npy")
    assert np.allclose(reduced, whole, rtol=2e-6, atol=1e-6)
