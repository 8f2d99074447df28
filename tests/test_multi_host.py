"""Multi-GPU schedule on the host side: sample-range partition and the film reduce, exercised with
world_size-2 gloo process groups on the CPU. The films come from the oracle here (no GPU in this suite); the same
zyg_b200.multi functions drive the device path in bench.py --gpus N and tests/test_multi_gpu.py."""

import os
import socket
import sys

import numpy as np
import pytest

from zyg_b200 import multi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("spp,world", [(64, 1), (64, 2), (64, 8), (7, 4), (3, 8), (4096, 8), (0, 2)])
def test_sample_ranges_tile_the_frame(spp, world):
    ranges = [multi.sample_range(r, world, spp) for r in range(world)]
    cursor = 0
    for first, count in ranges:
        assert first == cursor and count >= 0
        cursor += count
    assert cursor == spp
    counts = [c for _, c in ranges]
    assert max(counts) - min(counts) <= 1


def test_sample_range_rejects_bad_rank():
    with pytest.raises(ValueError):
        multi.sample_range(2, 2, 16)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, spp, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist

    import oracle_lib as oracle
    from zyg_b200 import scenes, su

    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        w = 24
        scenes.cornell_box(w, w, spp=spp, filter_name="Mitchell")
        scene, view = su.compile_scene()
        first, count = multi.sample_range(rank, world, spp)
        film = oracle.render(scene, view, w, w, first, count, threads=2)
        total = multi.reduce_host_films(film, rank, world)
        if 0 == rank:
            whole = oracle.render(scene, view, w, w, 0, spp, threads=2)
            np.save(os.path.join(out_dir, "reduced.npy"), total)
            np.save(os.path.join(out_dir, "whole.npy"), whole)
        su.release()
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_reduce_equals_single_process(tmp_path):
    import torch.multiprocessing as mp

    spp, world = 6, 2
    mp.spawn(_worker, args=(world, _free_port(), spp, str(tmp_path)), nprocs=world, join=True)
    reduced = np.load(tmp_path / "reduced.npy")
    whole = np.load(tmp_path / "whole.npy")
    # the weight channel is reduced too (non-uniform with a radius-2 filter): exact up to fp32 summation order
    assert np.allclose(reduced, whole, rtol=2e-6, atol=1e-6)
    assert reduced[..., 3].min() > 0
