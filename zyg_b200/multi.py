"""Multi-GPU schedule of the forward pass: sample-range split + one film reduce.

zyg already exposes the split: ``Driver.render(camera, frame, iteration, num_samples)`` renders samples
``[iteration, iteration + num_samples)`` of every pixel with the Sobol index derived from the absolute sample number
(src/core/rendering/driver.zig:115,141-142, worker.zig:145-149; CLI ``--sample / --num-samples``, options.zig:88-91),
and the film is a linear accumulator of (sum w*rgb, sum w) (buffer_opaque.zig:39-45). One process per GPU renders its
range with the whole scene replicated; a single reduce(sum, fp32) of the W*H*4 film to rank 0 finishes the frame.
The reduce is the C ABI's ``zygpu_reduce_film(dev, ncclComm_t, root)`` (include/zygpu.h), which a Zig host calls the same way;
torch.distributed only supplies the communicator here (and the gloo plumbing of the CPU tests).
"""

from __future__ import annotations

import ctypes as C

import numpy as np

from . import lib as _lib
from . import su


def sample_range(rank: int, world: int, spp: int) -> tuple[int, int]:
    """(iteration, num_samples) of `rank`: contiguous ranges that tile [0, spp) exactly, sizes differing by at most 1."""
    if world <= 0 or not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside world of size {world}")
    base, extra = divmod(spp, world)
    first = rank * base + min(rank, extra)
    return first, base + (1 if rank < extra else 0)


class _DeviceFilm:
    """__cuda_array_interface__ view of the device film (zygpu_film_device) so torch can wrap it without a copy."""

    def __init__(self, ptr: int, height: int, width: int):
        self.__cuda_array_interface__ = {"shape": (height, width, 4), "typestr": "<f4", "data": (ptr, False), "version": 2}


def device_film_tensor(width: int, height: int):
    """torch view (H, W, 4) of the film on the engine's device. The engine owns the memory."""
    import torch

    L = _lib.load_library()
    L.zygpu_film_device.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
    L.zygpu_film_device.restype = C.c_void_p
    n = C.c_uint64()
    ptr = L.zygpu_film_device(su.device_handle(), C.byref(n))
    if not ptr or n.value != width * height * 4:
        raise RuntimeError("no device film of that size: call su.start_frame first")
    return torch.as_tensor(_DeviceFilm(ptr, height, width), device=torch.device("cuda", device_ordinal()))


def device_ordinal() -> int:
    """CUDA ordinal of the engine's device (zyg_su_set_device), independent of torch's current device."""
    L = _lib.load_library()
    L.zygpu_device_ordinal.argtypes = [C.c_void_p]
    return int(L.zygpu_device_ordinal(su.device_handle()))


def nccl_comm(group=None) -> int:
    """The ncclComm_t of torch's NCCL process group for the engine's device (created by the first collective)."""
    import torch
    import torch.distributed as dist

    device = torch.device("cuda", device_ordinal())
    backend = (group or dist.distributed_c10d._get_default_group())._get_backend(device)
    try:
        ptr = backend._comm_ptr()
    except Exception:
        ptr = 0
    if not ptr:  # communicators are made lazily: run one collective on this device first
        dist.all_reduce(torch.zeros(1, device=device), group=group)
        torch.cuda.synchronize(device)
        ptr = backend._comm_ptr()
    return int(ptr)


def reduce_film(root: int = 0, group=None):
    """zygpu_reduce_film: one ncclReduce(sum, fp32) of the device film to `root`, enqueued on the render stream behind the passes."""
    L = _lib.load_library()
    L.zygpu_reduce_film.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    su._ok(L.zygpu_reduce_film(su.device_handle(), nccl_comm(group), root), "zygpu_reduce_film")


def synchronize():
    """Waits for the render stream (passes and a reduce enqueued behind them)."""
    L = _lib.load_library()
    L.zygpu_synchronize.argtypes = [C.c_void_p]
    su._ok(L.zygpu_synchronize(su.device_handle()), "zygpu_synchronize")


def render_stream():
    """The engine's render stream as a torch ExternalStream (events, ordering of the reduce after the pass)."""
    import torch

    L = _lib.load_library()
    L.zygpu_render_stream.argtypes = [C.c_void_p]
    L.zygpu_render_stream.restype = C.c_void_p
    return torch.cuda.ExternalStream(L.zygpu_render_stream(su.device_handle()))


def render_frame_distributed(width: int, height: int, spp: int, rank: int, world: int, frame: int = 0, reduce: bool = True):
    """su_start_frame + this rank's sample range + reduce to rank 0. Returns the film tensor (complete on rank 0); the render
    stream has been waited for, so the tensor can be read from any stream.

    The caller has built the same scene on every rank and initialised torch.distributed with the NCCL backend."""
    su.start_frame(frame)
    first, count = sample_range(rank, world, spp)
    L = _lib.load_library()
    L.zygpu_render.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32]
    if count > 0:
        su._ok(L.zygpu_render(su.device_handle(), first, count), "zygpu_render")
    film = device_film_tensor(width, height)
    if reduce and world > 1:
        reduce_film(0)
    synchronize()
    return film


def reduce_host_films(film: np.ndarray, rank: int, world: int) -> np.ndarray:
    """The same reduce for host films (CPU tests with the gloo backend)."""
    import torch
    import torch.distributed as dist

    t = torch.from_numpy(np.ascontiguousarray(film))
    if world > 1:
        dist.reduce(t, dst=0, op=dist.ReduceOp.SUM)
    return t.numpy()
