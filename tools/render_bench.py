"""Times the device render pass (zygpu_render, scene already uploaded) on config 1 (Cornell) and on the 1M-triangle
sphere scene. Diagnostic tool; bench.py carries the reported numbers."""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from zyg_b200 import lib, scenes, su  # noqa: E402


class Stats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("camera_samples", "closest_rays", "shadow_rays", "kernel_launches", "passes")]


def run(name, build, w, spp, reps):
    su.release()
    build()
    L = lib.load_library()
    L.zygpu_render.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32]
    L.zygpu_synchronize.argtypes = [C.c_void_p]
    L.zygpu_clear_film.argtypes = [C.c_void_p]
    L.zygpu_render_stats.argtypes = [C.c_void_p, C.POINTER(Stats)]
    su.render_frame_range(0, 0, min(spp, 4))  # compile + upload + warm-up
    dev = su.device_handle()
    times = []
    for _ in range(reps):
        L.zygpu_clear_film(dev)
        L.zygpu_synchronize(dev)
        t = time.perf_counter()
        assert 0 == L.zygpu_render(dev, 0, spp)
        assert 0 == L.zygpu_synchronize(dev), L.zygpu_last_error()
        times.append(time.perf_counter() - t)
    st = Stats()
    L.zygpu_render_stats(dev, C.byref(st))
    best = min(times)
    samples = w * w * spp
    print(f"{name}: {w}x{w} x {spp} spp: {best * 1e3:.1f} ms  {samples / best / 1e6:.1f} Msamples/s  "
          f"closest {st.closest_rays / samples:.2f}/sample shadow {st.shadow_rays / samples:.2f}/sample "
          f"-> {(st.closest_rays + st.shadow_rays) / best / 1e6:.0f} Mrays/s  launches {st.kernel_launches} passes {st.passes}",
          flush=True)
    su.release()


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    reps = int(os.environ.get("REPS", "3"))
    if which in ("all", "cornell"):
        run("cornell", lambda: scenes.cornell_box(512, 512, spp=64), 512, 64, reps)
    if which in ("all", "instanced"):
        g = int(os.environ.get("GRID", "100"))
        run("instanced", lambda: scenes.instanced_scene(1024, 1024, spp=16, grid=(g, g), prototypes=int(os.environ.get("PROTOS", "20")),
                                                        quads=tuple(int(q) for q in os.environ.get("QUADS", "500,250").split(","))), 1024, 16, reps)
    if which in ("all", "sphere"):
        run("sphere1M", lambda: scenes.sphere_scene(1024, 1024, spp=16, quads=(1000, 500)), 1024, 16, reps)
