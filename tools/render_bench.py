"""Times the device render pass (zygpu_render, scene already uploaded) on the path-tracing scenes of bench.py.
Diagnostic tool for tuning runs; bench.py carries the reported numbers.
usage: tools/render_bench.py [scene-name-substring ...]   (REPS=n; any ZYGPU_* tuning variable applies)"""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
from zyg_b200 import lib, su  # noqa: E402


class Stats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("camera_samples", "closest_rays", "shadow_rays", "kernel_launches", "passes", "overflow_retries")]


def run(name, builder, kwargs, w, h, spp, reps):
    su.release()
    t0 = time.time()
    bench.build_render_scene(builder, kwargs, w, h, spp)
    L = lib.load_library()
    L.zygpu_render.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32]
    L.zygpu_synchronize.argtypes = [C.c_void_p]
    L.zygpu_clear_film.argtypes = [C.c_void_p]
    L.zygpu_render_stats.argtypes = [C.c_void_p, C.POINTER(Stats)]
    su.render_frame_range(0, 0, min(spp, 2))  # compile + upload + warm-up
    build_s = time.time() - t0
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
    samples = w * h * spp
    print(f"{name}: {best * 1e3:.1f} ms  {samples / best / 1e6:.1f} Msamples/s  "
          f"closest {st.closest_rays / samples:.2f}/sample shadow {st.shadow_rays / samples:.2f}/sample "
          f"-> {(st.closest_rays + st.shadow_rays) / best / 1e6:.0f} Mrays/s  launches {st.kernel_launches} passes {st.passes} "
          f"retries {st.overflow_retries} (build {build_s:.1f}s)", flush=True)
    if os.environ.get("COUNT", "1") != "0":
        class Counts(C.Structure):
            _fields_ = [(n, C.c_uint64) for n in ("nodes", "triangles", "props", "node_steps", "triangle_steps", "prop_steps")]

        L.zygpu_set_counting.argtypes = [C.c_void_p, C.c_int]
        L.zygpu_traversal_counts.argtypes = [C.c_void_p, C.POINTER(Counts), C.POINTER(Counts)]
        if 0 == L.zygpu_set_counting(dev, 1):
            L.zygpu_clear_film(dev)
            assert 0 == L.zygpu_render(dev, 0, spp)
            a, b = Counts(), Counts()
            if 0 == L.zygpu_traversal_counts(dev, C.byref(a), C.byref(b)):
                st2 = Stats()
                L.zygpu_render_stats(dev, C.byref(st2))
                for kind, c, rays in (("closest", a, st2.closest_rays), ("shadow", b, st2.shadow_rays)):
                    if 0 == rays or 0 == c.node_steps:
                        continue
                    steps = c.node_steps + c.triangle_steps + c.prop_steps
                    print(f"    {kind}: per ray {c.nodes / rays:.1f} nodes {c.triangles / rays:.1f} tris {c.props / rays:.2f} props "
                          f"= {(c.nodes * 80 + c.triangles * 64 + c.props * 32) / rays:.0f} B | lanes per step: node {c.nodes / max(1, c.node_steps):.1f} "
                          f"tri {c.triangles / max(1, c.triangle_steps):.1f} prop {c.props / max(1, c.prop_steps):.1f} | "
                          f"warp steps per ray {steps / rays:.2f} (node {c.node_steps / rays:.2f} tri {c.triangle_steps / rays:.2f} prop {c.prop_steps / rays:.2f})",
                          flush=True)
            L.zygpu_set_counting(dev, 0)
    su.release()


if __name__ == "__main__":
    wanted = sys.argv[1:]
    reps = int(os.environ.get("REPS", "3"))
    for name, (builder, kwargs, w, h, spp, _cpu) in bench.RENDER_SCENES.items():
        if wanted and not any(x in name for x in wanted):
            continue
        run(name, builder, kwargs, w, h, spp, reps)
