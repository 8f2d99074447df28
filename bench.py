#!/usr/bin/env python
"""bench.py — BASELINE.json config 2: ray-intersection microbench on the 1M-triangle displaced sphere.

One step = one pass of the traversal hot path over the synthetic batch of this rank:
    closest-hit over 16 777 216 primary rays      (coherent, 4096 x 4096 pinhole)
  + closest-hit over 16 777 216 incoherent rays   (PCG32 origins in the ball, uniform directions)
  + any-hit     over 16 777 216 shadow segments   (same origins, segment to a second point)
`value` = rays of all ranks / device time (CUDA events, max over ranks), inputs resident in HBM.
`e2e`   = same step through zygpu_trace_batch with pinned HOST buffers (H2D + D2H inside).
`--impl reference` times the CPU restatement of zyg's own path (oracle/) on the host cores.

The same JSON line carries `path_tracing`: path-samples/s of the wavefront PathtracerMIS pass (the other half of
BASELINE.json's metric) on config 1 (Cornell box, 512 x 512 x 64 spp) and on the config-2 mesh as a lit scene
(1M triangles, 1024 x 1024 x 16 spp), device-timed with the scene resident, next to the CPU path on a bounded sample.
With N > 1 ranks the samples of the frame are split by range and the films reduced to rank 0 (strong scaling).
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

PRIMARY_RES = 4096
N_RAYS = PRIMARY_RES * PRIMARY_RES
MESH_QUADS = (1000, 500)
WORKLOAD = "config2: 1M-triangle displaced sphere; 16.8M primary + 16.8M incoherent closest-hit + 16.8M shadow any-hit rays"
CLASSES = ("primary_closest", "incoherent_closest", "shadow_any")


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def build_inputs(rank: int, n_rays: int, res: int):
    from zyg_b200 import lib, scenes

    t0 = time.time()
    positions, normals, uvs, indices = scenes.displaced_sphere(*MESH_QUADS)
    mesh = lib.Mesh(positions, indices, normals, uvs)
    info = mesh.info()
    log(f"[rank {rank}] mesh: {info.num_source_triangles} triangles -> {info.num_tree_triangles} references, "
        f"{info.num_binary_nodes} binary / {info.num_wide_nodes} wide nodes, built in {time.time() - t0:.1f}s")
    t0 = time.time()
    cache = os.environ.get("ZYG_BENCH_CACHE")  # optional: reuse generated rays across tuning runs
    path = os.path.join(cache, f"rays_{res}_{n_rays}_{rank}.npy") if cache else None
    if path and os.path.exists(path):
        packed = np.load(path)
        rays = {k: packed[i] for i, k in enumerate(CLASSES)}
    else:
        rays = {
            "primary_closest": scenes.primary_rays(res, res),
            "incoherent_closest": scenes.random_rays(n_rays, first=rank * n_rays),
            "shadow_any": scenes.random_rays(n_rays, shadow=True, first=rank * n_rays),
        }
        if path:
            os.makedirs(cache, exist_ok=True)
            np.save(path, np.stack([rays[k] for k in CLASSES]))
    log(f"[rank {rank}] rays generated in {time.time() - t0:.1f}s")
    return mesh, rays


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.rows, self.proc = index, [], None

    def run(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms",
                 "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc:
            self.proc.terminate()
        self.join(timeout=2)
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i] == "Active" for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def load_oracle():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib

    oracle_lib.load()
    return oracle_lib


def cpu_step(oracle, arrays, sample):
    nodes, tris, pos = arrays
    t0 = time.perf_counter()
    oracle.trace_closest(nodes, tris, pos, sample["primary_closest"])
    oracle.trace_closest(nodes, tris, pos, sample["incoherent_closest"])
    oracle.trace_any(nodes, tris, pos, sample["shadow_any"])
    return time.perf_counter() - t0


def build_reference_inputs(n_rays: int, res: int):
    """Mesh arrays for the CPU arm, built by the oracle's own restatement of zyg's BVH builder (oracle/builders.cpp): this
    process never maps libzyg_b200.so. zyg_b200.scenes is plain numpy (procedural vertices and rays)."""
    from zyg_b200 import scenes

    oracle = load_oracle()
    t0 = time.time()
    positions, normals, uvs, indices = scenes.displaced_sphere(*MESH_QUADS)
    mesh = oracle.BuiltMesh(positions, indices, normals, uvs)
    arrays = tuple(mesh.data(w) for w in (mesh.NODES, mesh.TRIANGLES, mesh.POSITIONS))
    log(f"[reference] mesh: {indices.shape[0]} triangles -> {arrays[1].size // 3} references, {arrays[0].size} binary nodes, "
        f"built by oracle/builders.cpp in {time.time() - t0:.1f}s")
    rays = {
        "primary_closest": scenes.primary_rays(res, res),
        "incoherent_closest": scenes.random_rays(n_rays),
        "shadow_any": scenes.random_rays(n_rays, shadow=True),
    }
    return oracle, arrays, rays


def run_reference(args):
    """CPU arm: zyg's own traversal on the box's host cores (restated in oracle/: the reference is Zig and cannot be built
    here), over a tree built by the oracle's own builder. The forward-pass legs need a compiled scene, which only the
    product's host model produces: they run in a child process (`--impl reference-render`), so this process stays free of
    libzyg_b200.so, and the child re-builds the scene's prop and light trees with the oracle's builders and compares."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    oracle, arrays, rays = build_reference_inputs(N_RAYS // 4, PRIMARY_RES // 2)
    n = sum(r.shape[0] for r in rays.values())
    for _ in range(args.warmup):
        cpu_step(oracle, arrays, rays)
    times = [cpu_step(oracle, arrays, rays) for _ in range(args.steps)]
    dt = sum(times)
    value = n * args.steps / dt / 1e6
    cores = os.cpu_count()
    sample = (f"per step: {PRIMARY_RES // 2}^2 primary + {N_RAYS // 4} incoherent + {N_RAYS // 4} shadow rays (1/4 of the workload: "
              f"a rate, so the ratio to the GPU arm stands), tree built by oracle/builders.cpp")

    path_tracing = {}
    if not args.no_render:
        cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference-render"]
        if args.scenes:
            cmd += ["--scenes"] + list(args.scenes)
        env = {k: v for k, v in os.environ.items() if k not in ("RANK", "LOCAL_RANK", "WORLD_SIZE")}
        out = subprocess.run(cmd, stdout=subprocess.PIPE, text=True, env=env)
        lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
        if 0 == out.returncode and lines:
            path_tracing = json.loads(lines[-1])
        else:
            path_tracing = {"unavailable": f"child process returned {out.returncode}"}

    loaded = sorted({l.split()[-1] for l in open("/proc/self/maps") if l.rstrip().endswith(".so") and ("zyg" in l)})
    print(json.dumps({
        "impl": "reference", "metric": "traversal_throughput", "value": value, "unit": "Mrays/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "rays_per_step": n, "sample": sample},
        "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "native_libraries": [os.path.basename(x) for x in loaded],
        "path_tracing": path_tracing,
    }))


def run_reference_render(args):
    """Child of the reference arm: the CPU path of the forward pass on a bounded sample of the render workloads. The scene is
    compiled by the product's host model (there is no other source of a compiled scene); its prop trees and light tree are
    re-built by the oracle's own builders and must be identical before the oracle renders."""
    from zyg_b200 import su

    oracle = load_oracle()
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import scene_view as sv

    out = {}
    cores = os.cpu_count()
    for name, (builder, kwargs, width, height, spp, cpu_spp) in RENDER_SCENES.items():
        if (args.scenes and name not in args.scenes) or 0 == cpu_spp:
            continue
        num_meshes = build_render_scene(builder, kwargs, width, height, spp)
        scene, view = su.compile_scene()
        identical = oracle_rebuild_matches(oracle, sv, scene)
        t0 = time.perf_counter()
        oracle.render(scene, view, width, height, 0, cpu_spp, num_meshes=num_meshes)
        dt_r = time.perf_counter() - t0
        out[name] = {"path_samples_per_s": width * height * cpu_spp / dt_r, "cores": cores,
                     "sample": f"{width}x{height} x {cpu_spp} spp of the same scene",
                     "scene_compiled_by": "product host model (child process)", "oracle_rebuilt_trees_identical": identical}
        su.release()
    print(json.dumps(out))


def oracle_rebuild_matches(oracle, sv, scene_address) -> bool:
    """Prop trees and the scene light tree of a compiled scene, re-built by oracle/builders.cpp from the scene's primary records."""
    s = sv.scene_at(scene_address)
    props = sv.view(s.props, sv.PROP_DTYPE, s.num_props)
    aabbs = sv.view(s.aabbs, sv.AABB_DTYPE, s.num_props)
    infinite = np.isin(props["shape"], (sv.SHAPE_CANOPY, sv.SHAPE_DISTANT, sv.SHAPE_DOME))
    unocc = (props["flags"] & sv.PROP_UNOCCLUDING) != 0
    ok = True
    for tree, want in ((s.solid_bvh, False), (s.unoccluding_bvh, True)):
        idx = sv.view(tree.indices, "<u4", tree.num_indices)
        members = np.zeros(s.num_props, bool)
        members[idx] = True
        ids = np.nonzero(members & ~infinite & (unocc == want))[0].astype(np.uint32)
        nodes, indices = oracle.build_prop_tree(ids, aabbs)
        ok = ok and nodes == sv.view(tree.nodes, sv.NODE_DTYPE, tree.num_nodes).tobytes() and indices == idx.tobytes()
    if s.num_lights > 0:
        lights = sv.view(s.lights, sv.LIGHT_DTYPE, s.num_lights)
        finite = ~np.isin(props["shape"][lights["prop"]], (sv.SHAPE_CANOPY, sv.SHAPE_DISTANT, sv.SHAPE_DOME))
        mine = oracle.build_light_tree(sv.view(s.light_aabbs, sv.AABB_DTYPE, s.num_lights), sv.view(s.light_cones, "<f4", 4 * s.num_lights),
                                       lights["two_sided"] != 0, finite)
        t = s.light_tree
        ok = ok and mine["nodes"] == sv.view(t.nodes, sv.LIGHT_NODE_DTYPE, t.num_nodes).tobytes()
        ok = ok and mine["mapping"] == sv.view(t.light_mapping, "<u4", t.num_lights).tobytes()
    return bool(ok)


# The forward pass on the BASELINE.json configs: name -> (scene builder, kwargs, width, height, spp per step, spp of the CPU
# sample). The CPU path renders the same compiled scene at the same resolution with fewer samples per pixel.
RENDER_SCENES = {
    # configs[0]: Cornell box 512 x 512 x 64 spp, PathtracerMIS, max 8 bounces
    "config1_cornell_512x512x64": ("cornell_box", {}, 512, 512, 64, 8),
    # the configs[1] mesh (1M triangles) as a lit scene
    "sphere1m_1024x1024x16": ("sphere_scene", {"quads": MESH_QUADS}, 1024, 1024, 16, 2),
    # configs[2]: 20 x 250k-triangle prototypes = 5M triangles, 10k prop instances, diffuse + rough metal + glass, Rectangle
    # light + Distant sun, 1920 x 1080 (8 of the 256 spp per step)
    "config3_instanced5m_1920x1080x8": ("instanced_scene", {"grid": (100, 100), "prototypes": 20, "quads": (500, 250), "sun": 60.0},
                                        1920, 1080, 8, 1),
    # configs[4]: the instanced scene at 3840 x 2160, the frame's samples split by range over the ranks and the 132.7 MB film
    # reduced to rank 0 (8 of the 4096 spp per step; 8 / N per GPU)
    "config5_instanced5m_3840x2160x8": ("instanced_scene", {"grid": (100, 100), "prototypes": 20, "quads": (500, 250), "sun": 60.0},
                                        3840, 2160, 8, 0),
    # configs[3]: 1000 emissive icosahedron meshes + a 576-triangle emitter in a room with 200k triangles of diffuse geometry,
    # sky = a 1024^2 radiance image on a Canopy (procedural: the arpraguesky dataset is not shipped) + Distant sun, light-tree
    # sampling with split threshold 0.5, 1920 x 1080 (8 of the 1024 spp per step, so that 8 ranks have a sample each)
    "config4_meshlights1k_sky_1920x1080x8": ("mesh_lights_scene", {"num_lights": 1000, "geometry_quads": (400, 250), "sun": 15.0, "sky": 1024, "max_depth": 8},
                                         1920, 1080, 8, 1),
}
MESH_SCENES = ("sphere_scene", "instanced_scene", "mesh_lights_scene")

# Path-state traffic of the wavefront stages per ray, in bytes (the 16-byte SoA words of device/render.cuh):
# a path vertex = one closest-hit ray: extend reads origin + direction (32) and writes max_t + hit (32); shade_a reads the six
# vertex words (96), the sampler (24) and the accumulators it adds to (32 read + write), writes throughput (16) and sampler (24);
# shade_b reads the vertex words again (96), sampler (24), accumulator (32), writes the next vertex (80) and sampler (24);
# queue entries 3 x 4.
STATE_BYTES_PER_VERTEX = 32 + 32 + (96 + 24 + 32 + 16 + 24) + (96 + 24 + 32 + 80 + 24) + 12
# a shadow ray: shade_a writes the record (48), the shadow stage reads 32 of it and writes the visibility flag (4), shade_b reads it
# (48); one queue entry
STATE_BYTES_PER_SHADOW_RAY = 48 + 32 + 4 + 48 + 4


def build_render_scene(builder, kwargs, width, height, spp):
    from zyg_b200 import scenes, su

    su.release()
    r = getattr(scenes, builder)(width, height, spp=spp, **kwargs)
    return r if builder in MESH_SCENES else 0


def bench_render(args, rank, world, local):
    """Path-samples/s of the device pass: per step clear + this rank's sample range + film reduce to rank 0."""
    import ctypes as C

    import torch
    import torch.distributed as dist

    from zyg_b200 import lib, multi, su

    L = lib.load_library()
    L.zygpu_clear_film.argtypes = [C.c_void_p]
    L.zygpu_render.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32]
    L.zygpu_synchronize.argtypes = [C.c_void_p]

    class Stats(C.Structure):
        _fields_ = [(n, C.c_uint64) for n in ("camera_samples", "closest_rays", "shadow_rays", "kernel_launches", "passes", "overflow_retries")]

    class Counts(C.Structure):
        _fields_ = [(n, C.c_uint64) for n in ("nodes", "triangles", "props", "node_steps", "triangle_steps", "prop_steps")]

    L.zygpu_render_stats.argtypes = [C.c_void_p, C.POINTER(Stats)]
    L.zygpu_set_counting.argtypes = [C.c_void_p, C.c_int]
    L.zygpu_traversal_counts.argtypes = [C.c_void_p, C.POINTER(Counts), C.POINTER(Counts)]
    L.zygpu_download_film.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
    peak, peak_src = peak_hbm()
    traffic_file = os.path.join(ROOT, "profiles", "traffic.json")
    traffic = json.load(open(traffic_file)) if os.path.exists(traffic_file) else {}

    out = {}
    launches = 0
    for name, (builder, kwargs, width, height, spp, cpu_spp) in RENDER_SCENES.items():
        if args.scenes and name not in args.scenes:
            continue
        t_build = time.time()
        num_meshes = build_render_scene(builder, kwargs, width, height, spp)
        su._ok(su._su().zyg_su_set_device(local), "zyg_su_set_device")
        su.start_frame(0)  # Scene.compile + upload + clear (not timed: resident-scene number)
        dev = su.device_handle()
        stream = multi.render_stream()
        first, count = multi.sample_range(rank, world, spp)
        film = multi.device_film_tensor(width, height)
        log(f"[rank {rank}] {name}: scene built, compiled and uploaded in {time.time() - t_build:.1f}s")

        def step():
            su._ok(L.zygpu_clear_film(dev), "zygpu_clear_film")
            if count > 0:
                su._ok(L.zygpu_render(dev, first, count), "zygpu_render")
            if world > 1:
                multi.reduce_film(0)  # zygpu_reduce_film: ncclReduce(sum, fp32) on the render stream behind the passes

        def barrier():
            su._ok(L.zygpu_synchronize(dev), "zygpu_synchronize")
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()

        for _ in range(args.warmup):
            step()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record()
        for _ in range(args.steps):
            step()
        with torch.cuda.stream(stream):
            e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        st = Stats()
        L.zygpu_render_stats(dev, C.byref(st))

        # the film of the timed step, for the checks below (rank 0 holds the reduced film when world > 1)
        timed_film = np.empty((height, width, 4), np.float32)
        su._ok(L.zygpu_download_film(dev, timed_film.ctypes.data, width * height), "zygpu_download_film")

        # end to end through zyg's C API: su_render_frame (scene unchanged since the last frame: no recompile, no upload) +
        # su_resolve_frame_to_buffer (resolve on the device, RGBA fp32 to the host)
        t0 = time.perf_counter()
        if count > 0:
            su.render_frame_range(0, first, count)
        rgba = su.resolve_frame_to_buffer(width, height)
        e2e_s = time.perf_counter() - t0

        reduce_ms = 0.0
        if world > 1:  # the film reduce on its own (it is inside `ms` as well)
            barrier()
            r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(stream):
                r0.record()
                for _ in range(args.steps):
                    multi.reduce_film(0)
                r1.record()
            barrier()
            reduce_ms = r0.elapsed_time(r1) / args.steps
            t = torch.tensor([ms, e2e_s, reduce_ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms, e2e_s, reduce_ms = t.tolist()

        samples = width * height * spp
        entry = {
            "path_samples_per_s": samples * args.steps / (ms * 1e-3), "ms_per_frame": ms / args.steps,
            "resolution": [width, height], "spp": spp, "samples_per_frame": samples,
            "closest_rays_per_sample": st.closest_rays / max(1, st.camera_samples),
            "shadow_rays_per_sample": st.shadow_rays / max(1, st.camera_samples),
            "mrays_per_s": (st.closest_rays + st.shadow_rays) / max(1, st.camera_samples) * samples * args.steps / (ms * 1e-3) / 1e6,
            "e2e_path_samples_per_s": samples / e2e_s, "e2e_d2h_bytes": int(rgba.nbytes),
            "film_reduce_ms": reduce_ms, "film_bytes": width * height * 16,
            "scaling": "strong (sample-range split, film reduce to rank 0)" if world > 1 else "single device",
        }
        launches += int(st.kernel_launches)

        # Roofline of the render path (SURVEY.md §8d "B_sample"): algorithmic bytes per path sample = counted traversal fetches
        # of an instrumented pass over the same samples (80 B nodes, 64 B triangle records, 64 B prop records) + ray / hit /
        # shadow-record / path-state traffic of the stages per ray + 16 B of film, times the measured samples/s.
        if count > 0 and 0 == L.zygpu_set_counting(dev, 1):
            su._ok(L.zygpu_clear_film(dev), "zygpu_clear_film")
            su._ok(L.zygpu_render(dev, first, count), "zygpu_render")
            a, b = Counts(), Counts()
            st2 = Stats()
            if 0 == L.zygpu_traversal_counts(dev, C.byref(a), C.byref(b)) and 0 == L.zygpu_render_stats(dev, C.byref(st2)):
                mine = width * height * count
                trav = ((a.nodes + b.nodes) * 80 + (a.triangles + b.triangles) * 64 + (a.props + b.props) * 64) / mine
                state = (st2.closest_rays * STATE_BYTES_PER_VERTEX + st2.shadow_rays * STATE_BYTES_PER_SHADOW_RAY) / mine + 16
                b_sample = trav + state
                achieved = b_sample * entry["path_samples_per_s"] / 1e9
                entry["roofline"] = {
                    "bound": "hbm", "achieved": achieved, "peak": peak * world, "unit": "GB/s", "frac": achieved / (peak * world),
                    "bytes_per_sample": b_sample, "traversal_bytes_per_sample": trav, "state_bytes_per_sample": state,
                    "traffic": traffic.get(name), "peak_source": peak_src,
                    "counted": "fused two-level kernel" if a.node_steps + b.node_steps > 0 else "top + mesh kernels are not instrumented: state bytes only",
                    "nodes_per_closest_ray": a.nodes / max(1, st2.closest_rays), "tris_per_closest_ray": a.triangles / max(1, st2.closest_rays),
                    "props_per_closest_ray": a.props / max(1, st2.closest_rays),
                    "lanes_per_node_step": a.nodes / max(1, a.node_steps), "lanes_per_triangle_step": a.triangles / max(1, a.triangle_steps)}
            L.zygpu_set_counting(dev, 0)

        if world > 1:
            # The NCCL film path checked where N GPUs exist: rank 0 renders all samples of the frame alone and compares with
            # the film the ranks reduced (same samples, same seeds: equal up to fp32 summation order).
            check = None
            if 0 == rank:
                su._ok(L.zygpu_clear_film(dev), "zygpu_clear_film")
                su._ok(L.zygpu_render(dev, 0, spp), "zygpu_render")
                alone = np.empty((height, width, 4), np.float32)
                su._ok(L.zygpu_download_film(dev, alone.ctypes.data, width * height), "zygpu_download_film")
                err = np.abs(timed_film - alone) / np.maximum(np.abs(alone), 1e-3)
                off = int((~np.isclose(timed_film, alone, rtol=2e-6, atol=1e-6)).any(-1).sum())
                check = {"allclose_2e-6": 0 == off, "pixels_off": off, "max_rel_error": float(err.max()),
                         "weights_equal": bool(np.array_equal(timed_film[..., 3], alone[..., 3]))}
                # a broken reduce moves every pixel; a handful of pixels is a path that met two hits closer together than fp32 tells
                # apart and took the other one: reported, not fatal
                assert check["weights_equal"] and off <= max(4, width * height // 1000000), \
                    f"{name}: the reduced film differs from the single-GPU film: {check}"
            entry["nccl_film_check"] = check
            dist.barrier()

        if rank == 0 and world == 1 and not args.no_cpu and cpu_spp > 0:
            oracle = load_oracle()
            scene, view = su.compile_scene()
            t0 = time.perf_counter()
            oracle.render(scene, view, width, height, 0, cpu_spp, num_meshes=num_meshes)
            dt = time.perf_counter() - t0
            entry["cpu_baseline"] = {"value": width * height * cpu_spp / dt, "unit": "path-samples/s", "cores": os.cpu_count(), "kind": "port",
                                     "sample": f"{width}x{height} x {cpu_spp} spp of the same scene, oracle/ restatement of zyg's PathtracerMIS"}
        su.release()
        out[name] = entry
    return out, launches


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="zyg_b200", choices=["zyg_b200", "reference", "reference-render"])
    ap.add_argument("--rays", type=int, default=PRIMARY_RES, help="primary resolution r (r*r rays per class)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-render", action="store_true", help="skip the path_tracing section")
    ap.add_argument("--scenes", nargs="*", default=None, help="path_tracing scenes to run (default: all of RENDER_SCENES)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    if args.impl == "reference":
        return run_reference(args)
    if args.impl == "reference-render":
        return run_reference_render(args)

    import torch
    import torch.distributed as dist

    from zyg_b200 import lib

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    # rank 0 prints exactly one JSON line on stdout: everything else that writes to fd 1 while the bench runs (NCCL's version
    # banner, library chatter) goes to stderr
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    res = args.rays
    n_rays = res * res
    mesh, rays = build_inputs(rank, n_rays, res)
    dev = lib.Device(local)
    mid = dev.upload_mesh(mesh)
    modes = {"primary_closest": lib.CLOSEST, "incoherent_closest": lib.CLOSEST, "shadow_any": lib.ANY}

    # resident inputs / outputs (torch is the allocator and the stream owner; kernels are ours)
    d_rays = {k: torch.from_numpy(v.view(np.float32).reshape(-1, 8)).cuda() for k, v in rays.items()}
    d_out = {k: torch.empty((n_rays, 1 if modes[k] == lib.ANY else 4), dtype=torch.float32, device="cuda")
             for k in CLASSES}
    stream = torch.cuda.current_stream().cuda_stream

    def launch(k, counters=None):
        dev.trace_batch_ptr(mid, modes[k], d_rays[k].data_ptr(), n_rays, d_out[k].data_ptr(), host=False,
                            stream=stream, counters=counters)

    # fetch counts from the instrumented build of the same kernels on the same batch
    bytes_per_ray = {}
    fetches = {}
    for k in CLASSES:
        c = lib.TraceCounters()
        launch(k, c)
        out_b = 4 if modes[k] == lib.ANY else 16
        bytes_per_ray[k] = 32 + out_b + (c.nodes * 80 + c.triangles * 64) / n_rays
        fetches[k] = {"nodes_per_ray": c.nodes / n_rays, "tris_per_ray": c.triangles / n_rays,
                      "bytes_per_ray": bytes_per_ray[k], "max_stack": int(c.max_stack)}
    torch.cuda.synchronize()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        for k in CLASSES:
            launch(k)
    barrier()

    sampler = ClockSampler(local)
    sampler.start()
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(len(CLASSES) + 1)] for _ in range(args.steps)]
    for s in range(args.steps):
        ev[s][0].record()
        for i, k in enumerate(CLASSES):
            launch(k)
            ev[s][i + 1].record()
    barrier()
    total_ms = ev[0][0].elapsed_time(ev[-1][-1])
    class_ms = {k: sum(ev[s][i].elapsed_time(ev[s][i + 1]) for s in range(args.steps)) / args.steps
                for i, k in enumerate(CLASSES)}
    launches = args.steps * len(CLASSES)

    # end to end: pinned host buffers through the public C-ABI call, copies inside the timed region
    h_rays = {k: torch.from_numpy(v.view(np.float32).reshape(-1, 8)).pin_memory() for k, v in rays.items()}
    h_out = {k: torch.empty((n_rays, 1 if modes[k] == lib.ANY else 4), dtype=torch.float32).pin_memory()
             for k in CLASSES}

    def e2e_step():
        for k in CLASSES:
            dev.trace_batch_ptr(mid, modes[k], h_rays[k].data_ptr(), n_rays, h_out[k].data_ptr(), host=True)

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0
    clocks = sampler.stop()

    if world > 1:
        t = torch.tensor([total_ms, e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, e2e_s = t.tolist()

    rays_per_step = n_rays * len(CLASSES) * world
    value = rays_per_step * args.steps / (total_ms * 1e-3) / 1e6
    e2e_value = rays_per_step * args.steps / e2e_s / 1e6
    h2d = n_rays * 32 * len(CLASSES)
    d2h = n_rays * (16 + 16 + 4)

    # spot parity of what was just timed (rank 0): device result == host-path result
    if rank == 0:
        k = "incoherent_closest"
        assert torch.equal(d_out[k][: 1 << 16].cpu().view(torch.int32), h_out[k][: 1 << 16].view(torch.int32)), \
            "device and host entry points disagree"

    peak, peak_src = peak_hbm()
    dom = max(CLASSES, key=lambda k: class_ms[k])
    achieved = bytes_per_ray[dom] * n_rays / (class_ms[dom] * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp)).get(dom)

    result = {
        "metric": "traversal_throughput", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "rays_per_step_per_gpu": n_rays * len(CLASSES), "mesh_triangles": 1000000,
                   "l2": "inputs larger than L2 (512 MiB of rays per class); the 66 MB BVH is L2-resident by nature",
                   "parallelism": f"replicated scene, independent ray batches x{world}"},
        "classes": {k: {"mrays_s": n_rays / (class_ms[k] * 1e-3) / 1e6, "ms": class_ms[k], **fetches[k]}
                    for k in CLASSES},
        "roofline": {"bound": "hbm", "kernel": f"traceWide ({dom})", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                     "note": "algorithmic bytes = 32 B ray + hit record + counted 80 B node and 64 B triangle fetches"},
        "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": launches,
        "clocks": clocks,
    }

    if not args.no_render:
        dev.close()
        dev = None
        render, render_launches = bench_render(args, rank, world, local)
        result["path_tracing"] = render
        result["gpu_launches"] = launches + render_launches

    if rank == 0 and world == 1 and not args.no_cpu:
        oracle = load_oracle()
        arrays = tuple(mesh.data(w) for w in (lib.MESH_BINARY_NODES, lib.MESH_TRIANGLES, lib.MESH_POSITIONS))
        stride = max(1, n_rays // (1 << 22))
        sample = {k: np.ascontiguousarray(v[::stride]) for k, v in rays.items()}
        ns = sum(v.shape[0] for v in sample.values())
        cpu_step(oracle, arrays, {k: v[: 1 << 16] for k, v in sample.items()})
        reps, dt = 0, 0.0
        while dt < 10.0 and reps < 8:
            dt += cpu_step(oracle, arrays, sample)
            reps += 1
        result["cpu_baseline"] = {"value": ns * reps / dt / 1e6, "unit": "Mrays/s", "cores": os.cpu_count(),
                                  "kind": "port",
                                  "sample": f"every {stride}th ray of each class ({ns} rays) x {reps} passes, "
                                            f"oracle/ restatement of zyg's traversal, std::thread per core"}
        # parity of the timed output against the oracle on the sample
        ref = oracle.trace_closest(*arrays, sample["incoherent_closest"][: 1 << 18])
        got = h_out["incoherent_closest"].numpy().view(lib.HIT_DTYPE).reshape(-1)[::stride][: 1 << 18]
        assert np.array_equal(ref["t"].view(np.uint32), got["t"].view(np.uint32)), "timed output differs from oracle"

    if rank == 0:
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(result) + "\n").encode())
    if dev is not None:
        dev.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
