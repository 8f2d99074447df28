"""Renders the Cornell box (config 1) on the GPU and with the CPU oracle and compares the films.
Test/diagnostic tool: the oracle is only the checker here."""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

from png import write_png  # noqa: E402
from zyg_b200 import scenes, su  # noqa: E402


import oracle_lib as oracle  # noqa: E402


def main():
    w = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    spp = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    filt = sys.argv[3] if len(sys.argv) > 3 and sys.argv[3] != "none" else None
    out = sys.argv[4] if len(sys.argv) > 4 else os.path.join(ROOT, "gpurun_out")
    os.makedirs(out, exist_ok=True)

    which = os.environ.get("SCENE", "cornell")
    num_meshes = 0
    if "cornell" == which:
        scenes.cornell_box(w, w, spp=spp, filter_name=filt)
    elif "instanced" == which:
        g = int(os.environ.get("GRID", "24"))
        num_meshes = scenes.instanced_scene(w, w, spp=spp, filter_name=filt, grid=(g, g), prototypes=int(os.environ.get("PROTOS", "4")),
                                            quads=tuple(int(q) for q in os.environ.get("QUADS", "60,30").split(",")))
    else:
        num_meshes = scenes.sphere_scene(w, w, spp=spp, filter_name=filt, quads=tuple(int(q) for q in os.environ.get("QUADS", "200,100").split(",")))
    scene, view = su.compile_scene()

    t = time.time()
    film_ref = oracle.render(scene, view, w, w, 0, spp, num_meshes=num_meshes)
    t_ref = time.time() - t

    t = time.time()
    su.render_frame(0)
    t_gpu = time.time() - t
    t = time.time()
    su.render_frame(0)
    t_gpu2 = time.time() - t

    from zyg_b200 import lib
    L = lib.load_library()
    L.zygpu_download_film.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
    film_gpu = np.zeros((w, w, 4), np.float32)
    assert 0 == L.zygpu_download_film(su.device_handle(), film_gpu.ctypes.data, w * w)
    rgba = su.resolve_frame_to_buffer(w, w)
    write_png(os.path.join(out, "cornell_gpu.png"), rgba)
    rgba_ref = oracle.resolve(view, film_ref)
    write_png(os.path.join(out, "cornell_ref.png"), rgba_ref)

    d = np.abs(film_gpu - film_ref)
    rel = d[..., :3].sum(-1) / np.maximum(np.abs(film_ref[..., :3]).sum(-1), 1e-6)
    print(f"{w}x{w} {spp} spp filter={filt}: oracle {t_ref:.2f}s, gpu first {t_gpu:.3f}s second {t_gpu2:.3f}s "
          f"({w * w * spp / t_gpu2 / 1e6:.1f} Msamples/s incl. compile+upload)")
    print("weights equal:", np.array_equal(film_gpu[..., 3], film_ref[..., 3]) if filt is None else
          float(np.abs(film_gpu[..., 3] - film_ref[..., 3]).max()))
    print("bit-identical pixels:", float((d[..., :3].max(-1) == 0).mean()))
    print("max abs diff", float(d.max()), "mean rel", float(rel.mean()), "p99 rel", float(np.percentile(rel, 99)),
          "max rel", float(rel.max()))
    print("pixels with rel > 1e-3:", int((rel > 1e-3).sum()), "of", w * w)
    print("mean gpu", film_gpu[..., :3].mean(), "mean ref", film_ref[..., :3].mean())
    su.release()


if __name__ == "__main__":
    main()
