"""Extracts DATA tables the reference ships (no code) so the backend evaluates the same LUTs:

  * zyg_b200/data/ggx_luts.f32 — the five energy-compensation tables of
    src/core/scene/material/ggx_integral.zig (E_m 32x32, E_m_avg 32, E 16^3, E_avg 16x16, E_s 16^3),
    concatenated as little-endian float32 in that order (ZygpuScene.ggx_luts).
  * tests/golden/sobol_directions.npy — the direction numbers of src/core/sampler/sobol.zig:194-245,
    used only as a golden vector: product and oracle regenerate them from the Joe-Kuo recurrence.

Run in the build container (needs /root/reference): python tools/extract_reference_tables.py
"""
import os
import re
import sys

import numpy as np

REF = "/root/reference/src"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def floats_of(text, name):
    m = re.search(r"pub const %s = \[[^\]]*\]f32\{(.*?)\};" % re.escape(name), text, re.S)
    body = re.sub(r"//[^\n]*", "", m.group(1))
    return np.array([float(t) for t in re.findall(r"[-+]?\d+\.\d+(?:e[-+]?\d+)?", body)], dtype=np.float32)


def main():
    text = open(os.path.join(REF, "core/scene/material/ggx_integral.zig")).read()
    sizes = {"E_m": 1024, "E_m_avg": 32, "E": 4096, "E_avg": 256, "E_s": 4096}
    parts = []
    for name, n in sizes.items():
        a = floats_of(text, name)
        assert a.size == n, (name, a.size)
        parts.append(a)
    luts = np.concatenate(parts)
    luts.astype("<f4").tofile(os.path.join(ROOT, "zyg_b200/data/ggx_luts.f32"))

    sob = open(os.path.join(REF, "core/sampler/sobol.zig")).read()
    body = sob[sob.index("const Directions") :]
    words = np.array([int(w, 16) for w in re.findall(r"0x[0-9a-fA-F]{8}", body)], dtype=np.uint32)
    assert words.size == 160
    np.save(os.path.join(ROOT, "tests/golden/sobol_directions.npy"), words.reshape(5, 32))
    print("wrote", luts.size, "LUT floats and", words.size, "direction numbers")


if __name__ == "__main__":
    sys.exit(main())
