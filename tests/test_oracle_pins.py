"""Pins for the oracle: the only externally published known answers on this path.

The reference ships no tests; PCG32 is a published generator whose demo vector
(pcg32-demo, seed 42 / stream 54) pins rnd.Generator (src/base/random/generator.zig:1-47).
"""

import numpy as np
import pytest

import oracle_lib as oracle
from zyg_b200 import scenes


def test_pcg32_published_vector():
    got = oracle.pcg32_uints(42, 54, 6)
    want = np.array([0xA15C02B7, 0x7B47F409, 0xBA1D3330, 0x83D2F293, 0xBFA4784B, 0xCBED606E], np.uint32)
    assert np.array_equal(got, want)


def test_pcg32_float_mantissa_trick():
    u = oracle.pcg32_uints(7, 3, 64)
    f = oracle.pcg32_floats(7, 3, 64)
    want = ((u & 0x007FFFFF) | 0x3F800000).view(np.float32) - np.float32(1.0)
    assert np.array_equal(f, want)
    assert (f >= 0).all() and (f < 1).all()


def test_host_pcg32_matches_oracle():
    seqs = np.array([0, 1, 2, 12345, 2**40 + 17], np.uint64)
    g = scenes.PCG32(0, seqs)
    draws = np.stack([g.uint() for _ in range(16)], axis=1)
    for row, s in zip(draws, seqs):
        assert np.array_equal(row, oracle.pcg32_uints(0, int(s), 16))


# ---- golden data the reference ships for the shading path -----------------------------------------------
# tests/golden/sobol_directions.npy and zyg_b200/data/ggx_luts.f32 are extracted from the reference by
# tools/extract_reference_tables.py (data tables, not code).

import os  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sobol_directions_match_reference_table():
    """The oracle regenerates the direction numbers from the Joe-Kuo recurrence; sobol.zig:194-245 is the pin."""
    golden = np.load(os.path.join(ROOT, "tests", "golden", "sobol_directions.npy"))
    assert np.array_equal(oracle.sobol_directions(), golden)


def test_ggx_restatement_reproduces_reference_E_m_table():
    """ggx_integral.zig's E_m table is integrate_micro_directional_albedo (ggx_integrate.zig:27-57, 207-255) over
    ggx.Iso.reflect with 1024 Hammersley points. Recomputing it through the oracle's VNDF sampling, distribution,
    visibility and pdf pins that code against numbers the reference holds (printed with 8 decimals)."""
    luts = np.fromfile(os.path.join(ROOT, "zyg_b200", "data", "ggx_luts.f32"), np.float32)
    e_m = luts[:1024].reshape(32, 32)
    step = np.float32(1.0 / 31.0)
    got = np.empty((32, 32), np.float32)
    alpha = np.float32(0.0)
    for a in range(32):
        n_dot_wo = np.float32(0.0)
        for i in range(32):
            got[a, i] = oracle.ggx_micro_directional_albedo(float(alpha), float(n_dot_wo))
            n_dot_wo = np.float32(n_dot_wo + step)
        alpha = np.float32(alpha + step)
    assert np.abs(got - e_m).max() < 5e-7


def test_glass_lobes_reproduce_reference_E_s_table():
    """ggx_integral.zig's E_s table is integrate_f_s_ss (ggx_integrate.zig:134-205, 422-484): reflectNoFresnel +
    refractNoFresnel over the VNDF samples of 1024 Hammersley points, weighted by schlick1, for ior = F0ToIor(f0).
    Recomputing it through the oracle pins the rough-dielectric lobes Glass samples and evaluates
    (ggx.zig:161-233, 252-257, 441-449) against numbers the reference holds."""
    luts = np.fromfile(os.path.join(ROOT, "zyg_b200", "data", "ggx_luts.f32"), np.float32)
    e_s = luts[1056 + 4096 + 256:].reshape(16, 16, 16)
    one, two = np.float32(1.0), np.float32(2.0)
    step = np.float32(1.0 / 15.0)
    got = np.empty((16, 16, 16), np.float32)
    f0 = np.float32(0.0)
    for z in range(16):
        r = np.sqrt(f0, dtype=np.float32)
        ior = np.float32(((-f0 - one) / (f0 - one)) - ((two * r) / ((r - one) * (r + one))))  # Schlick.F0ToIor, fresnel.zig:25-28
        alpha = np.float32(0.0)
        for a in range(16):
            n_dot_wo = np.float32(0.0)
            for i in range(16):
                got[z, a, i] = oracle.ggx_f_s_ss(float(alpha), float(f0), float(ior), float(n_dot_wo))
                n_dot_wo = np.float32(n_dot_wo + step)
            alpha = np.float32(alpha + step)
        f0 = np.float32(f0 + np.float32(0.25) * step)
    assert np.abs(got - e_s).max() < 5e-7


def test_sobol_owen_scrambled_stratification():
    """Nested uniform scrambling keeps the (0, m, 1)-net property: the first 2^m samples of a dimension fall into
    distinct strata of width 2^-m (sobol.zig:36-60)."""
    for seed in (0, 1, 77):
        for m in (4, 6, 8):
            n = 1 << m
            firsts = np.array([oracle.sobol_stream(s, seed, 5) for s in range(n)])
            for dim in range(5):
                strata = np.floor(firsts[:, dim].astype(np.float64) * n).astype(int)
                assert len(set(strata.tolist())) == n, (seed, m, dim)


def test_sobol_padding_blocks_differ():
    a = oracle.sobol_stream(5, 0, 10)
    b = oracle.sobol_stream(5, 0, 10, pad_every=5)
    assert np.array_equal(a, b)  # a refill after 5 draws happens with or without incrementPadding
    c = oracle.sobol_stream(5, 0, 6, pad_every=1)
    assert len(set(c.tolist())) == 6 and c[1] != a[1]  # padding after every draw: each draw is dim 0 of a new block


def test_multiscatter_term_reproduces_reference_E_table():
    """ggx_integral.zig's E table is integrate_directional_albedo (ggx_integrate.zig:89-116, 300-368): ggx.Iso.reflect with
    Schlick(f0) plus dspbrMicroEc over the E_m / E_m_avg tables, clamped to 1. Recomputing it through the oracle pins the
    Schlick term, the multi-scatter compensation and the bilinear / linear table evaluation of the Substitute lobe."""
    luts = np.fromfile(os.path.join(ROOT, "zyg_b200", "data", "ggx_luts.f32"), np.float32)
    e = luts[1056:1056 + 4096].reshape(16, 16, 16)
    step = np.float32(1.0 / 15.0)
    got = np.empty((16, 16, 16), np.float32)
    f0 = np.float32(0.0)
    for z in range(16):
        alpha = np.float32(0.0)
        for a in range(16):
            n_dot_wo = np.float32(0.0)
            for i in range(16):
                got[z, a, i] = oracle.ggx_directional_albedo(luts, float(alpha), float(f0), float(n_dot_wo))
                n_dot_wo = np.float32(n_dot_wo + step)
            alpha = np.float32(alpha + step)
        f0 = np.float32(f0 + step)
    assert np.abs(got - e).max() < 5e-7, np.abs(got - e).max()  # measured 2.4e-7, 2946 / 4096 entries equal at 8 decimals


def test_trilinear_lookup_reproduces_reference_E_avg_table():
    """E_avg is integrate_average_albedo (ggx_integrate.zig:118-132, 370-420): the mean of E.eval over 1024 cosine-distributed
    Hammersley directions, capped at 0.9997 — a pin for InterpolatedFunction3D.eval and smpl.hemisphereCosine."""
    luts = np.fromfile(os.path.join(ROOT, "zyg_b200", "data", "ggx_luts.f32"), np.float32)
    e_avg = luts[1056 + 4096:1056 + 4096 + 256].reshape(16, 16)
    step = np.float32(1.0 / 15.0)
    got = np.empty((16, 16), np.float32)
    f0 = np.float32(0.0)
    for a in range(16):
        alpha = np.float32(0.0)
        for i in range(16):
            got[a, i] = min(oracle.ggx_average_albedo(luts, float(alpha), float(f0)), np.float32(0.9997))
            alpha = np.float32(alpha + step)
        f0 = np.float32(f0 + step)
    assert np.abs(got - e_avg).max() < 2e-7, np.abs(got - e_avg).max()  # measured 6e-8, 244 / 256 entries equal


def test_bilinear_lookup_reproduces_reference_E_m_avg_table():
    """E_m_avg is integrate_micro_average_albedo (ggx_integrate.zig:59-73, 259-298): the mean of E_m.eval over 1024
    cosine-distributed Hammersley directions — a pin for InterpolatedFunction2D.eval."""
    luts = np.fromfile(os.path.join(ROOT, "zyg_b200", "data", "ggx_luts.f32"), np.float32)
    e_m_avg = luts[1024:1056]
    step = np.float32(1.0 / 31.0)
    alpha = np.float32(0.0)
    got = np.empty(32, np.float32)
    for a in range(32):
        got[a] = oracle.ggx_micro_average_albedo(luts, float(alpha))
        alpha = np.float32(alpha + step)
    assert np.abs(got - e_m_avg).max() < 2e-7, np.abs(got - e_m_avg).max()


def test_reference_record_sizes_are_pinned_in_the_abi_header(tmp_path):
    """The records the device ABI shares with the reference byte for byte carry the sizes the reference checks for itself
    (src/core/size_test.zig:36-47: ComposedTransformation 64, BvhNode 32, LightNode 32): static assertions in include/zygpu_scene.h,
    compiled here as C and as C++."""
    import shutil
    import subprocess

    include = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include")
    for compiler, name in (("gcc", "t.c"), ("g++", "t.cpp")):
        if not shutil.which(compiler):
            pytest.skip("no host compiler")
        src = tmp_path / name
        src.write_text('#include "zygpu_scene.h"\n#include "zygpu.h"\n#include "zyg_su.h"\nint main(void) { return 0; }\n')
        subprocess.run([compiler, "-I", include, "-fsyntax-only", str(src)], check=True)
