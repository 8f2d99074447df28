"""Code-generation guards for the render kernels (no GPU: reads the objects `build()` leaves in zyg_b200/csrc/build).

The shade kernels take the scene, the view and the path arrays as by-value kernel parameters. If a helper that receives them
by reference ends up out of line (a `__noinline__`, or a large callback the compiler outlines with its closure), every thread
first copies the parameter structs to local memory — 40-65 STL.128 in the kernel prologue — and reads them back from there:
measured 0.8 -> 1.6 ms per shade_a launch on the instanced scene. These tests keep that from coming back."""

import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "zyg_b200", "csrc", "build")
OBJS = [os.path.join(BUILD, "device_render.o"), os.path.join(BUILD, "device_render_trace.o")]  # shading stages, traversal stages
LOGS = [o[:-2] + ".ptxas.log" for o in OBJS]
OBJ = OBJS[0]

pytestmark = pytest.mark.skipif(not (all(os.path.exists(o) for o in OBJS) and shutil.which("cuobjdump") and shutil.which("c++filt")),
                                reason="needs the built device_render.o and the CUDA binary utilities")


def kernels():
    sass = "".join(subprocess.run(["cuobjdump", "-sass", o], capture_output=True, text=True, check=True).stdout for o in OBJS)
    out, name, n = {}, None, 0
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name, n = m.group(1), 0
            out[name] = {"prologue_stl": 0, "instructions": 0}
            continue
        if name and re.search(r"/\*[0-9a-f]{4,}\*/", line):
            n += 1
            out[name]["instructions"] += 1
            if n <= 300 and "STL" in line:
                out[name]["prologue_stl"] += 1
    names = subprocess.run(["c++filt"], input="\n".join(out), capture_output=True, text=True, check=True).stdout.splitlines()
    return {pretty: out[mangled] for pretty, mangled in zip(names, out)}


def test_kernel_parameters_stay_out_of_local_memory():
    ks = kernels()
    hot = {k: v for k, v in ks.items() if re.search(r"shade[AB]Kernel|lightSamplePersistent|topKernel|meshTracePersistent", k)}
    assert len(hot) >= 20
    for name, info in hot.items():
        assert info["prologue_stl"] <= 28, f"{name}: {info['prologue_stl']} local stores in the prologue (parameter structs copied to the stack?)"


def test_hot_shade_kernels_do_not_spill():
    text = "".join(open(log).read() for log in LOGS)
    entries = re.findall(r"Compiling entry function '([^']+)' for 'sm_100a'\n.*\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads",
                         text)
    names = subprocess.run(["c++filt"], input="\n".join(e[0] for e in entries), capture_output=True, text=True, check=True).stdout.splitlines()
    spills = {n: int(e[2]) + int(e[3]) for n, e in zip(names, entries)}
    # shade_a calls the out-of-line lightImportance (device/render.cu): the values live across those calls are saved around them
    for key, limit in (("shadeAKernel<0u>", 160), ("shadeBKernel<false, false>", 16), ("meshTracePersistent<true, 8>", 16)):
        match = [n for n in spills if key in n]
        assert match, key
        assert spills[match[0]] <= limit, f"{match[0]} spills {spills[match[0]]} bytes"
