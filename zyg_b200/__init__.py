"""zyg_b200 — B200-native backend for zyg's surface-integration path.

Python here is a thin ctypes mirror of the C ABI in ``include/`` (the same role
``src/capi-test/test.py`` plays for the reference's ``libzyg``); all work happens in
``libzyg_b200.so`` (C++ scene compile + sm_100a CUDA kernels). There is no CPU fallback:
importing :mod:`zyg_b200.lib` raises if the shared library has not been built.
"""

from .lib import Device, Mesh, load_library  # noqa: F401
