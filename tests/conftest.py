import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build the product library and the oracle once per session if they are missing."""
    from zyg_b200 import lib

    import oracle_lib

    if not os.path.exists(lib.LIB_PATH) or not os.path.exists(oracle_lib.ORACLE_PATH):
        import __graft_entry__

        __graft_entry__.build()
    yield


@pytest.fixture(scope="session")
def sphere_mesh():
    """40 000-triangle displaced sphere, compiled once."""
    from zyg_b200 import lib, scenes

    positions, normals, uvs, indices = scenes.displaced_sphere(200, 100)
    mesh = lib.Mesh(positions, indices, normals, uvs)
    return mesh, positions, indices


@pytest.fixture(scope="session")
def device():
    from zyg_b200 import lib

    dev = lib.Device(0)
    yield dev
    dev.close()
