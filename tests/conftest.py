import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import bindings
    return bindings.COracle()


@pytest.fixture(scope="session")
def ref():
    from oracle import bindings
    if not bindings.have_ref():
        if not bindings.build_ref():
            pytest.skip("compiled reference (oracle/_ref/libref_voxel.so) not available")
    return bindings.ref()


@pytest.fixture(scope="session")
def meshes():
    """The three RecastDemo meshes staged under oracle/_ref/meshes (git-ignored)."""
    from oracle import bindings
    from recastnavigation_b200 import geom
    out = {}
    for name in ("nav_test", "dungeon", "undulating"):
        p = bindings.mesh_path(name + ".obj")
        if os.path.exists(p):
            out[name] = geom.load_obj(p)
    if not out:
        pytest.skip("demo meshes not staged (run `make -C oracle meshes` where the reference exists)")
    return out
