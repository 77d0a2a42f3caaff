import math
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


HAS_GPU = _has_gpu()


def pytest_collection_modifyitems(config, items):
    if HAS_GPU:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def build_libs():
    """Make sure the checker and the product libraries are built (no-op when they already are)."""
    import __graft_entry__ as ge
    ge.build()
    return True


@pytest.fixture(scope="session")
def scene():
    from loam_velodyne_b200 import synth
    return synth.make_scene()


@pytest.fixture(scope="session")
def oracle(build_libs):
    """The restatement (always buildable)."""
    from oracle import pydriver
    return pydriver.load("restatement")


@pytest.fixture(scope="session")
def reference(build_libs):
    """The compiled unmodified reference, when oracle/_ref travelled here."""
    from oracle import pydriver
    if not pydriver.available("reference"):
        pytest.skip("oracle/_ref/libloam_ref.so not present (needs /root/reference to build)")
    return pydriver.load("reference")


@pytest.fixture(scope="session")
def checker(build_libs):
    """Strongest checker available: compiled reference, else the restatement."""
    from oracle import pydriver
    return pydriver.best()


@pytest.fixture(scope="session")
def sweeps_vlp16(scene):
    from loam_velodyne_b200 import synth
    lidar = synth.Lidar.vlp16()
    return [synth.make_sweep(scene, lidar, i, yaw_rate=math.radians(5.0)) for i in range(8)]


@pytest.fixture(scope="session")
def map_200k(scene):
    from loam_velodyne_b200 import synth
    return synth.make_map(scene, 200_000)


@pytest.fixture(scope="session")
def ctx(build_libs):
    from loam_velodyne_b200 import api
    c = api.Ctx(0)
    yield c
    c.close()
