"""CPU-only: the C-ABI library loads, exports every symbol include/*.h declares, and fails loudly without a GPU."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(loam_b200_\w+)\s*\(", text)))


@pytest.mark.parametrize("header", ["loam_b200.h", "loam_b200_host.h"])
def test_every_declared_symbol_is_exported(build_libs, header):
    from loam_velodyne_b200 import api
    L = ctypes.CDLL(api.LIB_PATH)
    names = _declared(header)
    assert len(names) > 15
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, f"{header}: not exported: {missing}"


def test_python_binding_covers_the_headers(build_libs):
    from loam_velodyne_b200 import api
    L = api.lib()
    bound = set(L._signatures)
    declared = set(_declared("loam_b200.h")) | set(_declared("loam_b200_host.h"))
    assert declared <= bound, sorted(declared - bound)


def test_no_cpu_fallback(build_libs):
    """Without a usable GPU the product refuses to run instead of computing on the CPU."""
    from conftest import HAS_GPU
    from loam_velodyne_b200 import api
    if HAS_GPU:
        pytest.skip("a GPU is present; the loud-failure path is exercised on CPU-only boxes")
    with pytest.raises(api.LoamB200Error, match="no usable CUDA device"):
        api.Ctx(0)
    import numpy as np
    p = api.Pipeline()
    with pytest.raises(api.LoamB200Error, match="no usable CUDA device"):
        p.sweep(np.zeros((64, 4), np.float32), [32, 32])
    assert api.lib().loam_b200_strerror(-3).decode().startswith("no usable CUDA device")


def test_product_does_not_reference_the_oracle():
    """The shipped package must not import, link or execute anything under oracle/."""
    pkg = os.path.join(ROOT, "loam_velodyne_b200")
    bad = []
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", "Makefile")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                if re.search(r"oracle/|from oracle|import oracle|liboracle|libloam_ref|loamdrv_", txt):
                    bad.append(os.path.relpath(os.path.join(dp, f), ROOT))
    assert not bad, bad


def test_ring_ranges_match_reference_convention():
    from loam_velodyne_b200 import api
    import numpy as np
    s, e = api.ring_ranges([0, 5, 0, 3])
    # BasicScanRegistration.cpp:38-41: (cloudSize_before, cloudSize_after - 1) with 0 when the cloud is still empty
    np.testing.assert_array_equal(s, [0, 0, 5, 5])
    np.testing.assert_array_equal(e, [0, 4, 4, 7])
