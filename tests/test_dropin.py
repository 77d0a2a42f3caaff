"""The drop-in boundary, checked the way a maintainer would use it: the reference's five ROS adapters
(src/lib/{ScanRegistration,MultiScanRegistration,LaserOdometry,LaserMapping,TransformMaintenance}.cpp) are compiled
UNCHANGED against the product's replacement headers and linked against libloam_b200.so.

The adapters include their siblings with quote includes (LaserOdometry.h: #include "BasicLaserOdometry.h"), which resolve
in the including file's directory first, so the product headers must REPLACE the reference's files: the
test builds exactly that overlay in a temporary directory -- the reference's include/loam_velodyne/ with the product's
loam_velodyne/*.h copied over it.  ROS itself is absent from this image; oracle/shim/{ros,tf,nav_msgs,sensor_msgs,
geometry_msgs,pcl_conversions} stands in for the few names the adapters mention (test infrastructure).
Needs /root/reference (present in the build container, not on the GPU box)."""
import glob
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
HOST = os.path.join(ROOT, "loam_velodyne_b200", "csrc", "host")
ADAPTERS = ["ScanRegistration", "MultiScanRegistration", "LaserOdometry", "LaserMapping", "TransformMaintenance"]

pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "src", "lib")), reason="needs /root/reference")


@pytest.fixture(scope="module")
def overlay(tmp_path_factory):
    d = tmp_path_factory.mktemp("dropin")
    inc = d / "loam_velodyne"
    inc.mkdir()
    for f in glob.glob(os.path.join(REF, "include", "loam_velodyne", "*")):
        shutil.copy(f, inc)
    replaced = []
    for f in glob.glob(os.path.join(HOST, "loam_velodyne", "*.h")):
        shutil.copy(f, inc)  # the product's headers replace their upstream namesakes
        replaced.append(os.path.basename(f))
    assert {"BasicScanRegistration.h", "BasicLaserOdometry.h", "BasicLaserMapping.h", "BasicTransformMaintenance.h"} <= set(replaced)
    return d


def _flags(overlay):
    return ["-std=c++14", "-fPIC", f"-I{overlay}", f"-I{HOST}", f"-I{HOST}/compat", f"-I{ROOT}/include",
            f"-I{ROOT}/oracle/shim", f"-I{REF}/src/lib"]


@pytest.mark.parametrize("name", ADAPTERS)
def test_reference_adapter_compiles_against_product_headers(overlay, name):
    r = subprocess.run(["g++", "-fsyntax-only", *_flags(overlay), os.path.join(REF, "src", "lib", name + ".cpp")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]


def test_reference_adapters_link_against_the_product_library(overlay, build_libs):
    """Every Basic* symbol the adapters use is exported by libloam_b200.so (no undefined reference)."""
    out = overlay / "libadapters.so"
    srcs = [os.path.join(REF, "src", "lib", n + ".cpp") for n in ADAPTERS]
    r = subprocess.run(["g++", "-shared", "-o", str(out), *_flags(overlay), *srcs, "-Wl,--no-undefined",
                        f"-L{ROOT}/loam_velodyne_b200", "-lloam_b200"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]


def test_product_headers_do_not_redeclare_adapter_classes():
    """loam::MultiScanMapper belongs to upstream's MultiScanRegistration.h:49; a second definition in the product headers
    broke upstream's MultiScanRegistration.cpp in round 1."""
    for f in glob.glob(os.path.join(HOST, "loam_velodyne", "*.h")):
        assert "class MultiScanMapper" not in open(f).read(), f
