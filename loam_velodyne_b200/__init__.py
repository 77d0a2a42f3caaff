"""loam_velodyne_b200 -- B200-native implementation of LOAM's per-sweep registration hot path
(feature extraction, scan-to-scan and scan-to-map Gauss-Newton) behind the reference's Basic* class API.

The package holds only what the path needs: ``csrc/`` (sm_100a kernels, the C ABI, the host-side C++ drop-in
classes), ``api`` (ctypes binding) and ``synth`` (the synthetic world used by tests and bench.py).
"""
from . import api, synth  # noqa: F401

__all__ = ["api", "synth"]
