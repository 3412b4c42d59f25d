"""pycnn_b200 -- B200-native backend for the PyCNN `cuda(True)` hot path.

    from pycnn_b200 import PyCNN, Evaluate      # same surface as `from pycnn import PyCNN, Evaluate`

Everything numeric runs in libpycnn_b200.so (hand-written sm_100a kernels behind a C ABI,
include/pycnn_b200.h); importing this package does not need a GPU, using it does.
"""
from .pycnn import DeviceFeatures, Evaluate, PyCNN  # noqa: F401

__all__ = ["PyCNN", "Evaluate", "DeviceFeatures"]
__version__ = "0.1.0"
