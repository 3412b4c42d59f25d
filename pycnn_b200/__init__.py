"""pycnn_b200 -- B200-native backend for the PyCNN `cuda(True)` hot path.

    from pycnn_b200 import PyCNN, Evaluate      # same surface as `from pycnn import PyCNN, Evaluate`

Everything numeric runs in libpycnn_b200.so (hand-written sm_100a kernels behind a C ABI,
include/pycnn_b200.h); importing this package does not need a GPU, using it does.
"""
from .pycnn import DeviceFeatures, Evaluate, PyCNN  # noqa: F401

__all__ = ["PyCNN", "Evaluate", "DeviceFeatures", "PyCNNTorchModel"]
__version__ = "0.1.0"


def __getattr__(name):   # PyCNNTorchModel needs torch: resolved on first use only
    if name == "PyCNNTorchModel":
        from . import torch_export
        return torch_export.PyCNNTorchModel
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")
