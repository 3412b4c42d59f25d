"""(This container only) run the UNMODIFIED reference class from /root/reference (TEST INFRASTRUCTURE).

The reference's pycnn/pycnn.py is executed *in place* (its source is read from
/root/reference at call time, never copied) inside a module whose ``__file__`` points into
oracle/_ref/pycnn/, so that (a) its ``lib/optimized.so`` lookup (pycnn/pycnn.py:18-22) and
(b) its ``from pycnn.modules.X import X`` imports (pycnn/pycnn.py:140-146) resolve to the
binaries oracle/build_ref.py compiled from the reference's own sources.  ``matplotlib`` is not
installed in this image and is imported unconditionally at pycnn/pycnn.py:1, so an empty stub
is injected (all uses are behind visualize=True).

Used by tests/golden/make_golden.py to emit golden vectors and by tests/test_oracle.py
(when /root/reference exists) to pin oracle/pycnn_oracle.py against the real thing.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

from . import build_ref

_cached = None


def available() -> bool:
    return build_ref.ref_available()


def load():
    """Returns the reference module (attributes: PyCNN, Evaluate, softmax, cross_entropy)."""
    global _cached
    if _cached is not None:
        return _cached
    if not available():
        raise RuntimeError("reference sources not present (expected only in the build container)")
    if not build_ref.build(verbose=False):
        raise RuntimeError("could not build oracle/_ref")
    if "matplotlib" not in sys.modules:
        try:
            importlib.import_module("matplotlib.pyplot")
        except Exception:
            mpl = types.ModuleType("matplotlib")
            plt = types.ModuleType("matplotlib.pyplot")
            mpl.pyplot = plt
            sys.modules["matplotlib"] = mpl
            sys.modules["matplotlib.pyplot"] = plt
    ref_pkg_dir = os.path.join(build_ref.OUT, "pycnn")
    if "pycnn" in sys.modules and not getattr(sys.modules["pycnn"], "__file__", "").startswith(build_ref.OUT):
        raise RuntimeError("another 'pycnn' package is already imported")
    if build_ref.OUT not in sys.path:
        sys.path.insert(0, build_ref.OUT)
    importlib.import_module("pycnn")          # oracle/_ref/pycnn (package marker + compiled kernels)
    src_path = os.path.join(build_ref.REF_ROOT, "pycnn", "pycnn.py")
    with open(src_path, "r") as f:
        source = f.read()
    mod = types.ModuleType("pycnn.pycnn")
    mod.__file__ = os.path.join(ref_pkg_dir, "pycnn.py")   # so base_dir/lib/optimized.so resolves
    mod.__package__ = "pycnn"
    sys.modules["pycnn.pycnn"] = mod
    exec(compile(source, src_path, "exec"), mod.__dict__)
    _cached = mod
    return mod
