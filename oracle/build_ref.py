"""Recipe: compile the reference's OWN native hot-path kernels into oracle/_ref/ (TEST INFRASTRUCTURE).

Sources are read where they lie under /root/reference (never copied into the repo):
  pycnn/lib/optimized.cpp                 -> oracle/_ref/pycnn/lib/optimized.so     (softmax, cross_entropy)
  pycnn/modules/{forward_pass,backward_pass,gradient_decent,max_pooling}.pyx
                                          -> oracle/_ref/pycnn/modules/<name>.so    (Cython -> C -> gcc)
Flags mirror the reference's portable (CI) build: pycnn/lib/Makefile:16-17 and setup.py:200-203
without -march=native/-ffast-math so the binaries travel to the GPU box's host CPU.
  tests/test.py                           -> oracle/_ref/tests/test_py.bin            (the reference's own unittest file,
                                             byte-compiled UNMODIFIED; tests/test_gpu_reference_own_suite.py runs it
                                             against the drop-in on the GPU box, where /root/reference does not exist)
Generated C goes to a temp dir; only compiled artefacts (.so, .pyc) and empty package markers land in oracle/_ref/
(git-ignored, NOT gpurun-ignored, so it ships with the snapshot).

The reference's setup.py is NOT run (it downloads wheels at import time, setup.py:94-95,190).
Usage: python -m oracle.build_ref   (no-op with a message when /root/reference is absent)
"""
from __future__ import annotations

import os
import subprocess
import sys
import sysconfig
import tempfile

REF_ROOT = os.environ.get("PYCNN_REFERENCE_ROOT", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
PYX = ["forward_pass", "backward_pass", "gradient_decent", "max_pooling"]


def ref_available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "pycnn", "lib", "optimized.cpp"))


def built() -> bool:
    if not os.path.isfile(os.path.join(OUT, "pycnn", "lib", "optimized.so")):
        return False
    return all(os.path.isfile(os.path.join(OUT, "pycnn", "modules", f"{n}.so")) for n in PYX)


def _newer(src: str, dst: str) -> bool:
    return (not os.path.exists(dst)) or os.path.getmtime(src) > os.path.getmtime(dst)


def build(force: bool = False, verbose: bool = True) -> bool:
    """Returns True if oracle/_ref is usable afterwards."""
    if not ref_available():
        if verbose:
            print(f"[oracle.build_ref] {REF_ROOT} not present; using prebuilt oracle/_ref: {built()}")
        return built()
    import numpy as np
    for sub in ("pycnn", "pycnn/lib", "pycnn/modules"):
        os.makedirs(os.path.join(OUT, sub), exist_ok=True)
        marker = os.path.join(OUT, sub, "__init__.py")
        if not os.path.exists(marker):
            with open(marker, "w") as f:
                f.write("# generated package marker (oracle/build_ref.py); compiled reference kernels live here\n")
    cflags = ["-O3", "-fPIC", "-funroll-loops"]
    # 1. optimized.cpp (reference Makefile uses `gcc` on the .cpp with -shared; we do the same)
    src = os.path.join(REF_ROOT, "pycnn", "lib", "optimized.cpp")
    dst = os.path.join(OUT, "pycnn", "lib", "optimized.so")
    if force or _newer(src, dst):
        subprocess.check_call(["gcc", *cflags, "-shared", "-o", dst, src, "-lm"])
    # 2. the four Cython modules
    inc = [f"-I{sysconfig.get_paths()['include']}", f"-I{np.get_include()}"]
    with tempfile.TemporaryDirectory(prefix="pycnn_ref_build_") as tmp:
        for name in PYX:
            src = os.path.join(REF_ROOT, "pycnn", "modules", f"{name}.pyx")
            dst = os.path.join(OUT, "pycnn", "modules", f"{name}.so")
            if not (force or _newer(src, dst)):
                continue
            c_file = os.path.join(tmp, f"{name}.c")
            subprocess.check_call([sys.executable, "-m", "cython", "-3", src, "-o", c_file])
            subprocess.check_call(["gcc", *cflags, "-shared", "-w",
                                   "-DNPY_NO_DEPRECATED_API=NPY_1_7_API_VERSION",
                                   *inc, "-o", dst, c_file])
    # 3. the reference's own test file, byte-compiled (no source text enters the repo)
    import py_compile
    src = os.path.join(REF_ROOT, "tests", "test.py")
    dst = os.path.join(OUT, "tests", "test_py.bin")     # byte code; not named .pyc: snapshot tools drop *.pyc
    if os.path.isfile(src) and (force or _newer(src, dst)):
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        py_compile.compile(src, cfile=dst, doraise=True)
    if verbose:
        print(f"[oracle.build_ref] built reference kernels into {OUT}")
    return built()


if __name__ == "__main__":
    ok = build(force="--force" in sys.argv)
    sys.exit(0 if ok else 1)
