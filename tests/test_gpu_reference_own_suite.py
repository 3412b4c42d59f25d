"""The reference's OWN unittest file (tests/test.py:9-223, 16 cases) run UNMODIFIED against the drop-in.

oracle/build_ref.py byte-compiles /root/reference/tests/test.py into oracle/_ref/tests/test_py.bin (git-ignored, travels
to the GPU box; no reference source text enters the repo).  Here that module is loaded under a `pycnn` package that
resolves to pycnn_b200 with the B200 backend switched on (tests/ref_shim), and all 16 cases must pass: API shape,
defaults, dataset loading, the max-pool known answer, save/load round trip, training, prediction, evaluation."""
import importlib.machinery
import importlib.util
import io
import os
import sys
import unittest

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PYC = os.path.join(ROOT, "oracle", "_ref", "tests", "test_py.bin")


@pytest.mark.skipif(not os.path.isfile(PYC), reason="oracle/_ref/tests/test_py.bin not built (python -m oracle.build_ref)")
def test_reference_unittest_file_passes_against_the_drop_in(capsys):
    shim = os.path.join(ROOT, "tests", "ref_shim")
    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k == "pycnn" or k.startswith("pycnn.")}
    sys.path.insert(0, shim)
    try:
        loader = importlib.machinery.SourcelessFileLoader("reference_tests_test", PYC)
        spec = importlib.util.spec_from_loader("reference_tests_test", loader)
        mod = importlib.util.module_from_spec(spec)
        loader.exec_module(mod)
        import pycnn.pycnn as served
        assert served.__file__.startswith(shim), "the `pycnn` the reference tests import must be the drop-in shim"
        suite = unittest.TestLoader().loadTestsFromTestCase(mod.TestPyCNN)
        stream = io.StringIO()
        result = unittest.TextTestRunner(stream=stream, verbosity=2).run(suite)
    finally:
        sys.path.remove(shim)
        for k in [k for k in sys.modules if k == "pycnn" or k.startswith("pycnn.")]:
            del sys.modules[k]
        sys.modules.update(saved)
    with capsys.disabled():
        print(f"\n[reference tests/test.py, unmodified, via pycnn -> pycnn_b200 + cuda(True)] ran {result.testsRun}, "
              f"failures {len(result.failures)}, errors {len(result.errors)}")
    assert result.testsRun == 16, stream.getvalue()[-3000:]
    assert result.wasSuccessful(), stream.getvalue()[-6000:]
