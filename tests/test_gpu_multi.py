"""Multi-GPU parity (needs >= 2 B200s on the box; skipped otherwise): NCCL all-reduce path and the fused
peer-memory exchange both reproduce the single-GPU result."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpus():
    from pycnn_b200 import _lib
    try:
        return _lib.device_count()
    except Exception:
        return 0


@pytest.mark.skipif(_ngpus() < 2, reason="needs two GPUs")
def test_two_rank_training_matches_single_gpu():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "dp_worker.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-5000:]
    assert res.stdout.count("DP_OK") == 6, res.stdout[-3000:]


@pytest.mark.skipif(_ngpus() < 2, reason="needs two GPUs")
def test_two_rank_pycnn_surface_tf32(tmp_path):
    """PyCNN.train_model + parallel.attach (peer-memory exchange, TF32, configs[1] widths), image-sharded
    predict / Evaluate, and checkpoint resume with the sharded optimizer state -- all against single-GPU runs."""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "dp_surface_worker.py")]
    env = dict(os.environ, PYCNN_B200_TEST_TMP=str(tmp_path))
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT, env=env)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-5000:]
    assert res.stdout.count("DP_SURFACE_OK") == 3, res.stdout[-3000:]
