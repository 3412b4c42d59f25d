"""CPU tier: the C-ABI library loads and exports every declared symbol, the host logic (model.bin
layout, batch sharding, world_size-2 gloo all-reduce of shard gradients) is right, and the product
refuses to run without a GPU instead of falling back."""
import contextlib
import io
import os
import re
import pickle
import socket
import subprocess
import sys
import textwrap

import numpy as np
import pytest

from oracle import pycnn_oracle as O
from pycnn_b200 import PyCNN, _lib, parallel
from tests.helpers import load_golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_header_symbol():
    lib = _lib.load()
    names = _lib.header_symbols()
    assert len(names) >= 50
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/pycnn_b200.h but not exported"
    assert set(names) == set(_lib._SIGNATURES), "ctypes binding and header disagree"
    assert lib.pycnn_b200_abi_version() == _lib.ABI_VERSION == int(re.search(r"PYCNN_B200_ABI_VERSION (\d+)", open(_lib.HEADER_PATH).read()).group(1))
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    assert set(names) <= exported


def test_library_is_sm100a_tcgen05_tma():
    """The shipped binary carries Blackwell tensor-core / TMA SASS (UTC*MMA, UTMALDG, LDTM)."""
    cuobjdump = "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in sass or "SM100" in sass.upper()
    # tcgen05.mma / TMA tensor loads / tcgen05.ld, plus what the cluster paths compile to: multicast TMA loads and
    # multicast tcgen05.commit (paired CTAs), cluster barriers (DSMEM split-K), TMA 1-D bulk copies (extractor)
    for mnemonic in ("UTCHMMA", "UTMALDG", "LDTM", "UTMALDG.2D.MULTICAST", "UTCBAR.MULTICAST", "UCGABAR_ARV", "UBLKCP"):
        assert mnemonic in sass, mnemonic
    # round 2: CTA-pair (cta_group::2) MMAs and their loads / commits, TMA stores of the fused tail, and the NVSwitch
    # multicast reduction load (multimem.ld_reduce) of the gradient exchange
    for mnemonic in ("UTCHMMA.2CTA", "UTMALDG.2D.2CTA", "UTCBAR.2CTA.MULTICAST", "UTMASTG.2D", "LDGMC.E.ADD.F32x4"):
        assert mnemonic in sass, mnemonic


def _gpu_present():
    try:
        return _lib.device_count() > 0
    except _lib.BackendError:
        return False


def test_no_cpu_fallback():
    m = PyCNN()
    with contextlib.redirect_stdout(io.StringIO()):
        m.init(layers=[8])
    assert m.use_cuda is False and m.optimizer == "sgd"
    with pytest.raises(RuntimeError, match="cuda\\(True\\)"):
        m._softmax_batch(np.ones((2, 3)))
    with pytest.raises(RuntimeError):
        m._preprocess_image(np.zeros((8, 8, 3), np.uint8), 8, m.filters)
    if not _gpu_present():
        with pytest.raises(_lib.BackendError, match="no CPU fallback"):
            m.cuda(True)
        assert m.use_cuda is False


def test_api_surface_matches_reference_defaults():
    m = PyCNN()
    assert m.filters == O.DEFAULT_FILTERS and len(m.filters) == 19
    with contextlib.redirect_stdout(io.StringIO()) as out:
        m.init(batch_size=16, layers=[64, 32], learning_rate=0.01, epochs=5)
        m.adam(beta1=0.8, beta2=0.99, eps=1e-7)
    assert (m.batch_size, m.layers, m.learning_rate, m.epochs) == (16, [64, 32], 0.01, 5)
    assert (m.optimizer, m.beta1, m.beta2, m.eps, m.t) == ("adam", 0.8, 0.99, 1e-7, 0)
    assert "Network initialized with 2 hidden layers: [64, 32]" in out.getvalue()
    assert "Adam optimizer enabled." in out.getvalue()
    custom = [[[1, 0, -1]] * 3, [[0, 1, 0], [1, -4, 1], [0, 1, 0]]]
    with contextlib.redirect_stdout(io.StringIO()):
        m.init(layers=[32], filters=custom)
    assert m.filters == custom
    for name in ("cuda", "adam", "init", "train_model", "predict", "save_model", "load_model",
                 "_preprocess_image", "_forward_pass", "_backward_pass", "_gradient_decent",
                 "_softmax_batch", "_cross_entropy_batch", "_max_pooling"):
        assert callable(getattr(m, name))
    assert callable(m.dataset.local) and callable(m.dataset.hf)


def test_init_layers_uses_reference_rng_stream():
    g = load_golden("mlp_small.npz")
    m = PyCNN()
    with contextlib.redirect_stdout(io.StringIO()):
        m.init(layers=g["layers"].tolist())
    m.dataset._init_layers(["0", "1", "2"], int(g["D"]))
    for k in m.weights:
        np.testing.assert_array_equal(m.weights[k], g[f"sgd_init_{k}"])
    for k in m.biases:
        np.testing.assert_array_equal(m.biases[k], g[f"sgd_init_{k}"])


def test_model_bin_layout_roundtrip(tmp_path, golden_dir):
    """model.bin written here has the reference's dict layout (pycnn/pycnn.py:693-701) and a model.bin
    written by the reference itself loads here."""
    m = PyCNN()
    with contextlib.redirect_stdout(io.StringIO()):
        m.init(layers=[6])
        m.load_model(os.path.join(golden_dir, "ref_model.bin"))
    assert m.classes == ["cat", "dog"] and m.layers == [6] and m.image_size == 8
    assert m.weights["w1"].shape == (6, 18) and m.weights["w1"].dtype == np.float64
    path = str(tmp_path / "model.bin")
    with contextlib.redirect_stdout(io.StringIO()):
        m.save_model(path)
    raw = open(path, "rb").read()
    assert raw[:2] == b"\x80\x04"
    ours = pickle.loads(raw)
    ref = pickle.load(open(os.path.join(golden_dir, "ref_model.bin"), "rb"))
    assert list(ours) == list(ref)
    for k in ref["weights"]:
        np.testing.assert_array_equal(ours["weights"][k], ref["weights"][k])
        assert ours["weights"][k].dtype == np.float64
    for k in ref["biases"]:
        np.testing.assert_array_equal(ours["biases"][k], ref["biases"][k])
    assert ours["filters"] == ref["filters"] and ours["classes"] == ref["classes"]
    assert ours["layers"] == ref["layers"] and ours["image_size"] == ref["image_size"]


def test_batch_slices_partition_every_batch():
    for n, bs, world in [(20, 8, 2), (4096, 4096, 8), (37, 5, 4), (3, 32, 8), (65536, 4096, 3)]:
        for start in range(0, n, bs):
            end = min(start + bs, n)
            got = []
            for r in range(world):
                lo, hi = parallel.batch_slice(start, end, r, world)
                assert start <= lo <= hi <= end
                got.extend(range(lo, hi))
            assert got == list(range(start, end))
    assert [parallel.image_shard(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 9), (9, 10)]


WORKER = textwrap.dedent('''
    import os, sys
    sys.path.insert(0, os.environ["PYCNN_ROOT"])
    import numpy as np, torch, torch.distributed as dist
    from oracle import pycnn_oracle as O
    from pycnn_b200 import parallel
    from tests.helpers import load_golden
    rank, world, _ = parallel.init_process_group(backend="gloo")
    assert world == 2
    # the 128-byte id travels from rank 0 exactly as the NCCL unique id does on the GPU box
    uid = parallel.broadcast_bytes(bytes(range(128)) if rank == 0 else None)
    assert uid == bytes(range(128))
    g = load_golden("mlp_small.npz")
    layers, D, C, bs = g["layers"].tolist(), int(g["D"]), int(g["C"]), int(g["bs"])
    data, y = g["data"], O.one_hot(g["labels"], C)
    w, b = O.init_layers(D, layers, C, seed=42)
    keys = list(w) + list(b)
    def flat(gw, gb):
        return np.concatenate([gw[k].ravel() for k in w] + [gb[k].ravel() for k in b])
    def grad_fn(idx):
        if len(idx) == 0:
            return np.zeros(sum(v.size for v in list(w.values()) + list(b.values()))), 0.0, 0
        p, acts, _ = O.forward(data[idx], w, b, layers)
        gw, gb = O.backward(p, y[idx], acts, w, layers)
        return flat(gw, gb), float(O.cross_entropy(p, y[idx]).sum()), int((p.argmax(1) == y[idx].argmax(1)).sum())
    def allreduce(a):
        t = torch.from_numpy(np.array(a, dtype=np.float64))
        dist.all_reduce(t)                 # SUM: gradients are batch sums, no 1/world
        return t.numpy()
    def update_fn(gflat):
        gw, gb, o = {}, {}, 0
        for k in w:
            gw[k] = gflat[o:o + w[k].size].reshape(w[k].shape); o += w[k].size
        for k in b:
            gb[k] = gflat[o:o + b[k].size].reshape(b[k].shape); o += b[k].size
        O.update_cpu_rule(w, b, gw, gb, 0.05, "sgd", {}, {})
    stepper = parallel.ShardedStepper(rank, world, grad_fn, allreduce, update_fn)
    perms = [np.random.permutation(len(data)) for _ in range(3)]   # same stream on both ranks (seed 42)
    totals = [stepper.epoch(p, bs) for p in perms]
    for k in w:
        np.testing.assert_allclose(w[k], g[f"sgd_trained_{k}"], atol=1e-12)   # == the reference's 1-process run
    s = parallel.allreduce_scalars([1.0, float(rank)])
    assert s == [2.0, 1.0]
    if rank == 0:
        print("GLOO_OK", totals[-1])
    dist.destroy_process_group()
''')


def test_world2_gloo_sharded_training_equals_single_process(tmp_path):
    """Two CPU ranks, each computing its slice of every batch with the oracle, all-reducing SUMs over
    gloo and applying the same update: the weights equal the reference's single-process train_model."""
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    env = dict(os.environ, PYCNN_ROOT=ROOT, OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)]
    res = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=240)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    assert "GLOO_OK" in res.stdout


_SHARD_WORKER = r"""
import os, sys
import numpy as np
sys.path.insert(0, {root!r})
import torch.distributed as dist
from pycnn_b200 import parallel

class FakeCtx:
    # stands in for the CUDA context: "predicts" a deterministic function of the pixels so that any mix-up of
    # shard order or a dropped / duplicated image shows in the gathered result
    def predict_images(self, images):
        s = images.reshape(len(images), -1).astype(np.int64).sum(1)
        return (s % 7).astype(np.int32), (s % 100 / 100.0).astype(np.float32)
    def evaluate_images(self, images, labels, bs):
        cls, _ = self.predict_images(images)
        return float((cls + 1).sum()), int((cls == labels).sum()), cls

rank, world, _ = parallel.init_process_group("gloo")
rng = np.random.default_rng(5)
for n in (0 + 5, 64, 1001):                   # fewer images than ranks x 4, even, ragged
    images = rng.integers(0, 256, (n, 4, 4, 3), dtype=np.uint8)
    labels = rng.integers(0, 7, n).astype(np.int32)
    want_cls, want_conf = FakeCtx().predict_images(images)
    cls, conf = parallel.predict_sharded(FakeCtx(), images, rank, world)
    assert np.array_equal(cls, want_cls) and np.array_equal(conf, want_conf), n
    loss, correct, pred = parallel.evaluate_sharded(FakeCtx(), images, labels, 32, rank, world)
    assert np.array_equal(pred, want_cls) and correct == int((want_cls == labels).sum()) and loss == float((want_cls + 1).sum()), n
    lo, hi = parallel.image_shard(n, rank, world)
    spans = [None] * world
    dist.all_gather_object(spans, (lo, hi))
    assert spans[0][0] == 0 and spans[-1][1] == n and all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
print("SHARD_OK", rank)
dist.destroy_process_group()
"""


def test_world2_gloo_image_sharded_predict_and_evaluate(tmp_path):
    """Extraction / predict / Evaluate shard per image with no collective in the data path: the host logic
    (contiguous N/W ranges, host-side gather and tally sums) on two gloo ranks with a stand-in context."""
    script = tmp_path / "shard_worker.py"
    script.write_text(_SHARD_WORKER.format(root=ROOT))
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(script)]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-3000:]
    assert res.stdout.count("SHARD_OK") == 2


_EXCHANGE_WORKER = r"""
import os, sys, tempfile
sys.path.insert(0, {root!r})
import torch.distributed as dist
from pycnn_b200 import parallel

class FakeCtx:
    # stands in for the CUDA context: the multicast "object" is a temp file whose descriptor travels to the other rank
    def __init__(self, rank, supported=True, floats=1000, bind_fails=False):
        self.rank, self.supported, self.floats, self.bind_fails = rank, supported, floats, bind_fails
        self.log, self.inode = [], None
    def grad_buffer(self):
        return 0, self.floats
    def comm_nvls_supported(self):
        return self.supported
    def comm_nvls_create(self, world):
        self.tmp = tempfile.NamedTemporaryFile(prefix="pycnn_b200_fake_mc_")
        self.log.append("create")
        return os.dup(self.tmp.fileno())
    def comm_nvls_join(self, fd, rank, world):
        self.log.append("join")
        self.inode = os.fstat(fd if fd >= 0 else self.tmp.fileno()).st_ino
        if fd >= 0:
            os.fstat(fd)                                   # a live descriptor in THIS process
    def comm_nvls_bind(self):
        if self.bind_fails:
            raise RuntimeError("bind failed on this rank")
        self.log.append("bind")
    def comm_nvls_detach(self):
        self.log.append("detach")
    def comm_ipc_export(self):
        self.log.append("ipc_export")
        return bytes([self.rank]) * 192
    def comm_ipc_attach(self, blobs, rank, world):
        assert len(blobs) == 192 * world and blobs[0] == 0 and blobs[192] == 1
        self.log.append("ipc_attach")

rank, world, _ = parallel.init_process_group("gloo")
os.environ.pop("PYCNN_B200_DP", None)
# 1. both ranks offer multicast: rank 0's descriptor arrives on rank 1 and names the same object
c = FakeCtx(rank)
assert parallel.attach_exchange(c, rank, world) == "nvls"
assert c.log == (["create", "join", "bind"] if rank == 0 else ["join", "bind"]), c.log
inodes = [None] * world
dist.all_gather_object(inodes, c.inode)
assert inodes[0] == inodes[1] and inodes[0] is not None
# 2. one rank without multicast support: every rank falls back to the peer-memory exchange
c = FakeCtx(rank, supported=(rank == 0))
assert parallel.attach_exchange(c, rank, world) == "p2p" and c.log == ["ipc_export", "ipc_attach"], c.log
# 3. bind fails on one rank: everybody detaches and falls back together
c = FakeCtx(rank, bind_fails=(rank == 1))
assert parallel.attach_exchange(c, rank, world) == "p2p"
assert "detach" in c.log and c.log[-2:] == ["ipc_export", "ipc_attach"], c.log
# 4. a large network takes the sharded peer-memory update, an explicit mode wins, nccl attaches nothing
c = FakeCtx(rank, floats=parallel.NVLS_MAX_FLOATS + 1)
assert parallel.attach_exchange(c, rank, world) == "p2p" and c.log == ["ipc_export", "ipc_attach"]
c = FakeCtx(rank, floats=parallel.NVLS_MAX_FLOATS + 1)
assert parallel.attach_exchange(c, rank, world, "nvls") == "nvls"
c = FakeCtx(rank)
assert parallel.attach_exchange(c, rank, world, "nccl") == "nccl" and c.log == []
print("EXCHANGE_OK", rank)
dist.destroy_process_group()
"""


def test_world2_gloo_exchange_selection_and_descriptor_handoff(tmp_path):
    """parallel.attach_exchange on two gloo ranks with a stand-in context: the multicast object's file descriptor
    reaches the other rank (SCM_RIGHTS over a unix socket) and names the same object, every rank takes the same
    fallback when one of them cannot attach, and the size rule / explicit modes are honoured."""
    script = tmp_path / "exchange_worker.py"
    script.write_text(_EXCHANGE_WORKER.format(root=ROOT))
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(script)]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-3000:]
    assert res.stdout.count("EXCHANGE_OK") == 2


def test_checkpoint_sidecar_does_not_shadow_model_bin(tmp_path):
    """the resumable checkpoint is a separate, tagged file; a reference-format model.bin is refused by
    load_checkpoint before any device work (so this runs without a GPU) and the tag is what save_checkpoint writes."""
    import pickle
    from pycnn_b200 import PyCNN
    m = PyCNN()
    model_bin = tmp_path / "model.bin"
    with open(model_bin, "wb") as f:
        pickle.dump({"weights": {}, "biases": {}, "filters": m.filters, "image_size": 32, "classes": ["a"],
                     "layers": [8], "trained_with_cuda": False}, f, protocol=4)
    with pytest.raises(ValueError, match="checkpoint"):
        m.load_checkpoint(str(model_bin))
    assert PyCNN.CHECKPOINT_FORMAT.startswith("pycnn_b200.checkpoint/")
    import inspect
    sig = inspect.signature(PyCNN.train_model)
    assert list(sig.parameters)[:3] == ["self", "visualize", "early_stop"]          # reference signature first (:572)
    assert sig.parameters["start_epoch"].default == 0 and sig.parameters["checkpoint"].default is None


def test_threaded_decode_keeps_sequential_semantics(tmp_path):
    """dataset ingest decodes on host threads but keeps the reference loop's meaning (pycnn.py:355-372):
    listing order, failures reported and skipped, max_image counts successes only."""
    from PIL import Image
    from pycnn_b200.pycnn import _decode_files, _to_u8_image
    rng = np.random.default_rng(3)
    paths = []
    for i in range(9):
        pth = tmp_path / f"img_{i}.png"
        if i in (1, 4):
            pth.write_bytes(b"not an image")                    # undecodable files in the middle of the listing
        else:
            Image.fromarray(rng.integers(0, 256, (20 + i, 20 + i, 3), dtype=np.uint8), mode="RGB").save(pth)
        paths.append(str(pth))
    ok, bad = [], []
    got = _decode_files(paths, 16, limit=5, on_ok=lambda p, n: ok.append((os.path.basename(p), n)),
                        on_fail=lambda p, e: bad.append(os.path.basename(p)), workers=4)
    assert [o[0] for o in ok] == ["img_0.png", "img_2.png", "img_3.png", "img_5.png", "img_6.png"]
    assert [o[1] for o in ok] == [1, 2, 3, 4, 5] and bad == ["img_1.png", "img_4.png"]
    assert len(got) == 5 and all(g.shape == (16, 16, 3) and g.dtype == np.uint8 for g in got)
    for g, name in zip(got, [o[0] for o in ok]):
        np.testing.assert_array_equal(g, _to_u8_image(str(tmp_path / name), 16))    # same pixels as the sequential path
    assert len(_decode_files(paths, 16)) == 7 and _decode_files(paths, 16, limit=0) == []


def test_oracle_is_test_infrastructure_only():
    """The oracle is the checker, never the product: nothing under pycnn_b200/ or scripts/ imports it, bench.py only
    in its cpu_baseline / --impl reference legs, __graft_entry__ only in build() (builds the checker) and smoke()."""
    import ast
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

    def oracle_imports(path):
        tree = ast.parse(open(path).read())
        hits = []
        for node in ast.walk(tree):
            if isinstance(node, ast.Import) and any(a.name.split(".")[0] == "oracle" for a in node.names):
                hits.append(node.lineno)
            if isinstance(node, ast.ImportFrom) and (node.module or "").split(".")[0] == "oracle":
                hits.append(node.lineno)
        return hits

    for sub in ("pycnn_b200", "scripts"):
        for dirpath, _, files in os.walk(os.path.join(root, sub)):
            for f in files:
                if f.endswith(".py"):
                    assert not oracle_imports(os.path.join(dirpath, f)), f"{sub}/{f} imports the oracle"
    # bench.py: module level must be oracle-free; the imports live inside the CPU-baseline helpers
    tree = ast.parse(open(os.path.join(root, "bench.py")).read())
    top = [n for n in tree.body if isinstance(n, (ast.Import, ast.ImportFrom))]
    assert not any(isinstance(n, ast.ImportFrom) and (n.module or "").startswith("oracle") for n in top)
    assert not any(isinstance(n, ast.Import) and any(a.name.startswith("oracle") for a in n.names) for n in top)
    fn_with_oracle = {n.name for n in tree.body if isinstance(n, ast.FunctionDef)
                      for m in ast.walk(n) if isinstance(m, ast.ImportFrom) and (m.module or "").startswith("oracle")}
    assert fn_with_oracle and all("cpu" in name or "reference" in name for name in fn_with_oracle), fn_with_oracle


def test_committed_bench_line_honours_the_contract():
    """The JSON line bench.py printed on the B200 (profiles/r2_bench_c1_1gpu.json) carries every key the driver reads."""
    import json
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    line = open(os.path.join(root, "profiles", "r2_bench_c1_1gpu.json")).read().strip().splitlines()[-1]
    d = json.loads(line)
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["metric"] == "train_images_per_sec" and d["unit"] == "images/s" and d["higher_is_better"] is True
    assert d["n_gpus"] == 1 and d["scaling"] == "weak" and d["data"] == "synthetic" and d["dtype"] == "tf32"
    assert d["warmup"] >= 3 and d["gpu_launches"] > 0 and "workload" in d["config"] and "l2" in d["config"]
    assert abs(d["value"] - d["steps"] * d["config"]["global_batch"] / (d["ms_per_step"] * d["steps"] / 1e3)) < 1e-6 * d["value"]
    e = d["e2e"]
    assert e["unit"] == d["unit"] and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["value"] < d["value"]
    r = d["roofline"]
    assert r["bound"] in ("hbm", "tensor") and r["unit"] in ("GB/s", "TFLOP/s") and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert r["traffic"] is None or r["traffic"] > 0
    c = d["cpu_baseline"]
    assert c["kind"] in ("reference", "port") and c["cores"] >= 1 and c["unit"] == d["unit"] and c["sample"]
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}


def test_committed_multi_gpu_lines_and_reference_arm():
    """profiles/r2_bench_c1_{2,4,8}gpu.json: same metric / workload as the 1-GPU line, weak scaling, whole-job value,
    dp_parity green; profiles/r2_bench_reference_arm.json: the reference arm's line on the same config."""
    import json
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    load = lambda name: json.loads(open(os.path.join(root, "profiles", name)).read().strip().splitlines()[-1])
    one = load("r2_bench_c1_1gpu.json")
    for n in (2, 4, 8):
        d = load(f"r2_bench_c1_{n}gpu.json")
        assert d["n_gpus"] == n and d["metric"] == one["metric"] and d["scaling"] == "weak" and d["steps"] == one["steps"]
        assert d["config"]["workload"] == one["config"]["workload"] and d["config"]["global_batch"] == n * one["config"]["global_batch"]
        assert abs(d["value"] - d["config"]["global_batch"] / (d["ms_per_step"] / 1e3)) < 1e-6 * d["value"]
        p = d["dp_parity"]
        assert p["ok"] and p["ranks_identical_params"] and p["loss_rel_dev"] < 1e-3 and p["param_max_rel_err"] < 2e-2
        assert d["value"] > 0.75 * n * one["value"]                      # the measured weak-scaling efficiency floor
        assert d["e2e"]["value"] < d["value"] and d["gpu_launches"] > 0
    ref = load("r2_bench_reference_arm.json")
    assert ref["impl"] == "reference" and ref["metric"] == one["metric"] and ref["unit"] == one["unit"]
    assert ref["config"]["workload"] == one["config"]["workload"]
    assert ref["e2e"]["h2d_bytes_per_step"] == 0 and ref["e2e"]["d2h_bytes_per_step"] == 0 and ref["e2e"]["value"] == ref["value"]
    assert ref["cpu_baseline"]["kind"] in ("reference", "port") and ref["cpu_baseline"]["cores"] >= 1


def test_push_dataset_reuploads_edited_labels():
    """model.labels replaced or edited while model.data stays the device-resident feature matrix: the next train_model
    must see the new labels (host logic only; the context is a recorder)."""
    from pycnn_b200.pycnn import DeviceFeatures

    class Recorder:
        def __init__(self):
            self.uploads = []

        def dataset_set_labels(self, row0, idx):
            self.uploads.append((row0, np.array(idx)))

    m = PyCNN()
    rec = Recorder()
    m._bind_network = lambda: rec
    m.data = DeviceFeatures(rec, 6, 4)
    m._dataset_token = (id(m.data), 6)
    m.labels = np.eye(3, dtype=np.int64)[[0, 1, 2, 0, 1, 2]]
    m._labels_uploaded = None
    m._push_dataset()
    assert len(rec.uploads) == 1 and rec.uploads[0][1].tolist() == [0, 1, 2, 0, 1, 2]
    m._push_dataset()                                                    # unchanged: nothing travels
    assert len(rec.uploads) == 1
    m.labels[0] = [0, 0, 1]                                              # edited in place
    m._push_dataset()
    assert len(rec.uploads) == 2 and rec.uploads[1][1].tolist() == [2, 1, 2, 0, 1, 2]
    m.labels = np.eye(3, dtype=np.int64)[[1, 1, 1, 1, 1]]                # wrong length
    with pytest.raises(ValueError):
        m._push_dataset()
