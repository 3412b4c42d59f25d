"""The reference's own conformance suite (/root/reference/tests/test.py, 16 cases) re-stated against
`pycnn_b200.PyCNN` with `cuda(True)`: same fixture (2 classes x 5 random 32x32 PNGs), same calls, same
assertions, plus numeric checks against the oracle where the reference only smoke-tests."""
import contextlib
import io
import os

import numpy as np
import pytest
from PIL import Image

from oracle import pycnn_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dataset_dir(tmp_path_factory):
    base = tmp_path_factory.mktemp("pycnn_ds")
    rng = np.random.default_rng(123)
    for cls in ("class_a", "class_b"):
        os.makedirs(base / cls)
        for i in range(5):
            if cls == "class_a":
                arr = rng.integers(150, 255, (32, 32, 3), dtype=np.uint8)
                arr[:, :, 1:] = rng.integers(0, 50, (32, 32, 2), dtype=np.uint8)
            else:
                arr = rng.integers(0, 50, (32, 32, 3), dtype=np.uint8)
                arr[:, :, 2] = rng.integers(150, 255, (32, 32), dtype=np.uint8)
            Image.fromarray(arr, mode="RGB").save(base / cls / f"img_{i}.png")
    return str(base)


def make(**kw):
    from pycnn_b200 import PyCNN
    m = PyCNN()
    with contextlib.redirect_stdout(io.StringIO()):
        m.init(**kw)
        m.set_math_mode("fp32")
        m.cuda(True)
    return m


def quiet(fn, *a, **k):
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        r = fn(*a, **k)
    return r, buf.getvalue()


def test_01_02_import_and_defaults():
    from pycnn_b200 import PyCNN
    m = PyCNN()
    assert m.optimizer == "sgd" and m.use_cuda is False
    with contextlib.redirect_stdout(io.StringIO()) as out:
        m.cuda(True)
    assert m.use_cuda is True and "CUDA backend initialized successfully" in out.getvalue()
    m.cuda(False)
    assert m.use_cuda is False


def test_05_load_local_dataset(dataset_dir):
    m = make(batch_size=4, layers=[32], learning_rate=0.01, epochs=2)
    _, out = quiet(m.dataset.local, dataset_dir, max_image=3, image_size=32)
    assert len(m.classes) == 2 and len(m.data) == 6 and len(m.data) == len(m.labels)
    assert "Loaded 6 images from 2 classes" in out
    feats = np.asarray(m.data)
    assert feats.shape == (6, 4275) and feats.dtype == np.float64
    # numerics the reference does not pin: the loaded feature rows equal the oracle's extractor on the same files
    classes = os.listdir(dataset_dir)
    first = [f for f in os.listdir(os.path.join(dataset_dir, classes[0])) if f.endswith(".png")][0]
    img = np.asarray(Image.open(os.path.join(dataset_dir, classes[0], first)).convert("RGB").resize((32, 32), Image.Resampling.LANCZOS))
    np.testing.assert_allclose(feats[0], O.extract_features(img), atol=2e-5)
    assert m.labels[0].tolist() == [1, 0]


def test_06_preprocessing():
    m = make(layers=[32])
    img = Image.new("RGB", (32, 32), color=(255, 0, 0))
    out = m._preprocess_image(img, 32, m.filters[:3])
    assert isinstance(out, np.ndarray) and out.ndim == 1 and out.shape[0] == 675
    np.testing.assert_allclose(out, O.extract_features(np.asarray(img), m.filters[:3]), atol=2e-5)


def test_07_11_12_small_dropins():
    m = make(layers=[8])
    x = np.array([[1.0, 2.0, 3.0], [4.0, 5.0, 6.0]])
    r = m._softmax_batch(x)
    assert r.shape == x.shape
    np.testing.assert_array_almost_equal(r.sum(axis=1), [1.0, 1.0], decimal=5)
    with pytest.raises(ValueError):
        m._softmax_batch(np.ones(3))
    preds = np.array([[0.7, 0.2, 0.1], [0.1, 0.8, 0.1]])
    labels = np.array([[1.0, 0.0, 0.0], [0.0, 1.0, 0.0]])
    loss = m._cross_entropy_batch(preds, labels)
    assert loss.shape[0] == 2 and np.all(loss >= 0)
    np.testing.assert_allclose(loss, O.cross_entropy(preds, labels), rtol=1e-12)
    fm = np.arange(1, 17, dtype=float).reshape(4, 4)
    pooled = m._max_pooling(fm)
    assert pooled.shape == (2, 2) and pooled.tolist() == [[6, 8], [14, 16]]


def test_08_training_matches_oracle(dataset_dir):
    m = make(batch_size=4, layers=[16], learning_rate=0.1, epochs=2)
    quiet(m.dataset.local, dataset_dir, max_image=3, image_size=32)
    w0 = {k: v.copy() for k, v in m.weights.items()}
    b0 = {k: v.copy() for k, v in m.biases.items()}
    feats, y = np.asarray(m.data), np.asarray(m.labels)
    state = np.random.get_state()
    _, out = quiet(m.train_model, visualize=False)
    assert "w1" in m.weights and "wo" in m.weights
    assert "Training model: epoch=2/2" in out and "correct=" in out and "loss=" in out
    np.random.set_state(state)
    O.train_epochs(feats, y, w0, b0, [16], 0.1, "sgd", 4, epochs=2)
    for k in w0:
        assert np.abs(m.weights[k] - w0[k]).max() / np.abs(w0[k]).max() < 1e-4, k


def test_09_save_and_load_model(dataset_dir, tmp_path):
    m = make(batch_size=4, layers=[16], learning_rate=0.1, epochs=1)
    quiet(m.dataset.local, dataset_dir, max_image=2, image_size=32)
    quiet(m.train_model, visualize=False)
    path = str(tmp_path / "test_model.bin")
    _, out = quiet(m.save_model, path)
    assert os.path.exists(path) and "Model saved in" in out
    m2 = make(layers=[16])
    quiet(m2.load_model, path)
    assert len(m2.classes) == len(m.classes) and m2.layers == m.layers
    for k in m.weights:
        np.testing.assert_array_almost_equal(m.weights[k], m2.weights[k], decimal=5)
    # and the reference-format file predicts the same thing
    img = os.path.join(dataset_dir, m.classes[0], "img_0.png")
    assert m.predict(img) == m2.predict(img)


def test_10_prediction(dataset_dir):
    m = make(batch_size=4, layers=[16], learning_rate=0.1, epochs=2)
    quiet(m.dataset.local, dataset_dir, max_image=3, image_size=32)
    quiet(m.train_model, visualize=False)
    img_path = os.path.join(dataset_dir, "class_a", "img_0.png")
    label, conf = m.predict(img_path)
    assert label in m.classes and 0.0 <= conf <= 1.0 and isinstance(conf, float)
    img = np.asarray(Image.open(img_path).convert("RGB").resize((32, 32), Image.Resampling.LANCZOS))
    idx, want_conf = O.predict_single(O.extract_features(img), m.weights, m.biases, m.layers)
    assert label == m.classes[idx] and abs(conf - want_conf) < 1e-4


def test_13_14_filters():
    from pycnn_b200 import PyCNN
    m = PyCNN()
    assert len(m.filters) == 19 and all(len(f) == 3 and len(f[0]) == 3 for f in m.filters)
    custom = [[[1, 0, -1], [1, 0, -1], [1, 0, -1]], [[0, 1, 0], [1, -4, 1], [0, 1, 0]]]
    with contextlib.redirect_stdout(io.StringIO()):
        m.init(layers=[32], filters=custom)
    assert m.filters == custom


def test_15_early_stopping(dataset_dir):
    m = make(batch_size=4, layers=[16], learning_rate=0.1, epochs=100)
    quiet(m.dataset.local, dataset_dir, max_image=3, image_size=32)
    _, out = quiet(m.train_model, visualize=False, early_stop=3)
    assert "w1" in m.weights
    assert "Early stopping at epoch" in out or "Model reached maximum learning" in out or "epoch=100/100" in out


def test_16_evaluation(dataset_dir):
    from pycnn_b200 import Evaluate
    m = make(batch_size=4, layers=[16], learning_rate=0.1, epochs=2)
    quiet(m.dataset.local, dataset_dir, max_image=3, image_size=32)
    quiet(m.train_model, visualize=False)
    ev = Evaluate(m)
    res, out = quiet(ev.local, dataset_dir, max_image=2)
    assert res is None and "Evaluation Results:" in out and "Acc:" in out and "class_a:" in out
    acc = ev.last_result["acc"]
    # the reference-signature path (feature rows + one-hot labels) gives the same tallies
    feats, y = np.asarray(m.data), np.asarray(m.labels)
    quiet(ev._run_inference, list(feats), list(y))
    assert 0.0 <= acc <= 1.0 and ev.last_result["per_class"]["class_a"]["total"] == 3


def test_adam_and_private_call_protocol(dataset_dir):
    """The reference's private loop (_forward_pass -> CE -> _backward_pass -> _gradient_decent) still works."""
    m = make(batch_size=4, layers=[16], learning_rate=0.01, epochs=1)
    with contextlib.redirect_stdout(io.StringIO()):
        m.adam()
    quiet(m.dataset.local, dataset_dir, max_image=3, image_size=32)
    feats, y = np.asarray(m.data), np.asarray(m.labels)
    w = {k: v.copy() for k, v in m.weights.items()}
    b = {k: v.copy() for k, v in m.biases.items()}
    gw = {k: np.zeros(v.shape) for k, v in m.weights.items()}
    gb = {k: np.zeros(v.shape) for k, v in m.biases.items()}
    acts, zs = [None] * 2, [None] * 2
    out = m._forward_pass(feats[:4], acts, zs)
    loss = m._cross_entropy_batch(out, y[:4]).sum()
    m._backward_pass(out, y[:4], acts, zs, gw, gb)
    m._gradient_decent(gw, gb)
    p, a, _ = O.forward(feats[:4], w, b, [16])
    assert abs(loss - O.cross_entropy(p, y[:4]).sum()) < 1e-4 * abs(loss)
    ogw, ogb = O.backward(p, y[:4], a, w, [16])
    for k in ogw:
        assert np.abs(gw[k] - ogw[k]).max() <= 1e-4 * max(np.abs(ogw[k]).max(), 1e-12), k
    mm, vv = {k: np.zeros_like(x) for k, x in {**w, **b}.items()}, {k: np.zeros_like(x) for k, x in {**w, **b}.items()}
    O.update_cpu_rule(w, b, ogw, ogb, 0.01, "adam", mm, vv)
    for k in w:
        assert np.abs(m.weights[k] - w[k]).max() / np.abs(w[k]).max() < 1e-3, k
    assert m.t == 0


@pytest.mark.parametrize("rule", ["cpu", "cupy"])
def test_checkpoint_resume_continues_the_run(tmp_path, rule):
    """an additive sidecar (model dict + Adam m, v, t + epoch + host RNG state).  2 epochs, checkpoint,
    fresh model + load_checkpoint + 2 more epochs == 4 epochs straight (Adam, FP32 mode; both update rules: the CuPy
    rule's bias correction needs the restored step count)."""
    from pycnn_b200 import PyCNN
    rng = np.random.default_rng(9)
    imgs = rng.integers(0, 256, (96, 32, 32, 3), dtype=np.uint8)
    lab = rng.integers(0, 3, 96)
    ckpt = str(tmp_path / "run.ckpt")

    def fresh(epochs):
        m = PyCNN()
        with contextlib.redirect_stdout(io.StringIO()):
            m.init(batch_size=32, layers=[24, 12], learning_rate=1e-3, epochs=epochs)
            m.adam()
            m.set_math_mode("fp32")
            m.set_update_rule(rule)
            m.cuda(True)
            np.random.seed(42)
            m.dataset.synthetic(imgs, lab, ["a", "b", "c"])
        return m

    straight = fresh(4)
    np.random.seed(7)
    quiet(straight.train_model)

    first = fresh(2)
    np.random.seed(7)
    quiet(first.train_model, checkpoint=ckpt)
    assert os.path.exists(ckpt) and not os.path.exists(ckpt + ".tmp")

    resumed = fresh(4)
    np.random.seed(999)                                   # the checkpoint, not this, decides the next permutation
    (epoch, _) = quiet(resumed.load_checkpoint, ckpt)
    assert epoch == 2 and resumed.optimizer == "adam"
    resumed.epochs = 4
    _, out = quiet(resumed.train_model, start_epoch=epoch)
    assert "epoch=3/4" in out and "epoch=4/4" in out and "epoch=1/4" not in out
    for k in straight.weights:
        assert np.abs(resumed.weights[k] - straight.weights[k]).max() <= 2e-6 * np.abs(straight.weights[k]).max(), k
    for k in straight.biases:
        assert np.abs(resumed.biases[k] - straight.biases[k]).max() <= 2e-6 * max(np.abs(straight.biases[k]).max(), 1e-3), k
    # the sidecar does not replace model.bin: load_model refuses nothing it used to accept, load_checkpoint refuses model.bin
    mb = str(tmp_path / "model.bin")
    quiet(resumed.save_model, mb)
    with pytest.raises(ValueError):
        resumed.load_checkpoint(mb)
