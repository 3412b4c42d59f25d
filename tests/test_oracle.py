"""CPU: pin oracle/pycnn_oracle.py against golden vectors produced by the reference itself
(tests/golden/make_golden.py) and against the reference's own known-answer tests."""
import os
import pickle

import numpy as np
import pytest

from oracle import pycnn_oracle as O
from oracle import ref_kernels, ref_loader
from tests.helpers import GOLDEN, load_golden, synth_images, synth_labels


def test_maxpool_known_answer():
    # the reference's only numeric KAT: tests/test.py:164-181
    fm = np.arange(1, 17, dtype=float).reshape(4, 4)
    np.testing.assert_array_equal(O.max_pool2x2(fm), [[6, 8], [14, 16]])
    # floor crop on odd sizes (pycnn/modules/max_pooling.pyx:13-14)
    assert O.max_pool2x2(np.arange(25, dtype=float).reshape(5, 5)).shape == (2, 2)


def test_softmax_and_ce_properties():
    # tests/test.py:96-104 and :154-162
    x = np.array([[1.0, 2.0, 3.0], [4.0, 5.0, 6.0]])
    np.testing.assert_allclose(O.softmax(x).sum(axis=1), [1.0, 1.0], atol=1e-12)
    preds = np.array([[0.7, 0.2, 0.1], [0.1, 0.8, 0.1]])
    labels = np.array([[1.0, 0.0, 0.0], [0.0, 1.0, 0.0]])
    loss = O.cross_entropy(preds, labels)
    np.testing.assert_allclose(loss, -np.log(np.array([0.7, 0.8]) + 1e-9), rtol=1e-14)


@pytest.mark.parametrize("tag,s", [("s32", 32), ("s33", 33), ("s64", 64), ("s7", 7)])
def test_extractor_matches_reference_golden(tag, s):
    g = load_golden("extract.npz")
    imgs = g[f"{tag}_images"]
    cf = g["custom_filters"].tolist()
    for name, filt in (("default", O.DEFAULT_FILTERS), ("custom", cf)):
        want = g[f"{tag}_{name}"]
        got = np.stack([O.extract_features(im, filt) for im in imgs])
        assert got.shape == want.shape == (len(imgs), O.feature_count(s, len(filt)))
        np.testing.assert_allclose(got, want, rtol=0, atol=2e-14)
        np.testing.assert_allclose(O.extract_batch(imgs, filt), want, rtol=0, atol=2e-14)


def test_extractor_red_image_first3():
    g = load_golden("extract.npz")
    red = np.zeros((32, 32, 3), dtype=np.uint8)
    red[..., 0] = 255
    np.testing.assert_allclose(O.extract_features(red, O.DEFAULT_FILTERS[:3]), g["red_first3"], atol=1e-14)


def _small_setup(g):
    layers = g["layers"].tolist()
    D, C, bs = int(g["D"]), int(g["C"]), int(g["bs"])
    data, lab = g["data"], g["labels"]
    return layers, D, C, bs, data, O.one_hot(lab, C)


@pytest.mark.parametrize("opt,lr", [("sgd", 0.05), ("adam", 0.01)])
def test_mlp_small_matches_reference_golden(opt, lr):
    g = load_golden("mlp_small.npz")
    layers, D, C, bs, data, y = _small_setup(g)
    w, b = O.init_layers(D, layers, C, seed=42)
    wk, bk = O.layer_keys(len(layers))
    for k in wk:
        np.testing.assert_array_equal(w[k], g[f"{opt}_init_{k}"])   # same RNG stream as the reference
    for k in bk:
        np.testing.assert_array_equal(b[k], g[f"{opt}_init_{k}"])
    m, v = {}, {}
    if opt == "adam":
        for d in (w, b):
            for k in d:
                m[k], v[k] = np.zeros(d[k].shape), np.zeros(d[k].shape)
    losses = []
    for step in range(5):
        idx = np.arange(step * 3, step * 3 + bs) % len(data)
        p, acts, _ = O.forward(data[idx], w, b, layers)
        losses.append(O.cross_entropy(p, y[idx]).sum())
        gw, gb = O.backward(p, y[idx], acts, w, layers)
        if step == 0:
            np.testing.assert_allclose(p, g[f"{opt}_probs0"], atol=1e-14)
            for k in wk:
                np.testing.assert_allclose(gw[k], g[f"{opt}_grad0_{k}"], atol=1e-13)
            for k in bk:
                np.testing.assert_allclose(gb[k], g[f"{opt}_grad0_{k}"], atol=1e-13)
        O.update_cpu_rule(w, b, gw, gb, lr, opt, m, v)
    np.testing.assert_allclose(losses, g[f"{opt}_step_losses"], rtol=1e-12)
    assert int(g[f"{opt}_t_after"]) == 0          # the CPU path never advances t
    for k in wk:
        np.testing.assert_allclose(w[k], g[f"{opt}_after5_{k}"], atol=1e-13)
    for k in bk:
        np.testing.assert_allclose(b[k], g[f"{opt}_after5_{k}"], atol=1e-13)


@pytest.mark.parametrize("opt,lr", [("sgd", 0.05), ("adam", 0.01)])
def test_train_loop_matches_reference_train_model(opt, lr):
    """Seed 42 -> init -> 3 epochs of np.random.permutation, N=20, bs=8 (short last batch)."""
    g = load_golden("mlp_small.npz")
    layers, D, C, bs, data, y = _small_setup(g)
    w, b = O.init_layers(D, layers, C, seed=42)
    O.train_epochs(data, y, w, b, layers, lr, opt, bs, epochs=3)
    for k in list(w) + list(b):
        got = w[k] if k in w else b[k]
        np.testing.assert_allclose(got, g[f"{opt}_trained_{k}"], atol=1e-12)
    idx, conf = O.predict_single(data[0], w, b, layers)
    want = g[f"{opt}_trained_probs_row0"]
    assert idx == int(np.argmax(want)) and abs(conf - want.max()) < 1e-12


@pytest.mark.parametrize("opt,lr", [("sgd", 1e-3), ("adam", 1e-4)])
def test_c1_trajectory_matches_reference(opt, lr):
    """BASELINE config 1 over 100 steps: the restatement follows the reference's loss trajectory."""
    g = load_golden("traj_c1.npz")
    S, C, bs, N, steps = int(g["S"]), int(g["C"]), int(g["bs"]), int(g["N"]), int(g["steps"])
    layers = g["layers"].tolist()
    feats = O.extract_batch(synth_images(N, S, 0), O.DEFAULT_FILTERS)
    d = g["feat_digest"]
    np.testing.assert_allclose([feats.sum(), np.abs(feats).sum(), feats.max(), feats[123, 456], feats[2047, 4274]],
                               d, rtol=1e-12)
    y = O.one_hot(synth_labels(N, C, 1), C)
    w, b = O.init_layers(feats.shape[1], layers, C, seed=42)
    res = O.train_epochs(feats, y, w, b, layers, lr, opt, bs, epochs=10, max_steps=steps)
    np.testing.assert_array_equal(res["perms"][0][:64], g[f"{opt}_perm0_head"])
    np.testing.assert_allclose(res["step_losses"], g[f"{opt}_step_losses"], rtol=1e-9)
    for k in list(w) + list(b):
        t = w[k] if k in w else b[k]
        np.testing.assert_allclose([t.sum(), np.abs(t).sum(), np.sqrt((t * t).sum())],
                                   g[f"{opt}_digest_{k}"], rtol=1e-9)
    p, _, _ = O.forward(feats[:64], w, b, layers)
    np.testing.assert_allclose(p, g[f"{opt}_probs_first64"], atol=1e-10)


def test_reference_model_bin_layout(golden_dir):
    """model.bin written by the reference's save_model: pickle protocol 4 dict."""
    with open(os.path.join(golden_dir, "ref_model.bin"), "rb") as f:
        raw = f.read()
    assert raw[:2] == b"\x80\x04"
    d = pickle.loads(raw)
    assert list(d) == ["weights", "biases", "filters", "image_size", "classes", "layers", "trained_with_cuda"]
    assert d["weights"]["w1"].dtype == np.float64 and d["weights"]["w1"].shape == (6, 18)
    assert d["classes"] == ["cat", "dog"] and d["layers"] == [6] and d["image_size"] == 8
    assert d["trained_with_cuda"] is False


@pytest.mark.skipif(not ref_kernels.available(), reason="oracle/_ref not built")
def test_compiled_reference_kernels_agree_with_restatement():
    """oracle/_ref (the reference's compiled Cython/C kernels) vs the numpy restatement."""
    g = load_golden("mlp_small.npz")
    layers, D, C, bs, data, y = _small_setup(g)
    for opt, lr in (("sgd", 0.05), ("adam", 0.01)):
        w, b = O.init_layers(D, layers, C, seed=42)
        tr = ref_kernels.RefTrainer(w, b, layers, lr, opt)
        m, v = {}, {}
        if opt == "adam":
            for d_ in (w, b):
                for k in d_:
                    m[k], v[k] = np.zeros(d_[k].shape), np.zeros(d_[k].shape)
        for step in range(5):
            idx = np.arange(step * 3, step * 3 + bs) % len(data)
            l_ref, c_ref = tr.step(data[idx], y[idx])
            l_o, c_o, _ = O.train_step(data[idx], y[idx], w, b, layers, lr, opt, m, v)
            assert c_ref == c_o and abs(l_ref - l_o) < 1e-10
        for k in w:
            np.testing.assert_allclose(tr.weights[k], w[k], atol=1e-13)
    img = synth_images(1, 32, 5)[0]
    np.testing.assert_allclose(ref_kernels.extract_features(img, O.DEFAULT_FILTERS),
                               O.extract_features(img, O.DEFAULT_FILTERS), atol=2e-14)
    x = np.random.default_rng(3).normal(size=(5, 7))
    np.testing.assert_allclose(ref_kernels.softmax(x), O.softmax(x), atol=1e-15)


@pytest.mark.skipif(not ref_loader.available(), reason="/root/reference only exists in the build container")
def test_live_reference_agrees_with_restatement():
    """Executes the unmodified reference class here and compares a fresh (non-golden) case."""
    import contextlib, io
    ref = ref_loader.load()
    m = ref.PyCNN()
    with contextlib.redirect_stdout(io.StringIO()):
        m.init(layers=[8])
    img = synth_images(1, 21, 99)[0]
    np.testing.assert_allclose(m._preprocess_image(img, 21, m.filters),
                               O.extract_features(img, O.DEFAULT_FILTERS), atol=2e-14)
    assert m.filters == O.DEFAULT_FILTERS
