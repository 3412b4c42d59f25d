"""Generate golden vectors by EXECUTING THE REFERENCE ITSELF (run in the build container only).

    python tests/golden/make_golden.py

Needs /root/reference (read-only) and oracle/_ref (built by oracle/build_ref.py from the
reference's own sources).  The reference's PyCNN class is run unmodified through
oracle/ref_loader.py; every array below comes out of the reference's own methods:
  _preprocess_image (pycnn/pycnn.py:267), _forward_pass (:474), _cross_entropy_batch (:245),
  _backward_pass (:498), _gradient_decent (:525), train_model (:572), predict (:737),
  DatasetLoader._init_layers (:293), save_model (:688).
Inputs are seeded exactly as SURVEY.md section 8(d) prescribes (images default_rng(0), labels
default_rng(1), custom filters default_rng(2); weights from the reference's np.random.seed(42) init).
Outputs: small .npz fixtures next to this file (committed).
"""
from __future__ import annotations

import contextlib
import io
import os
import pickle
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))

from oracle import ref_loader  # noqa: E402


def quiet():
    return contextlib.redirect_stdout(io.StringIO())


def custom_filters(F=8):
    return np.round(np.random.default_rng(2).uniform(-2, 2, (F, 3, 3)), 3).tolist()


def synth_images(n, s, seed=0):
    return np.random.default_rng(seed).integers(0, 256, (n, s, s, 3), dtype=np.uint8)


def synth_labels(n, c, seed=1):
    return np.random.default_rng(seed).integers(0, c, n)


def new_model(ref, layers, lr, bs, epochs, filters=None, adam=False):
    m = ref.PyCNN()
    with quiet():
        m.init(batch_size=bs, layers=list(layers), learning_rate=lr, epochs=epochs, filters=filters)
        if adam:
            m.adam()
    return m


def gen_extract(ref):
    out = {}
    m = new_model(ref, [8], 0.01, 4, 1)
    cf = custom_filters()
    out["custom_filters"] = np.array(cf)
    for tag, n, s in (("s32", 3, 32), ("s33", 2, 33), ("s64", 1, 64), ("s7", 2, 7)):
        imgs = synth_images(n, s, seed=10 + s)
        out[f"{tag}_images"] = imgs
        out[f"{tag}_default"] = np.stack([m._preprocess_image(im, s, m.filters) for im in imgs])
        out[f"{tag}_custom"] = np.stack([m._preprocess_image(im, s, cf) for im in imgs])
    # the reference's own test_06 input: solid red 32x32, first 3 filters (tests/test.py:85-94)
    red = np.zeros((32, 32, 3), dtype=np.uint8)
    red[..., 0] = 255
    out["red_first3"] = m._preprocess_image(red, 32, m.filters[:3])
    np.savez_compressed(os.path.join(HERE, "extract.npz"), **out)
    print("extract.npz", {k: v.shape for k, v in out.items()})


def gen_mlp_small(ref):
    """One forward/backward + 5 SGD and 5 Adam steps on a tiny net through the reference's own
    private methods, and a 3-epoch train_model run (N=20, bs=8 -> short last batch)."""
    D, layers, C, N, bs = 48, [16, 8], 3, 20, 8
    rng = np.random.default_rng(7)
    data = rng.uniform(0, 1.5, (N, D))
    lab = synth_labels(N, C, seed=8)
    y = np.zeros((N, C), dtype=np.int64)
    y[np.arange(N), lab] = 1
    out = dict(data=data, labels=lab, layers=np.array(layers), D=D, C=C, bs=bs)
    for opt in ("sgd", "adam"):
        m = new_model(ref, layers, 0.05 if opt == "sgd" else 0.01, bs, 3, adam=(opt == "adam"))
        with quiet():
            m.dataset._init_layers([str(i) for i in range(C)], D)
        for k, v in m.weights.items():
            out[f"{opt}_init_{k}"] = v.copy()
        for k, v in m.biases.items():
            out[f"{opt}_init_{k}"] = v.copy()
        gw = {k: np.zeros(v.shape) for k, v in m.weights.items()}
        gb = {k: np.zeros(v.shape) for k, v in m.biases.items()}
        acts, zs = [None] * (len(layers) + 1), [None] * (len(layers) + 1)
        losses = []
        for step in range(5):
            idx = np.arange(step * 3, step * 3 + bs) % N
            bx, by = data[idx], y[idx]
            p = m._forward_pass(bx, acts, zs)
            losses.append(m._cross_entropy_batch(p, by).sum())
            m._backward_pass(p, by, acts, zs, gw, gb)
            if step == 0:
                out[f"{opt}_probs0"] = p.copy()
                for k in gw:
                    out[f"{opt}_grad0_{k}"] = gw[k].copy()
                for k in gb:
                    out[f"{opt}_grad0_{k}"] = gb[k].copy()
            m._gradient_decent(gw, gb)
        out[f"{opt}_step_losses"] = np.array(losses)
        out[f"{opt}_t_after"] = m.t
        for k, v in m.weights.items():
            out[f"{opt}_after5_{k}"] = v.copy()
        for k, v in m.biases.items():
            out[f"{opt}_after5_{k}"] = v.copy()
        # full train_model run: fresh model, seed 42 -> init -> 3 epochs of permutations
        m2 = new_model(ref, layers, 0.05 if opt == "sgd" else 0.01, bs, 3, adam=(opt == "adam"))
        with quiet():
            m2.dataset._init_layers([str(i) for i in range(C)], D)
            m2.data = [row for row in data]
            m2.labels = [row for row in y]
            m2.train_model(visualize=False)
        for k, v in m2.weights.items():
            out[f"{opt}_trained_{k}"] = v.copy()
        for k, v in m2.biases.items():
            out[f"{opt}_trained_{k}"] = v.copy()
        # predict() on a feature vector: reproduce pycnn/pycnn.py:744-759 via the model's own dot/softmax
        cur = data[0]
        for i in range(len(layers)):
            cur = m2.maximum(m2.dot(m2.weights[f"w{i+1}"], cur) + m2.biases[f"b{i+1}"], 0)
        logits = m2.dot(m2.weights["wo"], cur) + m2.biases["bo"]
        out[f"{opt}_trained_probs_row0"] = m2._softmax_batch(logits.reshape(1, -1))[0]
    np.savez_compressed(os.path.join(HERE, "mlp_small.npz"), **out)
    print("mlp_small.npz", len(out), "arrays")


def gen_traj_c1(ref):
    """BASELINE config 1: 32x32 synthetic, 10 classes, layers [256,128,64], bs 32, N=2048,
    100 train steps through the reference's train_model internals; SGD lr 1e-3 (pycnn/pycnn.py:215)
    and Adam lr 1e-4.  Stores per-step loss sums, #correct and final-weight digests only."""
    S, C, layers, bs, N, steps = 32, 10, [256, 128, 64], 32, 2048, 100
    imgs = synth_images(N, S, seed=0)
    lab = synth_labels(N, C, seed=1)
    y = np.zeros((N, C), dtype=np.int64)
    y[np.arange(N), lab] = 1
    out = dict(S=S, C=C, layers=np.array(layers), bs=bs, N=N, steps=steps)
    m0 = new_model(ref, layers, 1e-3, bs, 1)
    feats = np.stack([m0._preprocess_image(im, S, m0.filters) for im in imgs])
    out["feat_digest"] = np.array([feats.sum(), np.abs(feats).sum(), feats.max(), feats[123, 456], feats[2047, 4274]])
    for opt, lr in (("sgd", 1e-3), ("adam", 1e-4)):
        m = new_model(ref, layers, lr, bs, 1, adam=(opt == "adam"))
        with quiet():
            m.dataset._init_layers([str(i) for i in range(C)], feats.shape[1])
        gw = {k: np.zeros(v.shape) for k, v in m.weights.items()}
        gb = {k: np.zeros(v.shape) for k, v in m.biases.items()}
        acts, zs = [None] * (len(layers) + 1), [None] * (len(layers) + 1)
        losses, corrects, perms = [], [], []
        step = 0
        while step < steps:
            perm = m.permutation(N)        # pycnn/pycnn.py:622, global stream after seed 42 + init
            perms.append(perm)
            for start in range(0, N, bs):
                idx = perm[start:start + bs]
                bx, by = feats[idx], y[idx]
                p = m._forward_pass(bx, acts, zs)
                losses.append(m._cross_entropy_batch(p, by).sum())
                corrects.append(int(np.sum(np.argmax(p, 1) == np.argmax(by, 1))))
                m._backward_pass(p, by, acts, zs, gw, gb)
                m._gradient_decent(gw, gb)
                step += 1
                if step >= steps:
                    break
        out[f"{opt}_step_losses"] = np.array(losses)
        out[f"{opt}_step_correct"] = np.array(corrects)
        out[f"{opt}_perm0_head"] = perms[0][:64]
        for k in list(m.weights) + list(m.biases):
            w = m.weights[k] if k in m.weights else m.biases[k]
            out[f"{opt}_digest_{k}"] = np.array([w.sum(), np.abs(w).sum(), np.sqrt((w * w).sum())])
        # logits/probs of the first 64 images after training, for class-parity checks
        p = m._forward_pass(feats[:64], acts, zs)
        out[f"{opt}_probs_first64"] = p
    np.savez_compressed(os.path.join(HERE, "traj_c1.npz"), **out)
    print("traj_c1.npz", len(out), "arrays")


def gen_model_bin(ref):
    """A model.bin written by the reference's own save_model (pycnn/pycnn.py:688-706)."""
    m = new_model(ref, [6], 0.01, 4, 1, filters=custom_filters(2))
    with quiet():
        m.dataset._init_layers(["cat", "dog"], ((8 - 2) // 2) ** 2 * 2)
        m.image_size = 8
        path = os.path.join(HERE, "ref_model.bin")
        m.save_model(path)
    with open(path, "rb") as f:
        d = pickle.load(f)
    print("ref_model.bin keys", list(d), "protocol byte", open(path, "rb").read(2))


if __name__ == "__main__":
    ref = ref_loader.load()
    gen_extract(ref)
    gen_mlp_small(ref)
    gen_traj_c1(ref)
    gen_model_bin(ref)
