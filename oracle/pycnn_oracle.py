"""numpy/float64 restatement of the PyCNN hot path (TEST INFRASTRUCTURE, see oracle/__init__.py).

Every function cites the reference file:line (relative to /root/reference) it restates.
The arithmetic is float64 end to end, like the reference (DTYPE = np.float64 in all four
.pyx modules; ``dtype=float`` at pycnn/pycnn.py:207-211,273).

Parity status: PINNED by tests/test_oracle.py against tests/golden/*.npz, which were
produced by executing the reference itself (tests/golden/make_golden.py), and against the
reference's max-pool known-answer test (tests/test.py:164-181).
"""
from __future__ import annotations

import numpy as np

# pycnn/pycnn.py:71-91 -- the default bank of 19 fixed 3x3 taps (user-replaceable via init(filters=)).
DEFAULT_FILTERS = [
    [[0, -1, 0], [-1, 5, -1], [0, -1, 0]],
    [[1, 0, -1], [1, 0, -1], [1, 0, -1]],
    [[-1, -1, -1], [-1, 8, -1], [-1, -1, -1]],
    [[-1, 0, 1], [-2, 0, 2], [-1, 0, 1]],
    [[-1, -2, -1], [0, 0, 0], [1, 2, 1]],
    [[1, 1, 1], [1, -8, 1], [1, 1, 1]],
    [[0.111, 0.111, 0.111], [0.111, 0.111, 0.111], [0.111, 0.111, 0.111]],
    [[0, 1, 0], [1, -4, 1], [0, 1, 0]],
    [[-1, 2, -1], [-1, 2, -1], [-1, 2, -1]],
    [[-1, -1, -1], [2, 2, 2], [-1, -1, -1]],
    [[2, -1, 0], [-1, 2, -1], [0, -1, 2]],
    [[0, -1, 2], [-1, 2, -1], [2, -1, 0]],
    [[-1, -1, 2], [-1, 2, -1], [2, -1, -1]],
    [[1, -2, 1], [-2, 5, -2], [1, -2, 1]],
    [[-1, 0, 1], [-1, 0, 1], [-1, 0, 1]],
    [[-1, -1, -1], [0, 0, 0], [1, 1, 1]],
    [[0, -0.5, 0], [-0.5, 3, -0.5], [0, -0.5, 0]],
    [[1, 1, 0], [1, -4, 0], [0, 0, 0]],
    [[0, -1, -2], [1, 1, -1], [2, 1, 0]],
]

MAX_GRAD_NORM = 5.0  # pycnn/modules/gradient_decent.pyx:26


# --------------------------------------------------------------------------------------
# Feature extractor
# --------------------------------------------------------------------------------------
def conv2d_valid(x: np.ndarray, k: np.ndarray) -> np.ndarray:
    """True 2-D convolution, 'valid' mode == scipy.signal.convolve2d(x, k, mode='valid')
    as called at pycnn/pycnn.py:280-282.  The kernel is flipped by 180 degrees:
    out[i, j] = sum_{a,b} x[i+a, j+b] * k[kh-1-a, kw-1-b].
    Accumulation order is row-major over (a, b); scipy's own order may differ in the last ulp.
    """
    x = np.asarray(x, dtype=np.float64)
    k = np.asarray(k, dtype=np.float64)
    kh, kw = k.shape
    oh, ow = x.shape[0] - kh + 1, x.shape[1] - kw + 1
    out = np.zeros((oh, ow), dtype=np.float64)
    for a in range(kh):
        for b in range(kw):
            out += x[a:a + oh, b:b + ow] * k[kh - 1 - a, kw - 1 - b]
    return out


def max_pool2x2(fm: np.ndarray) -> np.ndarray:
    """2x2 / stride-2 max-pool with floor crop.  pycnn/modules/max_pooling.pyx:10-30
    (CuPy branch pycnn/pycnn.py:252-263 computes the same thing)."""
    fm = np.asarray(fm, dtype=np.float64)
    nh, nw = fm.shape[0] // 2, fm.shape[1] // 2
    return fm[:nh * 2, :nw * 2].reshape(nh, 2, nw, 2).max(axis=(1, 3))


def extract_features(img, filters=DEFAULT_FILTERS) -> np.ndarray:
    """pycnn/pycnn.py:267-287 (_preprocess_image, array input branch :270-271).
    img: HxWx3 (uint8 or anything np.asarray takes) -> float64 [F * H' * W'], filter-major.
    /255 (:273); per filter conv(R)+conv(G)+conv(B) (:280-282); ReLU (:283); pool (:284); flatten (:287)."""
    a = np.asarray(img).astype(np.float64) / 255.0
    r, g, b = a[..., 0], a[..., 1], a[..., 2]
    maps = []
    for filt in filters:
        k = np.asarray(filt, dtype=np.float64)
        conv = conv2d_valid(r, k) + conv2d_valid(g, k) + conv2d_valid(b, k)
        conv = np.maximum(conv, 0)
        maps.append(max_pool2x2(conv))
    return np.array(maps).reshape(-1)


def extract_batch(images, filters=DEFAULT_FILTERS) -> np.ndarray:
    """Vectorised over a batch [N,H,W,3]: same arithmetic per image as extract_features
    (per-channel convolutions summed in R,G,B order), used where a python loop is too slow."""
    a = np.asarray(images).astype(np.float64) / 255.0
    n, h, w, _ = a.shape
    oh, ow = h - 2, w - 2
    ph, pw = oh // 2, ow // 2
    out = np.empty((n, len(filters), ph, pw), dtype=np.float64)
    for fi, filt in enumerate(filters):
        k = np.asarray(filt, dtype=np.float64)
        conv = None
        for c in range(3):
            acc = np.zeros((n, oh, ow), dtype=np.float64)
            for da in range(3):
                for db in range(3):
                    acc += a[:, da:da + oh, db:db + ow, c] * k[2 - da, 2 - db]
            conv = acc if conv is None else conv + acc
        conv = np.maximum(conv, 0)
        out[:, fi] = conv[:, :ph * 2, :pw * 2].reshape(n, ph, 2, pw, 2).max(axis=(2, 4))
    return out.reshape(n, -1)


def feature_count(image_size: int, num_filters: int) -> int:
    """pycnn/pycnn.py:330,422."""
    return ((image_size - 2) // 2) ** 2 * num_filters


# --------------------------------------------------------------------------------------
# Weights
# --------------------------------------------------------------------------------------
def layer_keys(num_hidden: int):
    ws = [f"w{i + 1}" for i in range(num_hidden)] + ["wo"]
    bs = [f"b{i + 1}" for i in range(num_hidden)] + ["bo"]
    return ws, bs


def init_layers(input_size: int, layers, num_classes: int, seed: int | None = 42):
    """pycnn/pycnn.py:225 (np.random.seed(42) in init()) + :293-311 (_init_layers).
    He-normal randn(out,in)*sqrt(2/in), bias randn*0.01; draw order w1,b1,...,wo,bo from the
    GLOBAL numpy stream.  seed=None keeps drawing from the current global stream."""
    if seed is not None:
        np.random.seed(seed)
    weights, biases = {}, {}
    prev = input_size
    for i, size in enumerate(layers):
        weights[f"w{i + 1}"] = np.random.randn(size, prev) * (2 / prev) ** 0.5
        biases[f"b{i + 1}"] = np.random.randn(size) * 0.01
        prev = size
    weights["wo"] = np.random.randn(num_classes, prev) * (2 / prev) ** 0.5
    biases["bo"] = np.random.randn(num_classes) * 0.01
    return weights, biases


# --------------------------------------------------------------------------------------
# Forward / loss
# --------------------------------------------------------------------------------------
def softmax(x: np.ndarray) -> np.ndarray:
    """pycnn/lib/optimized.cpp:9-31 (row max-subtracted softmax); identical maths at
    pycnn/modules/forward_pass.pyx:40-41."""
    x = np.asarray(x, dtype=np.float64)
    e = np.exp(x - np.max(x, axis=1, keepdims=True))
    return e / np.sum(e, axis=1, keepdims=True)


def cross_entropy(preds: np.ndarray, labels: np.ndarray) -> np.ndarray:
    """pycnn/lib/optimized.cpp:34-46: out[i] = -sum_j y_ij * log(p_ij + 1e-9)."""
    preds = np.asarray(preds, dtype=np.float64)
    labels = np.asarray(labels, dtype=np.float64)
    return -(labels * np.log(preds + 1e-9)).sum(axis=1)


def forward(batch_x, weights, biases, layers):
    """pycnn/modules/forward_pass.pyx:15-43.  Returns (probs, activations) where
    activations[0] = batch_x and activations[i+1] is the post-ReLU output of hidden layer i
    (z_values[i] aliases it in the reference, :31-34).  Also returns the output logits."""
    acts = [np.asarray(batch_x, dtype=np.float64)]
    cur = acts[0]
    for i in range(len(layers)):
        z = cur @ weights[f"w{i + 1}"].T
        z = z + biases[f"b{i + 1}"]
        z[z < 0] = 0  # relu_inplace, :8-13
        cur = z
        acts.append(cur)
    logits = cur @ weights["wo"].T + biases["bo"]
    return softmax(logits), acts, logits


def predict_single(feat, weights, biases, layers):
    """pycnn/pycnn.py:744-759: GEMV chain on one feature vector -> (argmax index, confidence)."""
    cur = np.asarray(feat, dtype=np.float64)
    for i in range(len(layers)):
        cur = np.maximum(np.dot(weights[f"w{i + 1}"], cur) + biases[f"b{i + 1}"], 0)
    logits = np.dot(weights["wo"], cur) + biases["bo"]
    p = softmax(logits.reshape(1, -1))[0]
    idx = int(np.argmax(p))
    return idx, float(p[idx])


# --------------------------------------------------------------------------------------
# Backward
# --------------------------------------------------------------------------------------
def backward(output, batch_y, acts, weights, layers):
    """pycnn/modules/backward_pass.pyx:37-97.  Gradients are SUMS over the batch (no 1/bs):
    dZ=p-y (:65); dW=dZ^T.A (:71,92); db=sum_rows dZ (:72,93); dA=dZ.W (:75,97);
    mask where z<=0 (:85-86; z aliases the post-ReLU activation); no dA for layer 1 (:95)."""
    gw, gb = {}, {}
    dz = np.asarray(output, dtype=np.float64) - np.asarray(batch_y, dtype=np.float64)
    gw["wo"] = dz.T @ acts[-1]
    gb["bo"] = dz.sum(axis=0)
    da = dz @ weights["wo"]
    for i in range(len(layers) - 1, -1, -1):
        dz = da.copy()
        dz[acts[i + 1] <= 0] = 0.0
        gw[f"w{i + 1}"] = dz.T @ acts[i]
        gb[f"b{i + 1}"] = dz.sum(axis=0)
        if i > 0:
            da = dz @ weights[f"w{i + 1}"]
    return gw, gb


# --------------------------------------------------------------------------------------
# Optimizer
# --------------------------------------------------------------------------------------
def _clip(g):
    """pycnn/modules/gradient_decent.pyx:33-35: per-tensor L2 clip to 5.0."""
    n = np.linalg.norm(g)
    if n > MAX_GRAD_NORM:
        g = g * (MAX_GRAD_NORM / n)
    return g


def update_cpu_rule(weights, biases, gw, gb, lr, optimizer, m, v, beta1=0.9, beta2=0.999, eps=1e-8, t=0):
    """pycnn/modules/gradient_decent.pyx:7-149 -- the CPU update rule, the parity target.
    Per tensor: clip to L2 5.0 (:33-35); NaN/Inf -> skip tensor (:37-39); SGD w -= lr*g (:41);
    Adam with t<=0 -> 1 (:58-59; train_model never advances t, pycnn/pycnn.py:564-570, so the
    bias corrections stay 1-beta1 and 1-beta2), v_hat floored at eps^2 (:98), NaN/Inf update ->
    skip after m,v were already mutated (:103-105).  Mutates the dicts like the reference."""
    if optimizer == "sgd":
        for params, grads in ((weights, gw), (biases, gb)):
            for key in grads:
                g = _clip(grads[key])
                if np.any(np.isnan(g)) or np.any(np.isinf(g)):
                    continue
                params[key] = params[key] - lr * g
        return
    if t <= 0:
        t = 1
    bc1 = 1.0 - beta1 ** t
    bc2 = 1.0 - beta2 ** t
    if abs(bc1) < 1e-8:
        bc1 = 1e-8
    if abs(bc2) < 1e-8:
        bc2 = 1e-8
    for params, grads in ((weights, gw), (biases, gb)):
        for key in grads:
            g = _clip(grads[key])
            if np.any(np.isnan(g)) or np.any(np.isinf(g)):
                continue
            if key not in m:
                m[key] = np.zeros_like(g)
            if key not in v:
                v[key] = np.zeros_like(g)
            m[key] *= beta1
            m[key] += (1.0 - beta1) * g
            v[key] *= beta2
            v[key] += (1.0 - beta2) * (g * g)
            m_hat = m[key] / bc1
            v_hat = np.maximum(v[key] / bc2, eps * eps)
            upd = lr * m_hat / (np.sqrt(v_hat) + eps)
            if np.any(np.isnan(upd)) or np.any(np.isinf(upd)):
                continue
            params[key] = params[key] - upd


def update_cupy_rule(weights, biases, gw, gb, lr, optimizer, m, v, beta1=0.9, beta2=0.999, eps=1e-8, t=0):
    """pycnn/pycnn.py:526-562 -- the CuPy branch's rule: no clip, no NaN skip, t advances.
    Returns the new t.  Offered by the product behind set_update_rule(CUPY)."""
    if optimizer == "sgd":
        for key in gw:
            weights[key] = weights[key] - lr * gw[key]
        for key in gb:
            biases[key] = biases[key] - lr * gb[key]
        return t
    t += 1
    for params, grads in ((weights, gw), (biases, gb)):
        for key in grads:
            g = grads[key]
            m[key] = m[key] * beta1 + (1 - beta1) * g
            v[key] = v[key] * beta2 + (1 - beta2) * (g * g)
            m_hat = m[key] / (1 - beta1 ** t)
            v_hat = v[key] / (1 - beta2 ** t)
            params[key] = params[key] - lr * m_hat / (np.sqrt(v_hat) + eps)
    return t


# --------------------------------------------------------------------------------------
# Training loop
# --------------------------------------------------------------------------------------
def train_step(batch_x, batch_y, weights, biases, layers, lr, optimizer, m, v,
               beta1=0.9, beta2=0.999, eps=1e-8, rule="cpu", t=0):
    """One iteration of the batch loop, pycnn/pycnn.py:629-642.
    Returns (sum of per-row CE losses, #correct, new t)."""
    out, acts, _ = forward(batch_x, weights, biases, layers)
    loss_sum = float(cross_entropy(out, batch_y).sum())                       # :634-635
    correct = int(np.sum(np.argmax(out, axis=1) == np.argmax(batch_y, axis=1)))  # :637-639
    gw, gb = backward(out, batch_y, acts, weights, layers)
    if rule == "cpu":
        update_cpu_rule(weights, biases, gw, gb, lr, optimizer, m, v, beta1, beta2, eps, 0)
    else:
        t = update_cupy_rule(weights, biases, gw, gb, lr, optimizer, m, v, beta1, beta2, eps, t)
    return loss_sum, correct, t


def train_epochs(data, labels, weights, biases, layers, lr, optimizer, batch_size, epochs,
                 beta1=0.9, beta2=0.999, eps=1e-8, rule="cpu", perms=None, max_steps=None,
                 early_stop=0):
    """pycnn/pycnn.py:621-681: per epoch a permutation from the global numpy stream (:622)
    unless ``perms`` supplies them; batch loop (:626-642); epoch loss = sum/num_samples (:644);
    early stop on loss < best-1e-8 with patience (:651-659); hard stop at loss<=1e-6 and
    acc>=0.999999 (:679-681).  Returns dict(step_losses, epoch_loss, epoch_correct, perms)."""
    data = np.asarray(data, dtype=np.float64)
    labels = np.asarray(labels)
    n = len(data)
    m, v = {}, {}
    if optimizer == "adam":
        for d in (weights, biases):
            for key in d:
                m[key] = np.zeros(d[key].shape)
                v[key] = np.zeros(d[key].shape)
    step_losses, epoch_loss, epoch_correct, used_perms = [], [], [], []
    best, stop_counter, t, steps = float("inf"), 0, 0, 0
    for epoch in range(epochs):
        perm = np.random.permutation(n) if perms is None else np.asarray(perms[epoch])
        used_perms.append(perm)
        correct, total = 0, 0.0
        done = False
        for start in range(0, n, batch_size):
            idx = perm[start:min(start + batch_size, n)]
            ls, c, t = train_step(data[idx], labels[idx], weights, biases, layers, lr, optimizer,
                                  m, v, beta1, beta2, eps, rule, t)
            step_losses.append(ls)
            total += ls
            correct += c
            steps += 1
            if max_steps is not None and steps >= max_steps:
                done = True
                break
        el, ea = total / max(1, n), correct / max(1, n)
        epoch_loss.append(el)
        epoch_correct.append(correct)
        if done:
            break
        if early_stop:
            if el < best - 1e-8:
                best, stop_counter = el, 0
            else:
                stop_counter += 1
                if stop_counter >= early_stop:
                    break
        if el <= 1e-6 and ea >= 0.999999:
            break
    return dict(step_losses=np.array(step_losses), epoch_loss=np.array(epoch_loss),
                epoch_correct=np.array(epoch_correct), perms=used_perms, m=m, v=v, t=t)


def one_hot(label_idx, num_classes):
    """pycnn/pycnn.py:353-355: int one-hot rows."""
    y = np.zeros((len(label_idx), num_classes), dtype=np.int64)
    y[np.arange(len(label_idx)), np.asarray(label_idx)] = 1
    return y
