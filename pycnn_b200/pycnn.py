"""Host-side mirror of the reference's public surface for its `cuda(True)` path.

Same class and method names, argument meaning, printed lines and error behaviour as
/root/reference/pycnn/pycnn.py (PyCNN :68-773, DatasetLoader :289-472, Evaluate :779-914), but every
compute step goes through the C ABI of libpycnn_b200.so (include/pycnn_b200.h): there is no CuPy, no
NumPy compute path and no CPU fallback.  What differs from the reference, on purpose:

* `cuda(True)` RAISES when the library or a B200 is missing (reference: silent CPU fallback, :127-132).
* Without `cuda(True)` the compute methods raise: this package is only the GPU backend.
* Weight init and the per-epoch permutation stay on the HOST numpy stream seeded by `init()` (:225, :303-311,
  :622) so trajectories are comparable with the reference's CPU path; the default update rule is the
  CPU one (clip 5.0, frozen t; pycnn/modules/gradient_decent.pyx) -- see SURVEY.md 2.4 items 9-10.
* The dataset lives on the device as a float32 feature matrix; `model.data` is a proxy that converts
  to float64 NumPy on demand.  Loaders decode on the host and extract whole batches in one launch.
"""
from __future__ import annotations

import json
import os
import pickle
import time

import numpy as np

from . import _lib

_IMAGE_EXT_TRAIN = (".png", ".jpg", ".jpeg", ".bmp", ".gif", ".tiff")   # pycnn/pycnn.py:339
_IMAGE_EXT_EVAL = (".png", ".jpg", ".jpeg", ".bmp")                    # pycnn/pycnn.py:804

# The reference's default bank (pycnn/pycnn.py:71-91): data, part of the API's behaviour.
_DEFAULT_FILTERS = (
    ((0, -1, 0), (-1, 5, -1), (0, -1, 0)), ((1, 0, -1), (1, 0, -1), (1, 0, -1)),
    ((-1, -1, -1), (-1, 8, -1), (-1, -1, -1)), ((-1, 0, 1), (-2, 0, 2), (-1, 0, 1)),
    ((-1, -2, -1), (0, 0, 0), (1, 2, 1)), ((1, 1, 1), (1, -8, 1), (1, 1, 1)),
    ((0.111, 0.111, 0.111), (0.111, 0.111, 0.111), (0.111, 0.111, 0.111)),
    ((0, 1, 0), (1, -4, 1), (0, 1, 0)), ((-1, 2, -1), (-1, 2, -1), (-1, 2, -1)),
    ((-1, -1, -1), (2, 2, 2), (-1, -1, -1)), ((2, -1, 0), (-1, 2, -1), (0, -1, 2)),
    ((0, -1, 2), (-1, 2, -1), (2, -1, 0)), ((-1, -1, 2), (-1, 2, -1), (2, -1, -1)),
    ((1, -2, 1), (-2, 5, -2), (1, -2, 1)), ((-1, 0, 1), (-1, 0, 1), (-1, 0, 1)),
    ((-1, -1, -1), (0, 0, 0), (1, 1, 1)), ((0, -0.5, 0), (-0.5, 3, -0.5), (0, -0.5, 0)),
    ((1, 1, 0), (1, -4, 0), (0, 0, 0)), ((0, -1, -2), (1, 1, -1), (2, 1, 0)),
)


def _default_filters():
    return [[list(row) for row in f] for f in _DEFAULT_FILTERS]


def _decode_files(paths, image_size, limit=None, on_ok=None, on_fail=None, workers=None):
    """Decode + RGB-convert + LANCZOS-resize image files on host threads (PIL releases the GIL while it decodes and
    resamples), keeping the reference's sequential semantics (pycnn/pycnn.py:355-372): files are taken in listing
    order, a file that fails is reported and skipped, and at most `limit` SUCCESSFUL images are kept.  Work is issued
    in windows of the number of images still wanted, so nothing beyond the limit is decoded needlessly.  Returns the
    list of uint8 [S,S,3] arrays; on_ok(path, count) / on_fail(path, exc) are called in file order."""
    from concurrent.futures import ThreadPoolExecutor
    paths = list(paths)
    if limit is not None and limit <= 0:
        return []
    workers = workers or min(32, (os.cpu_count() or 4))

    def work(path):
        try:
            return _to_u8_image(path, image_size)
        except Exception as e:          # reported by the caller, in order
            return e

    out, pos = [], 0
    with ThreadPoolExecutor(max_workers=workers) as pool:
        while pos < len(paths) and (limit is None or len(out) < limit):
            want = len(paths) - pos if limit is None else min(len(paths) - pos, max(limit - len(out), 1))
            window = paths[pos:pos + want]
            pos += want
            for path, res in zip(window, pool.map(work, window)):
                if isinstance(res, Exception):
                    if on_fail:
                        on_fail(path, res)
                    continue
                out.append(res)
                if on_ok:
                    on_ok(path, len(out))
    return out


def _to_u8_image(obj, image_size=None):
    """Path -> PIL open/RGB/LANCZOS resize (pycnn/pycnn.py:269); anything else is taken as pixels (:271)."""
    if isinstance(obj, str):
        from PIL import Image
        img = Image.open(obj).convert("RGB").resize((image_size, image_size), Image.Resampling.LANCZOS)
        return np.asarray(img, dtype=np.uint8)
    a = np.asarray(obj)
    if a.ndim != 3 or a.shape[2] < 3:
        raise ValueError(f"expected an RGB image (H, W, 3), got array of shape {a.shape}")
    a = a[..., :3]
    if a.dtype != np.uint8:
        if np.all(a == np.round(a)) and a.min() >= 0 and a.max() <= 255:
            a = a.astype(np.uint8)
        else:
            raise ValueError("pycnn_b200 extracts features from 8-bit RGB pixels; got non-integral values")
    if a.shape[0] != a.shape[1]:
        raise ValueError(f"pycnn_b200 needs square images, got {a.shape[0]}x{a.shape[1]}")
    return np.ascontiguousarray(a)


class DeviceFeatures:
    """`model.data` under cuda(True): rows live in HBM as float32 [N, D]; NumPy view on demand."""

    def __init__(self, ctx, rows, width):
        self._ctx, self._rows, self._width = ctx, int(rows), int(width)

    def __len__(self):
        return self._rows

    @property
    def shape(self):
        return (self._rows, self._width)

    def __array__(self, dtype=None, copy=None):
        a = self._ctx.dataset_get_features(0, self._rows)
        return a if dtype is None else a.astype(dtype)

    def __getitem__(self, key):
        return np.asarray(self)[key]


class PyCNN:
    def __init__(self):
        self.optimizer = "sgd"
        self.filters = _default_filters()
        self.use_cuda = False
        self.dataset = self.DatasetLoader(self)
        self.m, self.v, self.t = {}, {}, 0
        self.beta1, self.beta2, self.eps = 0.9, 0.999, 1e-8
        # host-side helpers the reference exposes as attributes (Evaluate reaches for some of them)
        self.zeros, self.maximum, self.dot, self.random = np.zeros, np.maximum, np.dot, np.random
        self.array, self.exp, self.log = np.array, np.exp, np.log
        self.argmax, self.sum, self.max, self.sqrt = np.argmax, np.sum, np.max, np.sqrt
        self.permutation = np.random.permutation
        self._ctx = None
        self._device = int(os.environ.get("LOCAL_RANK", os.environ.get("PYCNN_B200_DEVICE", "0")))
        self._math_mode = {"fp32": _lib.MATH_FP32_SIMT, "tf32": _lib.MATH_TF32}[
            os.environ.get("PYCNN_B200_MATH", "tf32").lower()]
        self._update_rule = {"cpu": _lib.RULE_CPU, "cupy": _lib.RULE_CUPY}[
            os.environ.get("PYCNN_B200_UPDATE_RULE", "cpu").lower()]
        self._extract_mode = {"simt": _lib.EXTRACT_SIMT, "tc": _lib.EXTRACT_TC}[
            os.environ.get("PYCNN_B200_EXTRACT", "tc").lower()]
        self._loaded_filters = None     # what the device currently holds
        self._net_key = None            # (D, layers, C) the device network was built for
        self._params_on_device = False
        self._opt_key = None
        self._resumed = False           # load_checkpoint() restored device state that train_model must keep
        self._rank, self._world = 0, 1

    # ------------------------------------------------------------------ backend seam (:104-194)
    def cuda(self, use_cuda=False, device=None):
        """Enable the B200 backend.  Raises if libpycnn_b200.so or a B200 is unavailable."""
        if device is not None:
            self._device = int(device)
        if use_cuda and self._ctx is None:
            ctx = _lib.Context(self._device)     # raises BackendError: no silent CPU fallback
            ctx.set_math_mode(self._math_mode)
            ctx.set_update_rule(self._update_rule)
            ctx.set_extract_mode(self._extract_mode)
            self._ctx = ctx
            self._loaded_filters = self._net_key = self._opt_key = None
            self._params_on_device = False
            print("CUDA backend initialized successfully")
        elif not use_cuda and self._ctx is not None:
            if isinstance(getattr(self, "data", None), DeviceFeatures):
                self.data = np.asarray(self.data)
            self._ctx.close()
            self._ctx = None
        self.use_cuda = bool(use_cuda)

    def set_math_mode(self, mode: str):
        """'tf32' (tcgen05 tensor cores, 1e-2) or 'fp32' (SIMT verification mode, 1e-5)."""
        self._math_mode = {"fp32": _lib.MATH_FP32_SIMT, "tf32": _lib.MATH_TF32}[mode.lower()]
        if self._ctx:
            self._ctx.set_math_mode(self._math_mode)

    def set_update_rule(self, rule: str):
        """'cpu' = gradient_decent.pyx semantics (default, the parity target); 'cupy' = pycnn.py:526-562."""
        self._update_rule = {"cpu": _lib.RULE_CPU, "cupy": _lib.RULE_CUPY}[rule.lower()]
        if self._ctx:
            self._ctx.set_update_rule(self._update_rule)

    def set_extract_mode(self, mode: str):
        self._extract_mode = {"simt": _lib.EXTRACT_SIMT, "tc": _lib.EXTRACT_TC}[mode.lower()]
        if self._ctx:
            self._ctx.set_extract_mode(self._extract_mode)

    def _require(self):
        if self._ctx is None:
            raise RuntimeError("pycnn_b200 is the cuda(True) backend only: call model.cuda(True) first "
                               "(there is no CPU fallback; use the reference package for CPU runs)")
        return self._ctx

    def _to_cpu(self, x):
        return np.asarray(x) if hasattr(x, "__array__") else x

    def _to_backend(self, x):
        return np.array(x)

    # ------------------------------------------------------------------ configuration (:196-227)
    def adam(self, beta1=0.9, beta2=0.999, eps=1e-8):
        self.optimizer = "adam"
        self.beta1, self.beta2, self.eps = beta1, beta2, eps
        self.t, self.m, self.v = 0, {}, {}
        self._opt_key = None    # forces a reset of the device-side m, v, t
        print("Adam optimizer enabled.")

    def init(self, batch_size=32, layers=[128, 64], learning_rate=0.001, epochs=50, filters=None):
        self.batch_size, self.learning_rate, self.epochs = batch_size, learning_rate, epochs
        self.layers = layers
        self.image_size = None
        if filters is not None:
            self.filters = filters
        np.random.seed(42)
        print(f"Network initialized with {len(layers)} hidden layers: {layers}")

    # ------------------------------------------------------------------ small drop-ins (:229-265)
    def _softmax_batch(self, x):
        x = np.asarray(x, dtype=np.float64)
        if x.ndim != 2:
            raise ValueError("Input must be a 2D array")
        return self._require().softmax(x)

    def _cross_entropy_batch(self, preds, labels):
        return self._require().cross_entropy(preds, labels)

    def _max_pooling(self, feature_map):
        return self._require().max_pool(feature_map)

    # ------------------------------------------------------------------ extractor (:267-287)
    def _sync_filters(self, filters):
        key = np.asarray(filters, dtype=np.float64).tobytes()
        if key != self._loaded_filters:
            self._require().set_filters(filters)
            self._loaded_filters = key

    def _preprocess_image(self, path, image_size, filters):
        ctx = self._require()
        self._sync_filters(filters)
        return ctx.extract(_to_u8_image(path, image_size))[0]

    # ------------------------------------------------------------------ dataset loading (:289-472)
    class DatasetLoader:
        def __init__(self, parent):
            self.parent = parent

        def _init_layers(self, classes, input_size):
            """He-normal init from the host numpy stream, draw order w1,b1,...,wo,bo (:293-311)."""
            p = self.parent
            p.classes = classes
            p.weights, p.biases = {}, {}
            prev = input_size
            for i, size in enumerate(p.layers):
                p.weights[f"w{i + 1}"] = np.random.randn(size, prev) * (2 / prev) ** 0.5
                p.biases[f"b{i + 1}"] = np.random.randn(size) * 0.01
                prev = size
            p.data, p.labels = [], []
            p.weights["wo"] = np.random.randn(len(classes), prev) * (2 / prev) ** 0.5
            p.biases["bo"] = np.random.randn(len(classes)) * 0.01
            p._params_on_device = False
            if p.optimizer == "adam":
                p.m, p.v = {}, {}
                p._opt_key = None

        def _upload(self, images, label_idx, num_classes, image_size, chunk=8192):
            """One fused conv+ReLU+pool launch per chunk straight into the device feature matrix."""
            p = self.parent
            ctx = p._require()
            p._sync_filters(p.filters)
            p._bind_network()
            n = len(images)
            ctx.dataset_alloc(n)
            for s in range(0, n, chunk):
                ctx.dataset_extract(s, np.stack(images[s:s + chunk]))
            ctx.dataset_set_labels(0, np.asarray(label_idx, dtype=np.int32))
            onehot = np.zeros((n, num_classes), dtype=np.int64)
            onehot[np.arange(n), label_idx] = 1
            p.data = DeviceFeatures(ctx, n, ctx.feature_count(image_size))
            p.labels = onehot
            p._dataset_token = (id(p.data), n)
            p._labels_uploaded = np.ascontiguousarray(np.asarray(label_idx, dtype=np.int32))

        def local(self, path, max_image=None, image_size=64):
            self.dataset_path = path
            p = self.parent
            try:
                p._require()
                classes = os.listdir(path)
                if not classes:
                    raise ValueError(f"No class directories found in {path}")
                self._init_layers(classes, ((image_size - 2) // 2) ** 2 * len(p.filters))
                p.image_size = image_size
                images, label_idx = [], []
                for idx, cls in enumerate(classes):
                    folder = f"{path}/{cls}"
                    if not os.path.isdir(folder):
                        print(f"Warning: {folder} is not a directory, skipping...")
                        continue
                    files = [f for f in os.listdir(folder) if f.lower().endswith(_IMAGE_EXT_TRAIN)]
                    if not files:
                        print(f"Warning: No image files found in {folder}")
                        continue
                    got = _decode_files(
                        [f"{folder}/{name}" for name in files], image_size, max_image,
                        on_ok=lambda pth, n, cls=cls: print(f"\r\033[KLoading local dataset: {cls}/{os.path.basename(pth)} ({n})", end="", flush=True),
                        on_fail=lambda pth, e: print(f"\nWarning: Failed to process {pth}: {e}"))
                    images.extend(got)
                    label_idx.extend([idx] * len(got))
                print(f"\nLoaded {len(images)} images from {len(classes)} classes")
                if not images:
                    raise ValueError("No images were successfully loaded")
                self._upload(images, label_idx, len(classes), image_size)
            except Exception as e:
                raise RuntimeError(f"Failed to load local dataset from {path}: {e}")

        def hf(self, name, max_image=None, split="train", cached=True):
            try:
                from datasets import load_dataset
            except ImportError:
                raise ImportError("datasets library is required for Hugging Face datasets. Install with: pip install datasets")
            p = self.parent
            p._require()
            try:
                ds = load_dataset(name, split=split) if cached else \
                    load_dataset(name, split=split, download_mode="force_redownload")
            except Exception as e:
                raise RuntimeError(f"Failed to load dataset '{name}': {e}")
            label_col = next((c for c in ("label", "labels", "target", "class") if c in ds.column_names), None)
            if label_col is None:
                raise ValueError(f"Dataset '{name}' does not have a recognized label column. Available columns: {ds.column_names}")
            image_col = next((c for c in ("image", "img", "picture", "photo") if c in ds.column_names), None)
            if image_col is None:
                raise ValueError(f"Dataset '{name}' does not have a recognized image column. Available columns: {ds.column_names}")
            feat = ds.features[label_col]
            if hasattr(feat, "names"):
                names = feat.names
            elif hasattr(feat, "_int2str"):
                names = [feat._int2str[i] for i in range(len(feat._int2str))]
            else:
                names = sorted({s[label_col] for s in ds})
                print(f"Warning: Could not extract class names from dataset features, using: {names}")
            image_size = np.asarray(next(iter(ds))[image_col]).shape[1]   # PIL size[0] == width
            self._init_layers(names, ((image_size - 2) // 2) ** 2 * len(p.filters))
            p.image_size = image_size
            counts = {i: 0 for i in range(len(names))}
            images, label_idx = [], []
            for i, sample in enumerate(ds):
                try:
                    lab = sample[label_col]
                    if isinstance(lab, (list, tuple)):
                        lab = lab[0]
                    if max_image is not None and counts[lab] >= max_image:
                        continue
                    lname = names[lab] if lab < len(names) else str(lab)
                    print(f"\r\033[KLoading HuggingFace dataset: {name} - {lname} ({counts[lab] + 1})", end="", flush=True)
                    img = sample[image_col]
                    if hasattr(img, "convert"):
                        img = img.convert("RGB")
                    images.append(_to_u8_image(img))
                    label_idx.append(lab)
                    counts[lab] += 1
                    if max_image is not None and all(c >= max_image for c in counts.values()):
                        print(f"\nAll classes reached maximum of {max_image} images each.")
                        break
                except Exception as e:
                    print(f"\nWarning: Failed to process sample {i}: {e}")
            print(f"\nLoaded {len(images)} total images:")
            for cname, cnt in zip(names, counts.values()):
                print(f"  {cname}: {cnt} images")
            if not images:
                raise ValueError("No images were successfully loaded from the dataset")
            self._upload(images, label_idx, len(names), image_size)

        def synthetic(self, images, label_idx, classes):
            """Extension (not in the reference): load an in-memory uint8 batch [n,S,S,3] + class indices."""
            p = self.parent
            images = np.asarray(images)
            size = images.shape[1]
            self._init_layers(list(classes), ((size - 2) // 2) ** 2 * len(p.filters))
            p.image_size = size
            self._upload(list(images), np.asarray(label_idx), len(classes), size)

    # ------------------------------------------------------------------ device state plumbing
    def _keys(self):
        n = len(self.layers)
        return [f"w{i + 1}" for i in range(n)] + ["wo"], [f"b{i + 1}" for i in range(n)] + ["bo"]

    def _bind_network(self):
        ctx = self._require()
        if not hasattr(self, "weights") or not hasattr(self, "biases"):
            raise ValueError("No trained model available. Train or load a model first.")
        wk, _ = self._keys()
        d_in = self.weights[wk[0]].shape[1]
        key = (d_in, tuple(self.layers), self.weights["wo"].shape[0])
        if key != self._net_key:
            ctx.set_network(key[0], list(key[1]), key[2])
            self._net_key, self._params_on_device, self._opt_key = key, False, None
        return ctx

    def _push_params(self, force=False):
        ctx = self._bind_network()
        if self._params_on_device and not force:
            return ctx
        wk, bk = self._keys()
        for l, (w, b) in enumerate(zip(wk, bk)):
            ctx.set_param(2 * l, self.weights[w])
            ctx.set_param(2 * l + 1, self.biases[b])
        self._params_on_device = True
        return ctx

    def _pull_params(self):
        ctx = self._require()
        wk, bk = self._keys()
        for l, (w, b) in enumerate(zip(wk, bk)):
            self.weights[w] = ctx.get_param(2 * l)
            self.biases[b] = ctx.get_param(2 * l + 1)

    def sync_weights(self):
        """Call after mutating model.weights / model.biases by hand."""
        self._params_on_device = False

    def _push_optimizer(self):
        ctx = self._require()
        key = (self.optimizer, float(self.learning_rate), self.beta1, self.beta2, self.eps, self._net_key)
        if key != self._opt_key:   # (re)configures and zeroes m, v, t like adam() does (:201-211)
            ctx.set_optimizer(_lib.OPT_ADAM if self.optimizer == "adam" else _lib.OPT_SGD,
                              self.learning_rate, self.beta1, self.beta2, self.eps)
            self._opt_key = key

    def optimizer_state(self, _synced=False):
        """Adam m, v as float64 dicts (downloaded on demand; the reference keeps them in self.m/self.v).
        Data-parallel runs: a collective (every rank calls it)."""
        ctx = self._require()
        if self._world > 1 and not _synced:
            ctx.comm_sync_opt_state()
        wk, bk = self._keys()
        m, v = {}, {}
        for l, (w, b) in enumerate(zip(wk, bk)):
            m[w], v[w] = ctx.get_opt_state(0, 2 * l), ctx.get_opt_state(1, 2 * l)
            m[b], v[b] = ctx.get_opt_state(0, 2 * l + 1), ctx.get_opt_state(1, 2 * l + 1)
        return m, v

    def _push_dataset(self):
        """Make the device dataset reflect model.data / model.labels (which the user may have replaced)."""
        ctx = self._bind_network()
        labels = np.asarray(self.labels)
        idx = np.ascontiguousarray((labels.argmax(axis=1) if labels.ndim == 2 else labels).astype(np.int32))
        if isinstance(self.data, DeviceFeatures) and getattr(self, "_dataset_token", None) == (id(self.data), len(self.data)):
            # the features are still the device-resident ones; the labels are cheap to compare and may have been
            # replaced or edited in place since they were uploaded
            if len(idx) != len(self.data):
                raise ValueError(f"model.labels has {len(idx)} rows, model.data {len(self.data)}")
            if getattr(self, "_labels_uploaded", None) is None or not np.array_equal(self._labels_uploaded, idx):
                ctx.dataset_set_labels(0, idx)
                self._labels_uploaded = idx
            return ctx, len(self.data)
        feats = np.asarray(self.data, dtype=np.float64)
        if feats.ndim != 2:
            raise ValueError(f"model.data must be [N, D], got shape {feats.shape}")
        ctx.dataset_alloc(len(feats))
        ctx.dataset_set_features(0, feats)
        ctx.dataset_set_labels(0, idx)
        self._labels_uploaded = idx
        self.data = DeviceFeatures(ctx, len(feats), feats.shape[1])
        self._dataset_token = (id(self.data), len(feats))
        return ctx, len(feats)

    # ------------------------------------------------------------------ reference-protocol methods (:474-570)
    def _forward_pass(self, batch_x, activations, z_values):
        ctx = self._push_params()
        batch_x = np.asarray(batch_x, dtype=np.float64)
        activations[0] = batch_x
        return ctx.forward_features(batch_x)

    def _backward_pass(self, output, batch_y, activations, z_values, gradients_w, gradients_b):
        ctx = self._require()
        y = np.asarray(batch_y)
        if y.ndim != 2 or not np.all((y == 0) | (y == 1)) or not np.all(y.sum(axis=1) == 1):
            raise ValueError("pycnn_b200 expects one-hot label rows (what the dataset loaders produce)")
        ctx.backward_last(y.argmax(axis=1))
        wk, bk = self._keys()
        for l, (w, b) in enumerate(zip(wk, bk)):
            gradients_w[w][:] = ctx.get_grad(2 * l)
            gradients_b[b][:] = ctx.get_grad(2 * l + 1)

    def _gradient_decent(self, gradients_w, gradients_b):
        ctx = self._require()
        self._push_optimizer()
        ctx.apply_update()
        self._pull_params()
        self.t = ctx.step_count()

    # ------------------------------------------------------------------ training (:572-686)
    def train_model(self, visualize=False, early_stop=0, start_epoch=0, checkpoint=None, checkpoint_every=1):
        """pycnn/pycnn.py:572-686.  Additive keywords: `start_epoch` resumes the epoch counter after
        load_checkpoint(); `checkpoint` = path of a resumable sidecar written every `checkpoint_every` epochs."""
        if not hasattr(self, "data") or len(self.data) == 0:
            raise ValueError("No data loaded. Please load a dataset first using dataset.local() or dataset.hf()")
        ctx = self._require()
        if self._rank == 0:
            print("=" * 50)
            print("Training the model...")
        self.visualize = visualize
        self._push_params(force=not self._resumed)
        ctx, num_samples = self._push_dataset()
        self._push_optimizer()
        self._resumed = False
        plot = _LivePlot(self.epochs) if visualize else None
        best_loss, stop_counter, patience = float("inf"), 0, early_stop
        perf_log = os.environ.get("PYCNN_B200_PERF_LOG") if self._rank == 0 else None   # one JSON line per epoch
        for epoch in range(int(start_epoch), self.epochs):
            perm = np.random.permutation(num_samples)                     # :622, host stream
            t_epoch, launches0 = time.perf_counter(), ctx.launch_count()
            loss_sum, correct = ctx.train_epoch(perm, self.batch_size)    # :626-642 in one call
            if perf_log:
                dt = time.perf_counter() - t_epoch
                with open(perf_log, "a") as f:
                    f.write(json.dumps({"epoch": epoch + 1, "seconds": dt, "images_per_s": num_samples / dt, "images": num_samples,
                                        "batch_size": self.batch_size, "loss": loss_sum / max(1, num_samples),
                                        "acc": correct / max(1, num_samples), "gpu_launches": ctx.launch_count() - launches0,
                                        "world": self._world}) + "\n")
            epoch_loss = loss_sum / max(1, num_samples)
            epoch_acc = correct / max(1, num_samples)
            if early_stop:
                if epoch_loss < best_loss - 1e-8:
                    best_loss, stop_counter = epoch_loss, 0
                else:
                    stop_counter += 1
                    if stop_counter >= patience:
                        if self._rank == 0:
                            print(f"\nEarly stopping at epoch {epoch + 1}")
                        break
            if self._rank == 0:
                print(f"\rTraining model: epoch={epoch + 1}/{self.epochs}  correct={correct}/{num_samples}  "
                      f"acc={epoch_acc:.4f}  loss={epoch_loss:.6f}   ", end="")
            if plot:
                plot.update(epoch, epoch_loss, epoch_acc)
            if checkpoint and (epoch + 1) % max(1, int(checkpoint_every)) == 0:
                self.save_checkpoint(checkpoint, epoch=epoch + 1, quiet=True)     # collective; rank 0 writes
            if epoch_loss <= 1e-6 and epoch_acc >= 0.999999 and epoch < self.epochs:
                if self._rank == 0:
                    print(f"\nModel reached maximum learning on epoch {epoch}")
                break
        if self._rank == 0:
            print()
        if plot:
            plot.finish()
        self._pull_params()
        self.t = ctx.step_count()

    # ------------------------------------------------------------------ persistence (:688-735)
    def save_model(self, path="model.bin"):
        if not hasattr(self, "weights") or not hasattr(self, "biases"):
            raise ValueError("No trained model to save. Train the model first.")
        blob = {
            "weights": {k: np.asarray(v, dtype=np.float64) for k, v in self.weights.items()},
            "biases": {k: np.asarray(v, dtype=np.float64) for k, v in self.biases.items()},
            "filters": self.filters,
            "image_size": int(self.image_size),
            "classes": list(self.classes),
            "layers": self.layers,
            "trained_with_cuda": self.use_cuda,
        }
        with open(path, "wb") as f:
            pickle.dump(blob, f, protocol=4)
        print(f"Model saved in {path} with {round(os.path.getsize(path) / (1024 * 1024), 2)}MB")

    def load_model(self, model_path):
        with open(model_path, "rb") as f:
            blob = pickle.load(f)
        self.weights = {k: np.array(v, dtype=np.float64) for k, v in blob["weights"].items()}
        self.biases = {k: np.array(v, dtype=np.float64) for k, v in blob["biases"].items()}
        self.filters = blob["filters"]
        self.image_size = blob["image_size"]
        self.classes = blob["classes"]
        self.layers = blob.get("layers", [])
        self._params_on_device = False
        trained_with_cuda = blob.get("trained_with_cuda", False)
        print("Model loaded successfully!")
        print(f"Size: {round(os.path.getsize(model_path) / (1024 * 1024), 2)}MB")
        print(f"Training backend: {'CUDA' if trained_with_cuda else 'CPU'}")
        print(f"Current backend: {'CUDA' if self.use_cuda else 'CPU'}")
        print(f"Network architecture: {len(self.layers)} hidden layers: {self.layers}")
        if trained_with_cuda != self.use_cuda:
            print("Backend conversion completed automatically.")

    def torch(self, path="model.pth", dtype=None):
        """Export to the PyTorch checkpoint the reference documents (README.md:184-250)."""
        from . import torch_export
        if self.use_cuda and self._ctx is not None and self._params_on_device:
            self._pull_params()
        net = torch_export.export(self, path, dtype=dtype)
        print(f"PyTorch model exported to {path}")
        return net

    # ------------------------------------------------------------------ resumable checkpoints
    CHECKPOINT_FORMAT = "pycnn_b200.checkpoint/1"

    def save_checkpoint(self, path="checkpoint.bin", epoch=None, quiet=False):
        """Additive sidecar to model.bin (whose dict layout is untouched, :688-706): the same model dict plus what
        the reference cannot restore -- Adam m, v, t, hyper-parameters, the epoch reached and the host NumPy RNG
        state that drives np.random.permutation (:622) -- so that load_checkpoint() + train_model(start_epoch=...)
        continues the run it was taken from.  Data-parallel runs: every rank calls it (the sharded optimizer state
        of the peer-memory exchange is gathered first); rank 0 writes the file."""
        ctx = self._require()
        if self._world > 1:
            ctx.comm_sync_opt_state()
            if self._rank != 0:
                return
        if self._params_on_device:
            self._pull_params()
        m, v = self.optimizer_state(_synced=True) if self._opt_key is not None else ({}, {})
        blob = {
            "format": self.CHECKPOINT_FORMAT,
            "model": {"weights": {k: np.asarray(x, dtype=np.float64) for k, x in self.weights.items()},
                      "biases": {k: np.asarray(x, dtype=np.float64) for k, x in self.biases.items()},
                      "filters": self.filters, "image_size": int(self.image_size), "classes": list(self.classes),
                      "layers": self.layers, "trained_with_cuda": self.use_cuda},
            "optimizer": {"kind": self.optimizer, "learning_rate": float(self.learning_rate), "beta1": self.beta1,
                          "beta2": self.beta2, "eps": self.eps, "m": m, "v": v, "t": int(ctx.step_count())},
            "epoch": None if epoch is None else int(epoch), "epochs": int(self.epochs), "batch_size": int(self.batch_size),
            "numpy_rng_state": np.random.get_state(),
        }
        tmp = f"{path}.tmp"
        with open(tmp, "wb") as f:
            pickle.dump(blob, f, protocol=4)
        os.replace(tmp, path)                      # a crash mid-write never leaves a truncated checkpoint
        if not quiet:
            print(f"Checkpoint saved in {path} with {round(os.path.getsize(path) / (1024 * 1024), 2)}MB")

    def load_checkpoint(self, path="checkpoint.bin", restore_rng=True):
        """Restore a save_checkpoint() file into this model and its device context; returns the epoch it was taken
        at (pass it as train_model(start_epoch=...))."""
        with open(path, "rb") as f:
            blob = pickle.load(f)
        if not isinstance(blob, dict) or blob.get("format") != self.CHECKPOINT_FORMAT:
            raise ValueError(f"{path} is not a {self.CHECKPOINT_FORMAT} file (model.bin files load with load_model)")
        mdl, opt = blob["model"], blob["optimizer"]
        self.weights = {k: np.array(x, dtype=np.float64) for k, x in mdl["weights"].items()}
        self.biases = {k: np.array(x, dtype=np.float64) for k, x in mdl["biases"].items()}
        self.filters, self.image_size = mdl["filters"], mdl["image_size"]
        self.classes, self.layers = mdl["classes"], mdl.get("layers", [])
        self.optimizer, self.learning_rate = opt["kind"], opt["learning_rate"]
        self.beta1, self.beta2, self.eps = opt["beta1"], opt["beta2"], opt["eps"]
        self.epochs, self.batch_size = blob.get("epochs", self.epochs), blob.get("batch_size", self.batch_size)
        self._params_on_device = False
        ctx = self._push_params(force=True)
        self._opt_key = None
        self._push_optimizer()                     # zeroes m, v, t ...
        wk, bk = self._keys()
        for l, (w, b) in enumerate(zip(wk, bk)):   # ... then put the saved state back
            for which, state in ((0, opt["m"]), (1, opt["v"])):
                if w in state:
                    ctx.set_opt_state(which, 2 * l, state[w])
                    ctx.set_opt_state(which, 2 * l + 1, state[b])
        ctx.set_step_count(int(opt.get("t", 0)))
        self.t = int(opt.get("t", 0))
        if restore_rng and blob.get("numpy_rng_state") is not None:
            np.random.set_state(blob["numpy_rng_state"])
        self._resumed = True                       # train_model must not overwrite the restored device state
        return blob.get("epoch")

    # ------------------------------------------------------------------ inference (:737-773)
    def predict(self, img_path):
        if not hasattr(self, "weights") or not hasattr(self, "biases"):
            raise ValueError("No trained model available. Train or load a model first.")
        ctx = self._push_params()
        self._sync_filters(self.filters)
        cls, conf = ctx.predict_images(_to_u8_image(img_path, self.image_size))
        return self.classes[int(cls[0])], float(conf[0])

    def predict_batch(self, images):
        """Extension: batched predict on uint8 [n,S,S,3] -> (class indices, confidences).  Under torchrun (after
        parallel.attach) the images are sharded over the ranks' GPUs -- no collective in the data path -- and every
        rank returns the full result (a collective call: every rank passes the same images)."""
        ctx = self._push_params()
        self._sync_filters(self.filters)
        if self._world > 1:
            from . import parallel
            return parallel.predict_sharded(ctx, np.asarray(images), self._rank, self._world)
        return ctx.predict_images(images)


class _LivePlot:
    """visualize=True (pycnn/pycnn.py:600-619, 663-677, 684-686).  matplotlib is optional here."""

    def __init__(self, epochs):
        try:
            import matplotlib.pyplot as plt
        except Exception:
            print("Warning: matplotlib is not installed; visualize=True is ignored.")
            self.plt = None
            return
        self.plt, self.epochs = plt, epochs
        plt.ion()
        self.fig, self.ax = plt.subplots()
        self.loss_line, = self.ax.plot([], [], label="loss", linewidth=2)
        self.acc_line, = self.ax.plot([], [], label="accuracy", linewidth=2)
        self.ax.set_xlabel("Epoch")
        self.ax.set_ylabel("Value")
        self.ax.set_title("Training: loss & accuracy per epoch")
        self.ax.legend()
        self.x, self.losses, self.accs = [], [], []

    def update(self, epoch, loss, acc):
        if not self.plt:
            return
        self.x.append(epoch + 1)
        self.losses.append(float(loss))
        self.accs.append(float(acc))
        self.loss_line.set_data(self.x, self.losses)
        self.acc_line.set_data(self.x, self.accs)
        self.ax.relim()
        self.ax.autoscale_view()
        self.ax.set_xlim(1, max(self.epochs, len(self.x)))
        if epoch % max(1, self.epochs // 100) == 0 or epoch == self.epochs - 1:
            self.fig.canvas.draw()
            self.fig.canvas.flush_events()

    def finish(self):
        if self.plt:
            self.plt.ioff()
            self.plt.show()


class Evaluate:
    """pycnn/pycnn.py:779-914: batched inference + loss/accuracy/per-class tallies; prints, returns None."""

    def __init__(self, model):
        self.model = model
        if not hasattr(model, "weights") or not hasattr(model, "biases"):
            raise ValueError("The provided model is not trained. Load weights first.")

    def local(self, path, max_image=None):
        if not os.path.exists(path):
            raise ValueError(f"Test dataset path does not exist: {path}")
        m = self.model
        images, label_idx = [], []
        print("=" * 50)
        print(f"Loading local test data from: {path}")
        for idx, cls in enumerate(m.classes):
            folder = os.path.join(path, cls)
            if not os.path.isdir(folder):
                continue
            names = [f for f in os.listdir(folder) if f.lower().endswith(_IMAGE_EXT_EVAL)]
            got = _decode_files([os.path.join(folder, name) for name in names], m.image_size, max_image,
                                on_fail=lambda pth, e: print(f"\nWarning: Failed to process {pth}: {e}"))
            images.extend(got)
            label_idx.extend([idx] * len(got))
        return self._run_images(images, label_idx)

    def hf(self, dataset_name, split="test", max_image=None):
        try:
            from datasets import load_dataset
        except ImportError:
            raise ImportError("datasets library is required. Install with: pip install datasets")
        m = self.model
        print("=" * 50)
        print(f"Loading HuggingFace dataset '{dataset_name}' (split: {split})...")
        ds = load_dataset(dataset_name, split=split)
        label_col = next((c for c in ("label", "labels", "target") if c in ds.column_names), None)
        image_col = next((c for c in ("image", "img", "photo") if c in ds.column_names), None)
        if not label_col or not image_col:
            raise ValueError(f"Could not find image/label columns in {ds.column_names}")
        hf_names = ds.features[label_col].names if hasattr(ds.features[label_col], "names") else []
        counts = {c: 0 for c in m.classes}
        images, label_idx = [], []
        for sample in ds:
            lab = sample[label_col]
            lname = hf_names[lab] if hf_names else str(lab)
            if lname not in m.classes or (max_image and counts[lname] >= max_image):
                continue
            img = sample[image_col]
            if hasattr(img, "convert"):
                img = img.convert("RGB")
            images.append(_to_u8_image(img))
            label_idx.append(m.classes.index(lname))
            counts[lname] += 1
        return self._run_images(images, label_idx)

    def _run_images(self, images, label_idx):
        """Whole evaluation on the device: extractor + forward + CE/argmax per chunk of batch_size."""
        if not images:
            raise ValueError("No valid data loaded for evaluation.")
        m = self.model
        ctx = m._push_params()
        m._sync_filters(m.filters)
        n = len(images)
        bs = getattr(m, "batch_size", 32)
        if m._world > 1:       # image-sharded over the ranks' GPUs: tallies are added on the host
            from . import parallel
            loss_sum, correct, pred = parallel.evaluate_sharded(ctx, np.stack(images), np.asarray(label_idx, dtype=np.int32),
                                                                bs, m._rank, m._world)
            if m._rank != 0:
                self.last_result = {"acc": correct / n, "loss": loss_sum / n}
                return
        else:
            loss_sum, correct, pred = ctx.evaluate_images(np.stack(images), np.asarray(label_idx, dtype=np.int32), bs)
        print(f"\rEvaluating: {n}/{n}", end="", flush=True)
        self._report(loss_sum, correct, pred, np.asarray(label_idx))

    def _run_inference(self, test_data, test_labels):
        """Reference signature (:873): precomputed feature rows + one-hot labels."""
        if len(test_data) == 0:
            raise ValueError("No valid data loaded for evaluation.")
        m = self.model
        feats = np.asarray(test_data, dtype=np.float64)
        labels = np.asarray(test_labels)
        n = len(feats)
        bs = getattr(m, "batch_size", 32)
        loss_sum, preds = 0.0, []
        for s in range(0, n, bs):
            out = m._forward_pass(feats[s:s + bs], [None] * 10, [None] * 10)
            loss_sum += m._cross_entropy_batch(out, labels[s:s + bs]).sum()
            preds.append(np.argmax(out, axis=1))
            print(f"\rEvaluating: {min(s + bs, n)}/{n}", end="", flush=True)
        pred = np.concatenate(preds)
        true = labels.argmax(axis=1)
        self._report(loss_sum, int((pred == true).sum()), pred, true)

    def _report(self, loss_sum, correct, pred, true):
        m = self.model
        n = len(true)
        stats = {c: {"correct": 0, "total": 0} for c in m.classes}
        for p, t in zip(pred, true):
            name = m.classes[int(t)]
            stats[name]["total"] += 1
            if p == t:
                stats[name]["correct"] += 1
        print(f"\n\nEvaluation Results:\nAcc: {correct / n:.4f} | Loss: {loss_sum / n:.4f}")
        for name, s in stats.items():
            print(name + ":", str(s["correct"]) + "/" + str(s["total"]))
        self.last_result = {"acc": correct / n, "loss": loss_sum / n, "per_class": stats}


def __getattr__(name):   # `from pycnn_b200.pycnn import PyCNNTorchModel`, as the reference README imports it
    if name == "PyCNNTorchModel":
        from . import torch_export
        return torch_export.PyCNNTorchModel
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")
