"""The reference's COMPILED kernels (oracle/_ref) behind a restated driver loop (TEST INFRASTRUCTURE).

This is what travels to the GPU box: the .so files built by oracle/build_ref.py from the
reference's own sources -- forward_pass / backward_pass / gradient_decent / max_pooling (Cython)
and softmax / cross_entropy (optimized.cpp) -- are loaded here and driven by a restatement of the
few Python lines that call them (pycnn/pycnn.py:267-287 for the extractor, :626-642 for the batch
loop).  scipy.signal.convolve2d is the same third-party call the reference makes (:139,280-282).
bench.py uses it as ``cpu_baseline.kind == "reference"`` and for ``--impl reference``;
tests use it as a second checker next to oracle/pycnn_oracle.py.
"""
from __future__ import annotations

import ctypes
import importlib.util
import os

import numpy as np

from . import build_ref

_mods = {}
_lib = None


def available() -> bool:
    return build_ref.built()


def _load_ext(name: str):
    if name not in _mods:
        path = os.path.join(build_ref.OUT, "pycnn", "modules", f"{name}.so")
        spec = importlib.util.spec_from_file_location(name, path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        _mods[name] = mod
    return _mods[name]


def _load_lib():
    global _lib
    if _lib is None:
        lib = ctypes.CDLL(os.path.join(build_ref.OUT, "pycnn", "lib", "optimized.so"))
        dp = ctypes.POINTER(ctypes.c_double)
        lib.softmax.argtypes = [dp, dp, ctypes.c_int, ctypes.c_int]            # pycnn/pycnn.py:28-33
        lib.cross_entropy.argtypes = [dp, dp, dp, ctypes.c_int, ctypes.c_int]  # pycnn/pycnn.py:48-54
        _lib = lib
    return _lib


def softmax(x):
    """ctypes wrapper equivalent to pycnn/pycnn.py:35-46."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    out = np.empty_like(x)
    dp = ctypes.POINTER(ctypes.c_double)
    _load_lib().softmax(x.ctypes.data_as(dp), out.ctypes.data_as(dp), x.shape[0], x.shape[1])
    return out


def cross_entropy(preds, labels):
    """ctypes wrapper equivalent to pycnn/pycnn.py:56-66."""
    preds = np.ascontiguousarray(preds, dtype=np.float64)
    labels = np.ascontiguousarray(labels, dtype=np.float64)
    out = np.empty(preds.shape[0], dtype=np.float64)
    dp = ctypes.POINTER(ctypes.c_double)
    _load_lib().cross_entropy(preds.ctypes.data_as(dp), labels.ctypes.data_as(dp),
                              out.ctypes.data_as(dp), preds.shape[0], preds.shape[1])
    return out


def max_pooling(fm):
    return _load_ext("max_pooling").max_pooling(np.ascontiguousarray(fm, dtype=np.float64))


def extract_features(img, filters):
    """Restates pycnn/pycnn.py:273-287 around scipy's convolve2d and the compiled max_pooling."""
    from scipy.signal import convolve2d
    a = np.array(np.asarray(img), dtype=float) / 255.0
    r, g, b = a[..., 0], a[..., 1], a[..., 2]
    maps = []
    for filt in filters:
        filt = np.array(filt)
        conv = convolve2d(r, filt, mode="valid") + convolve2d(g, filt, mode="valid") + \
            convolve2d(b, filt, mode="valid")
        conv = np.maximum(conv, 0)
        maps.append(max_pooling(conv))
    return np.array(maps).reshape(-1)


class RefTrainer:
    """Restates the state + batch loop of train_model (pycnn/pycnn.py:586-598, 626-642) around the
    compiled forward_pass / backward_pass / gradient_decent and optimized.so::cross_entropy."""

    def __init__(self, weights, biases, layers, lr, optimizer="sgd", beta1=0.9, beta2=0.999, eps=1e-8):
        self.fwd = _load_ext("forward_pass").forward_pass
        self.bwd = _load_ext("backward_pass").backward_pass
        self.upd = _load_ext("gradient_decent").gradient_decent
        self.weights = {k: np.array(v, dtype=np.float64) for k, v in weights.items()}
        self.biases = {k: np.array(v, dtype=np.float64) for k, v in biases.items()}
        self.layers = list(layers)
        self.lr, self.optimizer = float(lr), optimizer
        self.beta1, self.beta2, self.eps, self.t = beta1, beta2, eps, 0
        self.m, self.v = {}, {}
        if optimizer == "adam":                                   # pycnn/pycnn.py:205-211
            for d in (self.weights, self.biases):
                for k in d:
                    self.m[k] = np.zeros(d[k].shape, dtype=float)
                    self.v[k] = np.zeros(d[k].shape, dtype=float)
        self.gw = {k: np.zeros(v.shape) for k, v in self.weights.items()}
        self.gb = {k: np.zeros(v.shape) for k, v in self.biases.items()}
        n = len(self.layers) + 1
        self.acts, self.zs = [None] * n, [None] * n

    def forward(self, batch_x):
        return self.fwd(np.ascontiguousarray(batch_x, dtype=np.float64), self.acts, self.zs,
                        self.weights, self.biases, self.layers)

    def step(self, batch_x, batch_y):
        out = self.forward(batch_x)
        loss_sum = float(cross_entropy(out, batch_y).sum())
        correct = int(np.sum(np.argmax(out, axis=1) == np.argmax(batch_y, axis=1)))
        self.bwd(out, batch_y, self.acts, self.zs, self.gw, self.gb, self.weights, self.layers)
        # return value discarded exactly like pycnn/pycnn.py:564-570 (so t stays 0)
        self.upd(self.weights, self.biases, self.gw, self.gb, self.lr, self.optimizer,
                 self.m, self.v, self.beta1, self.beta2, self.eps, self.t)
        return loss_sum, correct
