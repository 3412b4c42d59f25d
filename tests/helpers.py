"""Shared test helpers: seeded synthetic inputs and the parity metric."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def synth_images(n, s, seed=0):
    return np.random.default_rng(seed).integers(0, 256, (n, s, s, 3), dtype=np.uint8)


def synth_labels(n, c, seed=1):
    return np.random.default_rng(seed).integers(0, c, n)


def custom_filters(F=8):
    return np.round(np.random.default_rng(2).uniform(-2, 2, (F, 3, 3)), 3).tolist()


def rel_err(a, b):
    """max|a-b| / max|b| -- the norm-wise 'relative' error the north star's 1e-5 / 1e-2 refer to
    ('Tolerance metric')."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def rel_fro(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))
