"""ctypes binding of libpycnn_b200.so -- the stub a PyCNN maintainer would add next to the existing
`optimized_lib = ctypes.CDLL(...)` block (pycnn/pycnn.py:18-66).  See include/pycnn_b200.h.

The library is loaded lazily; failure to load or to find a B200 raises -- there is no CPU fallback
(the reference silently falls back, pycnn/pycnn.py:127-132; the north star forbids that here).
"""
from __future__ import annotations

import ctypes
import os
import re

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpycnn_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "pycnn_b200.h")

ABI_VERSION = 2          # include/pycnn_b200.h: PYCNN_B200_ABI_VERSION
MATH_FP32_SIMT, MATH_TF32 = 0, 1
RULE_CPU, RULE_CUPY = 0, 1
OPT_SGD, OPT_ADAM = 0, 1
EXTRACT_SIMT, EXTRACT_TC = 0, 1

c_ctx = ctypes.c_void_p
_dp = ctypes.POINTER(ctypes.c_double)
_fp = ctypes.POINTER(ctypes.c_float)
_u8p = ctypes.POINTER(ctypes.c_uint8)
_i32p = ctypes.POINTER(ctypes.c_int32)
_i64p = ctypes.POINTER(ctypes.c_int64)
_ip = ctypes.POINTER(ctypes.c_int)
I, I64, D, SZ = ctypes.c_int, ctypes.c_int64, ctypes.c_double, ctypes.c_size_t

# name -> argtypes (restype is int unless listed in _RESTYPES)
_SIGNATURES = {
    "pycnn_b200_abi_version": [],
    "pycnn_b200_last_error": [],
    "pycnn_b200_device_count": [_ip],
    "pycnn_b200_create": [I, ctypes.POINTER(c_ctx)],
    "pycnn_b200_destroy": [c_ctx],
    "pycnn_b200_synchronize": [c_ctx],
    "pycnn_b200_set_math_mode": [c_ctx, I],
    "pycnn_b200_set_update_rule": [c_ctx, I],
    "pycnn_b200_set_extract_mode": [c_ctx, I],
    "pycnn_b200_device_info": [c_ctx, ctypes.c_char_p, SZ],
    "pycnn_b200_set_filters": [c_ctx, _dp, I],
    "pycnn_b200_feature_count": [c_ctx, I, _i64p],
    "pycnn_b200_extract": [c_ctx, _u8p, I64, I, _dp],
    "pycnn_b200_extract_f32": [c_ctx, _u8p, I64, I, _fp],
    "pycnn_b200_max_pool": [c_ctx, _dp, I, I, _dp],
    "pycnn_b200_softmax": [c_ctx, _dp, _dp, I, I],
    "pycnn_b200_cross_entropy": [c_ctx, _dp, _dp, _dp, I, I],
    "pycnn_b200_set_network": [c_ctx, I64, _ip, I, I],
    "pycnn_b200_num_params": [c_ctx, _ip],
    "pycnn_b200_param_shape": [c_ctx, I, _i64p, _i64p],
    "pycnn_b200_set_param": [c_ctx, I, _dp],
    "pycnn_b200_get_param": [c_ctx, I, _dp],
    "pycnn_b200_get_grad": [c_ctx, I, _dp],
    "pycnn_b200_get_opt_state": [c_ctx, I, I, _dp],
    "pycnn_b200_set_opt_state": [c_ctx, I, I, _dp],
    "pycnn_b200_set_optimizer": [c_ctx, I, D, D, D, D],
    "pycnn_b200_get_step_count": [c_ctx, _i64p],
    "pycnn_b200_trace_control": [c_ctx, ctypes.c_int, ctypes.c_int],
    "pycnn_b200_trace_read": [c_ctx, ctypes.POINTER(ctypes.c_uint64), ctypes.POINTER(ctypes.c_int32), ctypes.c_int, ctypes.POINTER(ctypes.c_int)],
    "pycnn_b200_set_step_count": [c_ctx, ctypes.c_int64],
    "pycnn_b200_dataset_alloc": [c_ctx, I64],
    "pycnn_b200_dataset_extract": [c_ctx, I64, _u8p, I64, I],
    "pycnn_b200_dataset_set_features": [c_ctx, I64, _dp, I64],
    "pycnn_b200_dataset_get_features": [c_ctx, I64, _dp, I64],
    "pycnn_b200_dataset_set_labels": [c_ctx, I64, _i32p, I64],
    "pycnn_b200_dataset_rows": [c_ctx, _i64p],
    "pycnn_b200_train_step": [c_ctx, _i64p, I, _dp, _i64p],
    "pycnn_b200_train_epoch": [c_ctx, _i64p, I64, I, _dp, _i64p],
    "pycnn_b200_train_step_images": [c_ctx, ctypes.c_void_p, ctypes.c_void_p, I, I, _dp, _i64p],
    "pycnn_b200_train_steps": [c_ctx, _i64p, I, I, _dp, _i64p],
    "pycnn_b200_train_epoch_images": [c_ctx, ctypes.c_void_p, ctypes.c_void_p, I64, _i64p, I64, I, I, _dp, _dp, _i64p],
    "pycnn_b200_forward_rows": [c_ctx, _i64p, I64, I, _dp],
    "pycnn_b200_forward_features": [c_ctx, _dp, I, _dp],
    "pycnn_b200_predict_images": [c_ctx, ctypes.c_void_p, I64, I, _i32p, _fp],
    "pycnn_b200_evaluate_rows": [c_ctx, I64, I64, I, _dp, _i64p, _i32p],
    "pycnn_b200_evaluate_images": [c_ctx, ctypes.c_void_p, _i32p, I64, I, I, _dp, _i64p, _i32p],
    "pycnn_b200_backward_last": [c_ctx, _i32p, I],
    "pycnn_b200_comm_unique_id": [ctypes.c_char_p],
    "pycnn_b200_comm_init": [c_ctx, ctypes.c_char_p, I, I],
    "pycnn_b200_comm_destroy": [c_ctx],
    "pycnn_b200_comm_ipc_export": [c_ctx, ctypes.c_char_p],
    "pycnn_b200_comm_ipc_attach": [c_ctx, ctypes.c_char_p, I, I],
    "pycnn_b200_comm_ipc_detach": [c_ctx],
    "pycnn_b200_comm_sync_opt_state": [c_ctx],
    "pycnn_b200_comm_nvls_supported": [c_ctx, ctypes.POINTER(ctypes.c_int)],
    "pycnn_b200_comm_nvls_create": [c_ctx, I, ctypes.POINTER(ctypes.c_int)],
    "pycnn_b200_comm_nvls_join": [c_ctx, I, I, I],
    "pycnn_b200_comm_nvls_bind": [c_ctx],
    "pycnn_b200_comm_nvls_detach": [c_ctx],
    "pycnn_b200_grad_buffer": [c_ctx, ctypes.POINTER(ctypes.c_void_p), _i64p],
    "pycnn_b200_forward_backward": [c_ctx, _i64p, I],
    "pycnn_b200_apply_update": [c_ctx],
    "pycnn_b200_read_step_stats": [c_ctx, _dp, _i64p, I],
    "pycnn_b200_event_record": [c_ctx, I],
    "pycnn_b200_event_elapsed_ms": [c_ctx, I, I, _fp],
    "pycnn_b200_set_kernel_timing": [c_ctx, I],
    "pycnn_b200_kernel_class_count": [],
    "pycnn_b200_kernel_class_name": [I],
    "pycnn_b200_kernel_stats": [c_ctx, I, _i64p, _dp, I],
    "pycnn_b200_launch_count": [c_ctx, _i64p],
    "pycnn_b200_flush_l2": [c_ctx],
    "pycnn_b200_gemm_probe": [c_ctx, _fp, _fp, _fp, I, I, I, I, I, I, _fp],
}
_RESTYPES = {"pycnn_b200_last_error": ctypes.c_char_p, "pycnn_b200_kernel_class_name": ctypes.c_char_p}

_lib = None


class BackendError(RuntimeError):
    """Raised for every non-zero return code of the C ABI (message = pycnn_b200_last_error())."""


def header_symbols():
    """Function names declared in include/pycnn_b200.h (used by the symbol-export test)."""
    with open(HEADER_PATH) as f:
        text = f.read()
    return sorted(set(re.findall(r"\b(pycnn_b200_[a-z0-9_]+)\s*\(", text)))


def load():
    """dlopen libpycnn_b200.so and attach the signatures.  Raises if the library is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise BackendError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C pycnn_b200/csrc`.  pycnn_b200 has no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, argtypes in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = _RESTYPES.get(name, ctypes.c_int)
    if lib.pycnn_b200_abi_version() != ABI_VERSION:
        raise BackendError(f"{LIB_PATH} has ABI v{lib.pycnn_b200_abi_version()}, this binding needs v{ABI_VERSION}: "
                           "rebuild with `make -C pycnn_b200/csrc`")
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        msg = load().pycnn_b200_last_error()
        raise BackendError(msg.decode("utf-8", "replace") if msg else f"pycnn_b200 error {rc}")


def ptr(arr, ctype):
    return arr.ctypes.data_as(ctypes.POINTER(ctype))


class Context:
    """Thin OO wrapper over the opaque pycnn_b200_ctx handle.  All arrays are numpy, host-side."""

    def __init__(self, device: int = 0):
        self.lib = load()
        self.handle = c_ctx()
        check(self.lib.pycnn_b200_create(int(device), ctypes.byref(self.handle)))
        self.device = int(device)
        self.num_filters = 0
        self.shape = None  # (D, layers, C)

    # -- lifecycle ---------------------------------------------------------------------------
    def close(self):
        if getattr(self, "handle", None) and self.handle.value:
            self.lib.pycnn_b200_destroy(self.handle)
            self.handle = c_ctx()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def info(self) -> str:
        buf = ctypes.create_string_buffer(256)
        check(self.lib.pycnn_b200_device_info(self.handle, buf, 256))
        return buf.value.decode()

    def synchronize(self):
        check(self.lib.pycnn_b200_synchronize(self.handle))

    def set_math_mode(self, mode):
        check(self.lib.pycnn_b200_set_math_mode(self.handle, int(mode)))

    def set_update_rule(self, rule):
        check(self.lib.pycnn_b200_set_update_rule(self.handle, int(rule)))

    def set_extract_mode(self, mode):
        check(self.lib.pycnn_b200_set_extract_mode(self.handle, int(mode)))

    # -- extractor ---------------------------------------------------------------------------
    def set_filters(self, filters):
        taps = np.ascontiguousarray(np.asarray(filters, dtype=np.float64))
        if taps.ndim != 3 or taps.shape[1:] != (3, 3):
            raise ValueError(f"filters must be [F][3][3], got shape {taps.shape}")
        check(self.lib.pycnn_b200_set_filters(self.handle, ptr(taps, ctypes.c_double), taps.shape[0]))
        self.num_filters = taps.shape[0]

    def feature_count(self, image_size: int) -> int:
        out = ctypes.c_int64()
        check(self.lib.pycnn_b200_feature_count(self.handle, int(image_size), ctypes.byref(out)))
        return out.value

    @staticmethod
    def _images(images):
        a = np.asarray(images)
        if a.dtype != np.uint8:
            raise TypeError(f"images must be uint8 HWC, got {a.dtype}")
        if a.ndim == 3:
            a = a[None]
        if a.ndim != 4 or a.shape[3] != 3 or a.shape[1] != a.shape[2]:
            raise ValueError(f"images must be [n, S, S, 3], got {a.shape}")
        return np.ascontiguousarray(a)

    def extract(self, images, dtype=np.float64):
        a = self._images(images)
        n, s = a.shape[0], a.shape[1]
        d = self.feature_count(s)
        out = np.empty((n, d), dtype=dtype)
        if dtype == np.float64:
            check(self.lib.pycnn_b200_extract(self.handle, ptr(a, ctypes.c_uint8), n, s, ptr(out, ctypes.c_double)))
        elif dtype == np.float32:
            check(self.lib.pycnn_b200_extract_f32(self.handle, ptr(a, ctypes.c_uint8), n, s, ptr(out, ctypes.c_float)))
        else:
            raise TypeError("dtype must be float64 or float32")
        return out

    def max_pool(self, fm):
        fm = np.ascontiguousarray(fm, dtype=np.float64)
        h, w = fm.shape
        out = np.empty((h // 2, w // 2), dtype=np.float64)
        check(self.lib.pycnn_b200_max_pool(self.handle, ptr(fm, ctypes.c_double), h, w, ptr(out, ctypes.c_double)))
        return out

    def softmax(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        out = np.empty_like(x)
        check(self.lib.pycnn_b200_softmax(self.handle, ptr(x, ctypes.c_double), ptr(out, ctypes.c_double), *x.shape))
        return out

    def cross_entropy(self, preds, labels):
        preds = np.ascontiguousarray(preds, dtype=np.float64)
        labels = np.ascontiguousarray(labels, dtype=np.float64)
        out = np.empty(preds.shape[0], dtype=np.float64)
        check(self.lib.pycnn_b200_cross_entropy(self.handle, ptr(preds, ctypes.c_double), ptr(labels, ctypes.c_double),
                                                ptr(out, ctypes.c_double), *preds.shape))
        return out

    # -- network -----------------------------------------------------------------------------
    def set_network(self, input_size, layers, num_classes):
        arr = (ctypes.c_int * max(1, len(layers)))(*[int(x) for x in layers])
        check(self.lib.pycnn_b200_set_network(self.handle, int(input_size), arr, len(layers), int(num_classes)))
        self.shape = (int(input_size), [int(x) for x in layers], int(num_classes))

    def num_params(self) -> int:
        n = ctypes.c_int()
        check(self.lib.pycnn_b200_num_params(self.handle, ctypes.byref(n)))
        return n.value

    def param_shape(self, index):
        r, c = ctypes.c_int64(), ctypes.c_int64()
        check(self.lib.pycnn_b200_param_shape(self.handle, index, ctypes.byref(r), ctypes.byref(c)))
        return (r.value, c.value) if index % 2 == 0 else (c.value,)

    def set_param(self, index, values):
        v = np.ascontiguousarray(values, dtype=np.float64)
        if v.shape != self.param_shape(index):
            raise ValueError(f"param {index}: expected shape {self.param_shape(index)}, got {v.shape}")
        check(self.lib.pycnn_b200_set_param(self.handle, index, ptr(v, ctypes.c_double)))

    def _get(self, fn, index, *pre):
        out = np.empty(self.param_shape(index), dtype=np.float64)
        check(fn(self.handle, *pre, index, ptr(out, ctypes.c_double)))
        return out

    def get_param(self, index):
        return self._get(self.lib.pycnn_b200_get_param, index)

    def get_grad(self, index):
        return self._get(self.lib.pycnn_b200_get_grad, index)

    def get_opt_state(self, which, index):
        return self._get(self.lib.pycnn_b200_get_opt_state, index, which)

    def set_opt_state(self, which, index, values):
        v = np.ascontiguousarray(values, dtype=np.float64)
        check(self.lib.pycnn_b200_set_opt_state(self.handle, which, index, ptr(v, ctypes.c_double)))

    def set_optimizer(self, kind, lr, beta1=0.9, beta2=0.999, eps=1e-8):
        check(self.lib.pycnn_b200_set_optimizer(self.handle, int(kind), float(lr), float(beta1), float(beta2), float(eps)))

    def step_count(self) -> int:
        t = ctypes.c_int64()
        check(self.lib.pycnn_b200_get_step_count(self.handle, ctypes.byref(t)))
        return t.value

    # -- in-kernel timeline (diagnostics) ----------------------------------------------------------
    def trace_enable(self, capacity=4096):
        check(self.lib.pycnn_b200_trace_control(self.handle, 0, int(capacity)))
        self._trace_cap = int(capacity)

    def trace_reset(self):
        check(self.lib.pycnn_b200_trace_control(self.handle, 1, 0))

    def trace_disable(self):
        check(self.lib.pycnn_b200_trace_control(self.handle, 2, 0))

    def trace_read(self, marks=False):
        """-> list of (kernel class name, stream id, start_ns, end_ns) in launch order; untouched slots are skipped.
        marks=True appends the three (min_ns, max_ns) phase-mark pairs (None when the kernel did not set them)."""
        cap = self._trace_cap
        times = np.zeros(8 * cap, np.uint64)
        kinds = np.zeros(cap, np.int32)
        n = ctypes.c_int()
        check(self.lib.pycnn_b200_trace_read(self.handle, ptr(times, ctypes.c_uint64), ptr(kinds, ctypes.c_int32), cap, ctypes.byref(n)))
        names = self.kernel_class_names()
        out = []
        for i in range(n.value):
            t0, t1 = int(times[8 * i]), int(times[8 * i + 1])
            if t1 == 0:
                continue
            row = (names[int(kinds[i]) % 100], int(kinds[i]) // 100, t0, t1)
            if marks:
                row += tuple((int(times[8 * i + 2 * k]), int(times[8 * i + 2 * k + 1])) if int(times[8 * i + 2 * k + 1]) else None
                             for k in (1, 2, 3))
            out.append(row)
        return out

    def set_step_count(self, t: int):
        check(self.lib.pycnn_b200_set_step_count(self.handle, int(t)))

    # -- dataset -----------------------------------------------------------------------------
    def dataset_alloc(self, rows):
        check(self.lib.pycnn_b200_dataset_alloc(self.handle, int(rows)))

    def dataset_extract(self, row0, images):
        a = self._images(images)
        check(self.lib.pycnn_b200_dataset_extract(self.handle, int(row0), ptr(a, ctypes.c_uint8), a.shape[0], a.shape[1]))

    def dataset_set_features(self, row0, feats):
        f = np.ascontiguousarray(feats, dtype=np.float64)
        if f.ndim != 2 or f.shape[1] != self.shape[0]:
            raise ValueError(f"features must be [n, {self.shape[0]}], got {f.shape}")
        check(self.lib.pycnn_b200_dataset_set_features(self.handle, int(row0), ptr(f, ctypes.c_double), f.shape[0]))

    def dataset_get_features(self, row0, n):
        out = np.empty((int(n), self.shape[0]), dtype=np.float64)
        check(self.lib.pycnn_b200_dataset_get_features(self.handle, int(row0), ptr(out, ctypes.c_double), int(n)))
        return out

    def dataset_set_labels(self, row0, labels):
        l = np.ascontiguousarray(labels, dtype=np.int32)
        check(self.lib.pycnn_b200_dataset_set_labels(self.handle, int(row0), ptr(l, ctypes.c_int32), l.shape[0]))

    def dataset_rows(self) -> int:
        r = ctypes.c_int64()
        check(self.lib.pycnn_b200_dataset_rows(self.handle, ctypes.byref(r)))
        return r.value

    # -- training ----------------------------------------------------------------------------
    def train_step(self, idx, want_stats=True):
        idx = np.ascontiguousarray(idx, dtype=np.int64)
        loss, corr = ctypes.c_double(), ctypes.c_int64()
        check(self.lib.pycnn_b200_train_step(self.handle, ptr(idx, ctypes.c_int64), idx.shape[0],
                                             ctypes.byref(loss) if want_stats else None,
                                             ctypes.byref(corr) if want_stats else None))
        return (loss.value, corr.value) if want_stats else None

    def train_steps(self, idx, bs, steps, want_stats=True):
        idx = np.ascontiguousarray(idx, dtype=np.int64)
        assert idx.size >= bs * steps
        loss, corr = ctypes.c_double(), ctypes.c_int64()
        check(self.lib.pycnn_b200_train_steps(self.handle, ptr(idx, ctypes.c_int64), int(bs), int(steps),
                                              ctypes.byref(loss) if want_stats else None,
                                              ctypes.byref(corr) if want_stats else None))
        return (loss.value, corr.value) if want_stats else None

    def train_epoch(self, perm, batch_size):
        perm = np.ascontiguousarray(perm, dtype=np.int64)
        loss, corr = ctypes.c_double(), ctypes.c_int64()
        check(self.lib.pycnn_b200_train_epoch(self.handle, ptr(perm, ctypes.c_int64), perm.shape[0], int(batch_size),
                                              ctypes.byref(loss), ctypes.byref(corr)))
        return loss.value, corr.value

    def train_step_images(self, images, labels, want_stats=True, images_ptr=None, labels_ptr=None, n=None, size=None):
        """images/labels: numpy arrays, or raw host pointers (pinned buffers) via images_ptr/labels_ptr."""
        if images_ptr is None:
            a = self._images(images)
            l = np.ascontiguousarray(labels, dtype=np.int32)
            images_ptr, labels_ptr, n, size = a.ctypes.data, l.ctypes.data, a.shape[0], a.shape[1]
        loss, corr = ctypes.c_double(), ctypes.c_int64()
        check(self.lib.pycnn_b200_train_step_images(self.handle, ctypes.c_void_p(images_ptr), ctypes.c_void_p(labels_ptr),
                                                    int(n), int(size),
                                                    ctypes.byref(loss) if want_stats else None,
                                                    ctypes.byref(corr) if want_stats else None))
        return (loss.value, corr.value) if want_stats else None

    def train_epoch_images(self, images_ptr, labels_ptr, n, size, batch_size, perm=None, want_step_stats=False,
                           num_images=None):
        """Streaming epoch from host memory (raw pointers so pinned buffers pass through untouched).  The buffers
        hold `num_images` rows (default n); perm entries are validated against it before anything is launched."""
        steps = -(-int(n) // int(batch_size))
        st = np.empty((steps, 2), dtype=np.float64) if want_step_stats else None
        if perm is not None:
            perm = np.ascontiguousarray(perm, dtype=np.int64)
            assert perm.shape[0] == n
        loss, corr = ctypes.c_double(), ctypes.c_int64()
        check(self.lib.pycnn_b200_train_epoch_images(
            self.handle, ctypes.c_void_p(images_ptr), ctypes.c_void_p(labels_ptr), int(n if num_images is None else num_images),
            ptr(perm, ctypes.c_int64) if perm is not None else None, int(n), int(size), int(batch_size),
            ptr(st, ctypes.c_double) if st is not None else None, ctypes.byref(loss), ctypes.byref(corr)))
        return (loss.value, corr.value, st) if want_step_stats else (loss.value, corr.value)

    def forward_backward(self, idx):
        idx = np.ascontiguousarray(idx, dtype=np.int64)
        check(self.lib.pycnn_b200_forward_backward(self.handle, ptr(idx, ctypes.c_int64) if idx.size else None, idx.shape[0]))

    def backward_last(self, labels):
        l = np.ascontiguousarray(labels, dtype=np.int32)
        check(self.lib.pycnn_b200_backward_last(self.handle, ptr(l, ctypes.c_int32), l.shape[0]))

    def apply_update(self):
        check(self.lib.pycnn_b200_apply_update(self.handle))

    def read_step_stats(self, reset=True):
        loss, corr = ctypes.c_double(), ctypes.c_int64()
        check(self.lib.pycnn_b200_read_step_stats(self.handle, ctypes.byref(loss), ctypes.byref(corr), int(reset)))
        return loss.value, corr.value

    # -- inference ---------------------------------------------------------------------------
    def forward_rows(self, idx=None, row0=0, bs=None):
        c = self.shape[2]
        if idx is not None:
            idx = np.ascontiguousarray(idx, dtype=np.int64)
            bs = idx.shape[0]
        out = np.empty((int(bs), c), dtype=np.float64)
        check(self.lib.pycnn_b200_forward_rows(self.handle, ptr(idx, ctypes.c_int64) if idx is not None else None,
                                               int(row0), int(bs), ptr(out, ctypes.c_double)))
        return out

    def forward_features(self, feats):
        f = np.ascontiguousarray(feats, dtype=np.float64)
        if f.ndim != 2 or f.shape[1] != self.shape[0]:
            raise ValueError(f"features must be [bs, {self.shape[0]}], got {f.shape}")
        out = np.empty((f.shape[0], self.shape[2]), dtype=np.float64)
        check(self.lib.pycnn_b200_forward_features(self.handle, ptr(f, ctypes.c_double), f.shape[0], ptr(out, ctypes.c_double)))
        return out

    def predict_images(self, images):
        a = self._images(images)
        cls = np.empty(a.shape[0], dtype=np.int32)
        conf = np.empty(a.shape[0], dtype=np.float32)
        check(self.lib.pycnn_b200_predict_images(self.handle, ctypes.c_void_p(a.ctypes.data), a.shape[0], a.shape[1],
                                                 ptr(cls, ctypes.c_int32), ptr(conf, ctypes.c_float)))
        return cls, conf

    def evaluate_rows(self, row0, n, batch_size, want_pred=True):
        loss, corr = ctypes.c_double(), ctypes.c_int64()
        pred = np.empty(int(n), dtype=np.int32) if want_pred else None
        check(self.lib.pycnn_b200_evaluate_rows(self.handle, int(row0), int(n), int(batch_size), ctypes.byref(loss),
                                                ctypes.byref(corr), ptr(pred, ctypes.c_int32) if want_pred else None))
        return loss.value, corr.value, pred

    def evaluate_images(self, images, labels, batch_size, want_pred=True):
        a = self._images(images)
        l = np.ascontiguousarray(labels, dtype=np.int32)
        loss, corr = ctypes.c_double(), ctypes.c_int64()
        pred = np.empty(a.shape[0], dtype=np.int32) if want_pred else None
        check(self.lib.pycnn_b200_evaluate_images(self.handle, ctypes.c_void_p(a.ctypes.data), ptr(l, ctypes.c_int32),
                                                  a.shape[0], a.shape[1], int(batch_size), ctypes.byref(loss),
                                                  ctypes.byref(corr), ptr(pred, ctypes.c_int32) if want_pred else None))
        return loss.value, corr.value, pred

    # -- data parallel -----------------------------------------------------------------------
    def comm_init(self, unique_id: bytes, rank: int, world: int):
        assert len(unique_id) == 128
        check(self.lib.pycnn_b200_comm_init(self.handle, unique_id, int(rank), int(world)))

    def comm_destroy(self):
        check(self.lib.pycnn_b200_comm_destroy(self.handle))

    def comm_ipc_export(self) -> bytes:
        buf = ctypes.create_string_buffer(192)
        check(self.lib.pycnn_b200_comm_ipc_export(self.handle, buf))
        return buf.raw

    def comm_ipc_attach(self, blobs: bytes, rank: int, world: int):
        assert len(blobs) == 192 * world
        check(self.lib.pycnn_b200_comm_ipc_attach(self.handle, blobs, int(rank), int(world)))

    def comm_ipc_detach(self):
        check(self.lib.pycnn_b200_comm_ipc_detach(self.handle))

    def comm_nvls_supported(self) -> bool:
        out = ctypes.c_int(0)
        check(self.lib.pycnn_b200_comm_nvls_supported(self.handle, ctypes.byref(out)))
        return bool(out.value)

    def comm_nvls_create(self, world: int) -> int:
        """Rank 0: create the multicast object; returns a file descriptor for the other ranks (close it afterwards)."""
        fd = ctypes.c_int(-1)
        check(self.lib.pycnn_b200_comm_nvls_create(self.handle, int(world), ctypes.byref(fd)))
        return fd.value

    def comm_nvls_join(self, fd: int, rank: int, world: int):
        check(self.lib.pycnn_b200_comm_nvls_join(self.handle, int(fd), int(rank), int(world)))

    def comm_nvls_bind(self):
        check(self.lib.pycnn_b200_comm_nvls_bind(self.handle))

    def comm_nvls_detach(self):
        check(self.lib.pycnn_b200_comm_nvls_detach(self.handle))

    def comm_sync_opt_state(self):
        """Collective: assemble the sharded Adam m, v of the peer-memory mode on every rank (no-op otherwise)."""
        check(self.lib.pycnn_b200_comm_sync_opt_state(self.handle))

    def grad_buffer(self):
        p, n = ctypes.c_void_p(), ctypes.c_int64()
        check(self.lib.pycnn_b200_grad_buffer(self.handle, ctypes.byref(p), ctypes.byref(n)))
        return p.value, n.value

    # -- measurement -------------------------------------------------------------------------
    def event_record(self, slot):
        check(self.lib.pycnn_b200_event_record(self.handle, int(slot)))

    def event_elapsed_ms(self, a, b) -> float:
        ms = ctypes.c_float()
        check(self.lib.pycnn_b200_event_elapsed_ms(self.handle, int(a), int(b), ctypes.byref(ms)))
        return ms.value

    def set_kernel_timing(self, on: bool):
        check(self.lib.pycnn_b200_set_kernel_timing(self.handle, int(bool(on))))

    def kernel_class_names(self):
        return [self.lib.pycnn_b200_kernel_class_name(k).decode() for k in range(self.lib.pycnn_b200_kernel_class_count())]

    def kernel_stats(self, reset=False):
        out = {}
        for k in range(self.lib.pycnn_b200_kernel_class_count()):
            n, ms = ctypes.c_int64(), ctypes.c_double()
            check(self.lib.pycnn_b200_kernel_stats(self.handle, k, ctypes.byref(n), ctypes.byref(ms), int(reset)))
            out[self.lib.pycnn_b200_kernel_class_name(k).decode()] = (n.value, ms.value)
        return out

    def launch_count(self) -> int:
        n = ctypes.c_int64()
        check(self.lib.pycnn_b200_launch_count(self.handle, ctypes.byref(n)))
        return n.value

    def flush_l2(self):
        check(self.lib.pycnn_b200_flush_l2(self.handle))

    def gemm_probe(self, A, B, a_kmajor=True, b_kmajor=True, iters=1):
        A = np.ascontiguousarray(A, dtype=np.float32)
        B = np.ascontiguousarray(B, dtype=np.float32)
        M, K = A.shape if a_kmajor else A.shape[::-1]
        N, K2 = B.shape if b_kmajor else B.shape[::-1]
        assert K == K2
        C = np.empty((M, N), dtype=np.float32)
        ms = ctypes.c_float()
        check(self.lib.pycnn_b200_gemm_probe(self.handle, ptr(A, ctypes.c_float), ptr(B, ctypes.c_float),
                                             ptr(C, ctypes.c_float), M, N, K, int(a_kmajor), int(b_kmajor), int(iters),
                                             ctypes.byref(ms)))
        return C, ms.value


def comm_unique_id() -> bytes:
    buf = ctypes.create_string_buffer(128)
    check(load().pycnn_b200_comm_unique_id(buf))
    return buf.raw


def device_count() -> int:
    n = ctypes.c_int()
    check(load().pycnn_b200_device_count(ctypes.byref(n)))
    return n.value
