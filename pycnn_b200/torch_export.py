"""`pycnn.torch("model.pth")` / `PyCNNTorchModel` -- the PyTorch export the reference documents
(/root/reference/README.md:184-250) but does not ship.

The exported module is the whole inference path as plain torch ops, bit-for-bit the reference's semantics:
    x [B,3,S,S] in [0,1] (already /255, as in the documented usage)
      -> channel sum -> true convolution with every 3x3 tap bank entry (pycnn/pycnn.py:276-283; F.conv2d is a
         correlation, so the stored kernel is the 180-degree flip) -> ReLU -> 2x2/s2 max-pool, floor crop (:251-265)
      -> flatten filter-major (:285-287) -> Linear+ReLU per hidden layer -> Linear -> softmax (forward_pass.pyx:28-41)
Checkpoint keys are the documented ones: layers, num_classes, filters, image_size, classes, model_state_dict.
torch is imported lazily: the training path never needs it.
"""
from __future__ import annotations

import numpy as np


def _torch():
    import torch
    return torch


def build_module_class():
    torch = _torch()
    nn, F = torch.nn, torch.nn.functional

    class PyCNNTorchModel(nn.Module):
        def __init__(self, layers, num_classes, filters, image_size, dtype=None):
            super().__init__()
            dtype = dtype or torch.float32
            self.layers_spec, self.num_classes, self.image_size = list(layers), int(num_classes), int(image_size)
            taps = np.asarray(filters, dtype=np.float64).reshape(-1, 3, 3)
            flipped = np.ascontiguousarray(taps[:, ::-1, ::-1])             # convolve2d flips, conv2d does not
            self.register_buffer("conv_kernel", torch.tensor(flipped, dtype=dtype).unsqueeze(1), persistent=False)
            pooled = (self.image_size - 2) // 2
            prev = pooled * pooled * taps.shape[0]
            self.hidden = nn.ModuleList()
            for size in self.layers_spec:
                self.hidden.append(nn.Linear(prev, size, dtype=dtype))
                prev = size
            self.out = nn.Linear(prev, self.num_classes, dtype=dtype)

        def features(self, x):
            """[B,3,S,S] in [0,1] -> [B, F*P*P] filter-major, what _preprocess_image returns per image."""
            s = x.to(self.conv_kernel.dtype).sum(dim=1, keepdim=True)          # conv is linear: sum channels first
            fm = F.relu(F.conv2d(s, self.conv_kernel))
            return F.max_pool2d(fm, 2).flatten(1)

        def forward(self, x):
            a = self.features(x)
            for lin in self.hidden:
                a = F.relu(lin(a))
            return F.softmax(self.out(a), dim=1)

    return PyCNNTorchModel


_CLASS = None


def __getattr__(name):            # `from pycnn_b200.torch_export import PyCNNTorchModel` without importing torch eagerly
    global _CLASS
    if name == "PyCNNTorchModel":
        if _CLASS is None:
            _CLASS = build_module_class()
        return _CLASS
    raise AttributeError(name)


def export(model, path="model.pth", dtype=None):
    """Write the documented checkpoint for a trained / loaded PyCNN-like object (weights, biases, filters, ...)."""
    torch = _torch()
    if not hasattr(model, "weights") or not hasattr(model, "biases"):
        raise ValueError("No trained model to export. Train or load a model first.")
    cls = __getattr__("PyCNNTorchModel")
    net = cls(model.layers, len(model.classes), model.filters, model.image_size, dtype=dtype)
    dt = net.out.weight.dtype
    with torch.no_grad():
        for i, lin in enumerate(net.hidden):
            lin.weight.copy_(torch.tensor(np.asarray(model.weights[f"w{i + 1}"]), dtype=dt))
            lin.bias.copy_(torch.tensor(np.asarray(model.biases[f"b{i + 1}"]), dtype=dt))
        net.out.weight.copy_(torch.tensor(np.asarray(model.weights["wo"]), dtype=dt))
        net.out.bias.copy_(torch.tensor(np.asarray(model.biases["bo"]), dtype=dt))
    torch.save({"layers": list(model.layers), "num_classes": len(model.classes),
                "filters": [[list(map(float, row)) for row in k] for k in np.asarray(model.filters, dtype=float)],
                "image_size": int(model.image_size), "classes": list(model.classes),
                "model_state_dict": net.state_dict()}, path)
    return net
