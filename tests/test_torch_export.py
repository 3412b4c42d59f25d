"""the PyTorch export the reference documents (README.md:184-250) -- checkpoint keys, module signature
and numerics against the oracle.  CPU only."""
import contextlib
import io

import numpy as np
import pytest

from oracle import pycnn_oracle as O

torch = pytest.importorskip("torch")


def _model(layers=(24, 12), classes=("a", "b", "c"), S=32, filters=None):
    from pycnn_b200 import PyCNN
    m = PyCNN()
    if filters is not None:
        m.filters = filters
    m.layers, m.classes, m.image_size = list(layers), list(classes), S
    D = O.feature_count(S, len(m.filters))
    w, b = O.init_layers(D, list(layers), len(classes), seed=42)
    m.weights, m.biases = w, b
    return m


@pytest.mark.parametrize("dtype,tol", [(None, 2e-5), ("float64", 1e-12)])
def test_export_matches_oracle(tmp_path, dtype, tol):
    custom = np.round(np.random.default_rng(2).uniform(-2, 2, (5, 3, 3)), 3).tolist()     # asymmetric taps: the flip matters
    for filters in (None, custom):
        m = _model(filters=filters)
        path = str(tmp_path / "model.pth")
        with contextlib.redirect_stdout(io.StringIO()) as out:
            m.torch(path, dtype=getattr(torch, dtype) if dtype else None)
        assert "exported" in out.getvalue()
        ckpt = torch.load(path, map_location="cpu")
        assert {"layers", "num_classes", "filters", "image_size", "classes", "model_state_dict"} <= set(ckpt)
        from pycnn_b200.pycnn import PyCNNTorchModel                       # the import path the README shows
        net = PyCNNTorchModel(ckpt["layers"], ckpt["num_classes"], ckpt["filters"], ckpt["image_size"],
                              **({"dtype": getattr(torch, dtype)} if dtype else {}))
        net.load_state_dict(ckpt["model_state_dict"])
        net.eval()
        imgs = np.random.default_rng(0).integers(0, 256, (6, 32, 32, 3), dtype=np.uint8)
        x = torch.from_numpy(imgs.astype(np.float64 if dtype else np.float32) / 255.0).permute(0, 3, 1, 2)
        with torch.no_grad():
            feats = net.features(x).numpy()
            probs = net(x).numpy()
        want_feats = O.extract_batch(imgs, m.filters)
        assert np.abs(feats - want_feats).max() <= tol * max(1.0, np.abs(want_feats).max())
        want, _, _ = O.forward(want_feats, m.weights, m.biases, m.layers)
        assert np.abs(probs - want).max() <= tol
        assert (probs.argmax(1) == want.argmax(1)).all()
        np.testing.assert_allclose(probs.sum(1), 1.0, atol=1e-5)


def test_export_needs_a_model():
    from pycnn_b200 import PyCNN
    with pytest.raises(ValueError):
        PyCNN().torch("unused.pth")
