"""GPU diagnostics (not a test): layout variants of the tcgen05 GEMM, TC extractor parity and timings."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import pycnn_oracle as O
from pycnn_b200 import _lib as lib
from tests.helpers import custom_filters, rel_err, synth_images


def gemm_diag(ctx):
    ctx.set_math_mode(lib.MATH_TF32)
    for variant in ("0", "1"):
        os.environ["PYCNN_B200_MN_VARIANT"] = variant
        for (M, N, K) in [(128, 256, 64), (128, 64, 8), (77, 33, 45), (300, 513, 96), (256, 4275, 512)]:
            rng = np.random.default_rng(1)
            A = rng.normal(size=(M, K)).astype(np.float32)
            B = rng.normal(size=(N, K)).astype(np.float32)
            want = A.astype(np.float64) @ B.astype(np.float64).T
            for (ak, bk) in [(1, 1), (1, 0), (0, 0)]:
                try:
                    C, ms = ctx.gemm_probe(A if ak else np.ascontiguousarray(A.T), B if bk else np.ascontiguousarray(B.T),
                                           a_kmajor=bool(ak), b_kmajor=bool(bk))
                    print(f"gemm variant={variant} MNK={(M, N, K)} layout={(ak, bk)} rel_err={rel_err(C, want):.3e}", flush=True)
                except Exception as e:
                    print(f"gemm variant={variant} MNK={(M, N, K)} layout={(ak, bk)} ERROR {e}", flush=True)
    os.environ["PYCNN_B200_MN_VARIANT"] = "0"


def extract_diag(ctx):
    for filt_name, filt in (("default", O.DEFAULT_FILTERS), ("custom", custom_filters())):
        ctx.set_filters(filt)
        for S, n in ((32, 5), (33, 3), (64, 2), (7, 40), (4, 300), (224, 1), (21, 37)):
            imgs = synth_images(n, S, seed=S)
            want = O.extract_batch(imgs, filt)
            for mode, mname in ((lib.EXTRACT_SIMT, "simt"), (lib.EXTRACT_TC, "tc")):
                ctx.set_extract_mode(mode)
                try:
                    got = ctx.extract(imgs)
                    print(f"extract {filt_name} S={S} n={n} {mname}: rel_err={rel_err(got, want):.3e}", flush=True)
                except Exception as e:
                    print(f"extract {filt_name} S={S} n={n} {mname}: ERROR {e}", flush=True)


def extract_timing(ctx):
    ctx.set_filters(O.DEFAULT_FILTERS)
    ctx.set_network(4275, [8], 2)
    n = 65536
    imgs = synth_images(n, 32, seed=0)
    ctx.dataset_alloc(n)
    for mode, mname in ((lib.EXTRACT_SIMT, "simt"), (lib.EXTRACT_TC, "tc")):
        ctx.set_extract_mode(mode)
        try:
            ctx.dataset_extract(0, imgs)
            ctx.set_kernel_timing(True)
            ctx.kernel_stats(reset=True)
            for _ in range(5):
                ctx.dataset_extract(0, imgs)
            st = ctx.kernel_stats(reset=True)["extract_conv_relu_pool"]
            ctx.set_kernel_timing(False)
            ms = st[1] / st[0]
            print(f"extract timing {mname}: {ms:.4f} ms / {n} imgs = {n / ms / 1e3:.1f} M img/s, "
                  f"{n * 20172 / ms / 1e6:.0f} GB/s algorithmic", flush=True)
        except Exception as e:
            print(f"extract timing {mname}: ERROR {e}", flush=True)


def extract_ablate(ctx):
    """Where does the tensor-core extractor's time go?  Debug switches knock out one stage at a time."""
    ctx.set_filters(O.DEFAULT_FILTERS)
    ctx.set_network(4275, [8], 2)
    n = 65536
    imgs = synth_images(n, 32, seed=0)
    ctx.dataset_alloc(n)
    ctx.set_extract_mode(lib.EXTRACT_TC)
    for dbg, what in ((0, "full"), (1, "no global stores"), (2, "no patch build"), (3, "no stores, no build")):
        os.environ["PYCNN_B200_EXTRACT_DEBUG"] = str(dbg)
        ctx.dataset_extract(0, imgs)
        ctx.set_kernel_timing(True)
        ctx.kernel_stats(reset=True)
        for _ in range(5):
            ctx.dataset_extract(0, imgs)
        st = ctx.kernel_stats(reset=True)["extract_conv_relu_pool"]
        ctx.set_kernel_timing(False)
        print(f"extract ablation [{what}]: {st[1] / st[0]:.4f} ms / {n} imgs", flush=True)
    os.environ["PYCNN_B200_EXTRACT_DEBUG"] = "0"


def gemm_timing(ctx):
    ctx.set_math_mode(lib.MATH_TF32)
    for (M, N, K, ak, bk, name) in [(4096, 256, 4275, 1, 1, "fwd1"), (256, 4275, 4096, 0, 0, "dW1"),
                                     (4096, 128, 256, 1, 1, "fwd2"), (4096, 256, 128, 1, 0, "dA2"),
                                     (4096, 1024, 18259, 1, 1, "c3_fwd1"), (1024, 18259, 4096, 0, 0, "c3_dW1")]:
        rng = np.random.default_rng(0)
        A = rng.normal(size=(M, K) if ak else (K, M)).astype(np.float32)
        B = rng.normal(size=(N, K) if bk else (K, N)).astype(np.float32)
        try:
            _, ms = ctx.gemm_probe(A, B, a_kmajor=bool(ak), b_kmajor=bool(bk), iters=20)
            print(f"gemm timing {name} MNK={(M, N, K)}: {ms * 1e3:.1f} us, {2.0 * M * N * K / ms / 1e9:.1f} TFLOP/s", flush=True)
        except Exception as e:
            print(f"gemm timing {name}: ERROR {e}", flush=True)


if __name__ == "__main__":
    what = sys.argv[1:] or ["gemm", "extract", "extract_timing", "gemm_timing"]
    for w in what:
        ctx = lib.Context(0)
        try:
            {"gemm": gemm_diag, "extract": extract_diag, "extract_timing": extract_timing, "gemm_timing": gemm_timing,
             "extract_ablate": extract_ablate}[w](ctx)
        except Exception as e:
            print(f"{w}: FATAL {e}", flush=True)
        try:
            ctx.close()
        except Exception:
            pass
