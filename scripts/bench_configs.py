"""Side measurements for the other BASELINE.json configurations (bench.py's JSON line is configs[1]).

    python scripts/bench_configs.py [c1] [c3] [c4] [c4s] [c5] [c5f19]      (c4s also under torchrun: image-sharded predict)

c1  32x32, 10 classes, [256,128,64], batch 32, SGD                      -> launch-latency-bound step
c3  64x64, 100 classes, [1024,512,256], Adam, batch 4096                -> tensor-bound step
c4  1M synthetic 32x32 images: extractor alone (device), batched predict from host, predict device-resident
c5  224x224, custom 8-filter bank, [2048,1024,512], 1000 classes, Adam, batch 512
Prints one JSON object per config on stdout.  Device-resident timing with CUDA events on the library stream.
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pycnn_b200 import _lib as lib
from pycnn_b200.pycnn import PyCNN


def he_init(ctx, dims, seed=42):
    rng = np.random.default_rng(seed)
    for l in range(len(dims) - 1):
        w = rng.standard_normal((dims[l + 1], dims[l]), dtype=np.float32) * np.float32((2 / dims[l]) ** 0.5)
        ctx.set_param(2 * l, w.astype(np.float64))
        ctx.set_param(2 * l + 1, rng.standard_normal(dims[l + 1]) * 0.01)


def custom_filters(F=8):
    return np.round(np.random.default_rng(2).uniform(-2, 2, (F, 3, 3)), 3).tolist()


def train_config(name, S, C, layers, bs, opt, lr, n_rows, steps, filters=None):
    ctx = lib.Context(0)
    ctx.set_math_mode(lib.MATH_TF32)
    ctx.set_filters(filters or PyCNN().filters)
    D = ctx.feature_count(S)
    dims = [D] + layers + [C]
    ctx.set_network(D, layers, C)
    he_init(ctx, dims)
    ctx.set_optimizer(lib.OPT_ADAM if opt == "adam" else lib.OPT_SGD, lr)
    rng = np.random.default_rng(0)
    ctx.dataset_alloc(n_rows)
    chunk = max(1, min(n_rows, (256 << 20) // (S * S * 3)))
    for s in range(0, n_rows, chunk):
        m = min(chunk, n_rows - s)
        ctx.dataset_extract(s, rng.integers(0, 256, (m, S, S, 3), dtype=np.uint8))
    ctx.dataset_set_labels(0, rng.integers(0, C, n_rows).astype(np.int32))
    idx = rng.integers(0, n_rows, steps * bs)
    for _ in range(3):
        ctx.train_epoch(idx, bs)
    ctx.synchronize()
    ctx.event_record(0)
    loss, _ = ctx.train_epoch(idx, bs)
    ctx.event_record(1)
    ctx.synchronize()
    ms = ctx.event_elapsed_ms(0, 1) / steps
    fwd = 2 * sum(a * b for a, b in zip(dims[:-1], dims[1:]))
    train_flop = 3 * fwd - 2 * dims[0] * dims[1]
    out = {"config": name, "image": S, "classes": C, "layers": layers, "batch": bs, "optimizer": opt, "feature_dim": D,
           "ms_per_step": round(ms, 4), "images_per_s": round(bs / ms * 1e3), "mlp_tflops": round(bs * train_flop / ms / 1e9, 1),
           "loss_per_image": round(loss / (steps * bs), 4)}
    ctx.close()
    return out


def config4(n_total=1_000_000):
    ctx = lib.Context(0)
    ctx.set_math_mode(lib.MATH_TF32)
    ctx.set_filters(PyCNN().filters)
    S, C, layers = 32, 10, [256, 128, 64]
    D = ctx.feature_count(S)
    ctx.set_network(D, layers, C)
    he_init(ctx, [D] + layers + [C])
    rng = np.random.default_rng(0)
    out = {"config": "c4", "images": n_total}
    # (a) extractor alone, device-resident: 1M images in launches of 65536 straight into the feature matrix
    n_dev = 262144
    imgs = rng.integers(0, 256, (65536, S, S, 3), dtype=np.uint8)
    ctx.dataset_alloc(n_dev)
    for s in range(0, n_dev, 65536):
        ctx.dataset_extract(s, imgs)
    ctx.set_kernel_timing(True)
    ctx.kernel_stats(reset=True)
    reps = n_total // 65536 + 1
    for r in range(reps):
        ctx.dataset_extract((r % 4) * 65536, imgs)
    n_l, ms = ctx.kernel_stats(reset=True)["extract_conv_relu_pool"]
    ctx.set_kernel_timing(False)
    per_img = ms / (n_l * 65536)
    out["extractor_images_per_s"] = round(1e3 / per_img)
    out["extractor_GBps"] = round(20172 / per_img / 1e6, 1)
    # (b) batched predict from HOST uint8 images (H2D + extractor + MLP forward + arg-max + D2H), 1M images
    import torch
    pool = 131072
    h = torch.empty((pool, S, S, 3), dtype=torch.uint8, pin_memory=True)
    h.numpy()[:] = rng.integers(0, 256, (pool, S, S, 3), dtype=np.uint8)
    cls = np.empty(pool, np.int32)
    conf = np.empty(pool, np.float32)
    import ctypes
    def predict_pool():
        lib.check(ctx.lib.pycnn_b200_predict_images(ctx.handle, ctypes.c_void_p(h.data_ptr()), pool, S,
                                                    lib.ptr(cls, ctypes.c_int32), lib.ptr(conf, ctypes.c_float)))
    predict_pool()
    t0 = time.perf_counter()
    done = 0
    while done < n_total:
        predict_pool()
        done += pool
    dt = time.perf_counter() - t0
    out["predict_from_host_images_per_s"] = round(done / dt)
    out["predict_from_host_GBps_h2d"] = round(done * S * S * 3 / dt / 1e9, 1)
    # (c) device-resident evaluation of the feature matrix (MLP forward + softmax/arg-max only)
    ctx.dataset_set_labels(0, rng.integers(0, C, n_dev).astype(np.int32))
    ctx.evaluate_rows(0, n_dev, 16384, want_pred=False)
    t0 = time.perf_counter()
    for _ in range(4):
        ctx.evaluate_rows(0, n_dev, 16384, want_pred=False)
    dt = time.perf_counter() - t0
    out["forward_resident_images_per_s"] = round(4 * n_dev / dt)
    ctx.close()
    return out


def config4_sharded(n_total=1_000_000):
    """BASELINE configs[3] over every GPU of the job (torchrun): each rank predicts its contiguous N/W image range from
    its own pinned host pool (H2D + extractor + MLP forward + arg-max + D2H); no collective in the data path.  Wall
    clock between two barriers, max over ranks; throughput = all images / that time."""
    import ctypes
    import torch
    from pycnn_b200 import parallel
    rank, world, local = parallel.dist_env()
    if world > 1:
        parallel.init_process_group("nccl")
    ctx = lib.Context(local)
    ctx.set_math_mode(lib.MATH_TF32)
    ctx.set_filters(PyCNN().filters)
    S, C, layers = 32, 10, [256, 128, 64]
    D = ctx.feature_count(S)
    ctx.set_network(D, layers, C)
    he_init(ctx, [D] + layers + [C])
    lo, hi = parallel.image_shard(n_total, rank, world)
    mine = hi - lo
    pool = 131072
    rng = np.random.default_rng(rank)
    h = torch.empty((pool, S, S, 3), dtype=torch.uint8, pin_memory=True)
    h.numpy()[:] = rng.integers(0, 256, (pool, S, S, 3), dtype=np.uint8)
    cls, conf = np.empty(pool, np.int32), np.empty(pool, np.float32)

    def predict(n):
        lib.check(ctx.lib.pycnn_b200_predict_images(ctx.handle, ctypes.c_void_p(h.data_ptr()), n, S,
                                                    lib.ptr(cls, ctypes.c_int32), lib.ptr(conf, ctypes.c_float)))

    def barrier():
        ctx.synchronize()
        if world > 1:
            torch.distributed.barrier()

    predict(pool)
    times = []
    for _ in range(3):
        barrier()
        t0 = time.perf_counter()
        done = 0
        while done < mine:
            n = min(pool, mine - done)
            predict(n)
            done += n
        ctx.synchronize()
        dt = time.perf_counter() - t0
        if world > 1:
            dt = parallel.allreduce_scalars([dt], op="max")[0]
        times.append(dt)
    ctx.close()
    dt = sorted(times)[1]
    out = {"config": "c4 sharded predict", "images": n_total, "n_gpus": world, "images_per_rank": mine,
           "predict_from_host_images_per_s": round(n_total / dt), "seconds": round(dt, 4),
           "h2d_GBps_aggregate": round(n_total * S * S * 3 / dt / 1e9, 1), "statistic": "median of 3, max over ranks"}
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()
    return out if rank == 0 else None


if __name__ == "__main__":
    what = sys.argv[1:] or ["c1", "c3", "c4", "c5"]
    for w in what:
        if w == "c1":
            print(json.dumps(train_config("c1", 32, 10, [256, 128, 64], 32, "sgd", 1e-3, 2048, 512)), flush=True)
        elif w == "c3":
            print(json.dumps(train_config("c3", 64, 100, [1024, 512, 256], 4096, "adam", 1e-4, 16384, 20)), flush=True)
        elif w == "c4":
            print(json.dumps(config4()), flush=True)
        elif w == "c4s":
            r = config4_sharded()
            if r is not None:
                print(json.dumps(r), flush=True)
        elif w == "c5f19":   # configs[4] with the default 19-filter bank: D = 234 099, 482 M parameters (1.93 GB f32)
            print(json.dumps(train_config("c5f19", 224, 1000, [2048, 1024, 512], 512, "adam", 1e-4, 1024, 6)), flush=True)
        elif w == "c5":
            print(json.dumps(train_config("c5", 224, 1000, [2048, 1024, 512], 512, "adam", 1e-4, 2048, 10,
                                          filters=custom_filters(8))), flush=True)
