"""2+ rank data-parallel worker (launched with torchrun by tests/test_gpu_multi.py or by hand):
trains the same small job with (a) NCCL all-reduce + replicated update and (b) the fused peer-memory exchange,
and checks both against a single-process run of the full batches on rank 0's GPU."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch.distributed as dist  # noqa: E402

from oracle import pycnn_oracle as O  # noqa: E402
from pycnn_b200 import _lib as lib  # noqa: E402
from pycnn_b200 import parallel  # noqa: E402
from tests.helpers import rel_err, synth_images, synth_labels  # noqa: E402


def make_ctx(device, opt, D_layers_C, feats, labels, math):
    D, layers, C = D_layers_C
    ctx = lib.Context(device)
    ctx.set_math_mode(math)
    ctx.set_filters(O.DEFAULT_FILTERS)
    ctx.set_network(D, layers, C)
    w, b = O.init_layers(D, layers, C, seed=42)
    wk, bk = O.layer_keys(len(layers))
    for l, (wn, bn) in enumerate(zip(wk, bk)):
        ctx.set_param(2 * l, w[wn])
        ctx.set_param(2 * l + 1, b[bn])
    ctx.set_optimizer(lib.OPT_ADAM if opt == "adam" else lib.OPT_SGD, 1e-3)
    ctx.dataset_alloc(len(feats))
    ctx.dataset_set_features(0, feats)
    ctx.dataset_set_labels(0, labels)
    return ctx


MODES = ("nccl", "p2p", "nvls")


def main():
    rank, world, local = parallel.init_process_group("nccl")
    S, C, layers, n, bs = 32, 10, [96, 48], 1000, 256          # batches 256,256,256,232 -> uneven rank slices
    imgs, labels = synth_images(n, S, 21), synth_labels(n, C, 22).astype(np.int32)
    feats = O.extract_batch(imgs)
    shape = (feats.shape[1], layers, C)
    perms = [np.random.default_rng(100 + e).permutation(n) for e in range(2)]
    results = {}
    for opt in ("sgd", "adam"):
        for mode in MODES:
            ctx = make_ctx(local, opt, shape, feats, labels, lib.MATH_FP32_SIMT)
            uid = parallel.broadcast_bytes(lib.comm_unique_id() if rank == 0 else None)
            ctx.comm_init(uid, rank, world)
            used = parallel.attach_exchange(ctx, rank, world, mode)
            assert used == mode or mode == "nvls", (used, mode)       # nvls falls back to p2p on a box without multicast
            if used != mode:
                print(f"DP_NOTE world={world}: no multicast objects here, {mode} ran as {used}", flush=True)
            stats = [ctx.train_epoch(p, bs) for p in perms]
            params = [ctx.get_param(i) for i in range(ctx.num_params())]
            results[(opt, mode)] = (stats, params)
            # every rank must hold identical parameters
            flat = np.concatenate([p.ravel() for p in params])
            gathered = [None] * world
            dist.all_gather_object(gathered, flat)
            for other in gathered[1:]:
                np.testing.assert_array_equal(gathered[0], other)
            ctx.close()
            dist.barrier()
        if rank == 0:
            single = make_ctx(local, opt, shape, feats, labels, lib.MATH_FP32_SIMT)
            want_stats = [single.train_epoch(p, bs) for p in perms]
            want = [single.get_param(i) for i in range(single.num_params())]
            single.close()
            for mode in MODES:
                stats, params = results[(opt, mode)]
                for (l1, c1), (l2, c2) in zip(stats, want_stats):
                    assert c1 == c2 and abs(l1 - l2) < 1e-6 * abs(l2), (opt, mode, l1, l2, c1, c2)
                tol = 1e-5 if opt == "sgd" else 5e-4   # Adam amplifies summation-order noise
                errs = [rel_err(a, b) for a, b in zip(params, want)]
                assert max(errs) < tol, (opt, mode, errs)
                print(f"DP_OK world={world} {opt}/{mode}: max param rel err vs single GPU {max(errs):.2e}, "
                      f"epoch losses {[round(s[0], 4) for s in stats]}", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
