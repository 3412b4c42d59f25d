"""2+ rank worker (torchrun; launched by tests/test_gpu_multi.py): data-parallel training, image-sharded predict /
Evaluate and checkpoint resume THROUGH THE PyCNN SURFACE (PyCNN.train_model + parallel.attach, peer-memory exchange,
TF32 tensor-core path, BASELINE configs[1] widths), each checked against a single-GPU run on rank 0's GPU."""
import contextlib
import io
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch.distributed as dist  # noqa: E402

from pycnn_b200 import PyCNN, parallel  # noqa: E402
from pycnn_b200.pycnn import Evaluate  # noqa: E402
from tests.helpers import rel_err, synth_images, synth_labels  # noqa: E402

S, C, LAYERS, N, BS, EPOCHS = 32, 10, [256, 128, 64], 2 * 2048 + 300, 2048, 2


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def make_model(images, labels, attach, epochs=EPOCHS):
    m = PyCNN()
    quiet(m.init, batch_size=BS, layers=LAYERS, learning_rate=1e-4, epochs=epochs)    # seeds np.random with 42
    quiet(m.adam)
    quiet(m.cuda, True)
    quiet(m.dataset.synthetic, images, labels, [str(i) for i in range(C)])
    if attach:
        parallel.attach(m)
    return m


def main():
    import faulthandler
    faulthandler.dump_traceback_later(int(os.environ.get("PYCNN_B200_WORKER_DEADLINE", "240")), exit=True)   # a stuck collective shows WHERE
    rank, world, local = parallel.init_process_group("nccl")
    images, labels = synth_images(N, S, 31), synth_labels(N, C, 32)

    # ---- (1) train_model under DP == single GPU, every rank ends with identical parameters
    m = make_model(images, labels, attach=True)
    assert m._world == world and m._require() is not None
    quiet(m.train_model)
    flat = np.concatenate([m.weights[k].ravel() for k in sorted(m.weights)] + [m.biases[k].ravel() for k in sorted(m.biases)])
    gathered = [None] * world
    dist.all_gather_object(gathered, flat)
    for other in gathered[1:]:
        np.testing.assert_array_equal(gathered[0], other)
    if rank == 0:
        one = make_model(images, labels, attach=False)
        quiet(one.train_model)
        errs = [rel_err(m.weights[k], one.weights[k]) for k in m.weights]
        assert max(errs) < 2e-2, errs          # TF32, different partial-sum grouping of the batch rows
        print(f"DP_SURFACE_OK train world={world}: max weight rel err vs single GPU {max(errs):.2e}", flush=True)
        one.cuda(False)
    dist.barrier()

    # ---- (2) image-sharded predict / Evaluate vs single GPU on the same weights (no collective in the data path).  A rank's
    # shard is a smaller batch than the whole set, so the TF32 layer-1 GEMM may cut K into a different number of slices:
    # equal up to TF32 summation order, not bit for bit
    test_imgs, test_lab = synth_images(1003, S, 33), synth_labels(1003, C, 34)
    cls, conf = m.predict_batch(test_imgs)
    ev = Evaluate(m)
    quiet(ev._run_images, list(test_imgs), list(test_lab))
    if rank == 0:
        one = make_model(images, labels, attach=False)
        one.weights = {k: v.copy() for k, v in m.weights.items()}
        one.biases = {k: v.copy() for k, v in m.biases.items()}
        one.sync_weights()
        cls1, conf1 = one.predict_batch(test_imgs)
        assert np.abs(conf - conf1).max() < 5e-3, np.abs(conf - conf1).max()
        flips = cls != cls1
        assert flips.mean() < 0.01 and np.all(np.abs(conf - conf1)[flips] < 5e-3), int(flips.sum())
        ev1 = Evaluate(one)
        quiet(ev1._run_images, list(test_imgs), list(test_lab))
        assert abs(ev.last_result["acc"] - ev1.last_result["acc"]) <= 0.01, (ev.last_result, ev1.last_result)
        assert abs(ev.last_result["loss"] - ev1.last_result["loss"]) <= 1e-3 * abs(ev1.last_result["loss"])
        assert sum(v["total"] for v in ev.last_result["per_class"].values()) == len(test_imgs)
        print(f"DP_SURFACE_OK sharded predict/evaluate world={world}: {len(cls)} images, {int(flips.sum())} near-tie class flips vs single GPU, "
              f"acc {ev.last_result['acc']:.4f}", flush=True)
        one.cuda(False)
    dist.barrier()
    m.cuda(False)

    # ---- (3) checkpoint under the peer-memory exchange (sharded Adam state): 1 epoch + resume + 1 epoch == 2 epochs
    ckpt = os.path.join(os.environ.get("PYCNN_B200_TEST_TMP", "/tmp"), f"pycnn_b200_dp_ckpt_{os.environ.get('MASTER_PORT', '0')}.bin")
    a = make_model(images, labels, attach=True)
    quiet(a.train_model)                                                  # 2 epochs straight
    b = make_model(images, labels, attach=True, epochs=1)
    quiet(b.train_model, checkpoint=ckpt)                                 # 1 epoch, checkpoint written by rank 0
    dist.barrier()
    b.cuda(False)
    r = make_model(images, labels, attach=True)
    epoch = quiet(r.load_checkpoint, ckpt)
    assert epoch == 1
    r.epochs = EPOCHS
    quiet(r.train_model, start_epoch=epoch)
    for k in a.weights:
        np.testing.assert_array_equal(a.weights[k], r.weights[k])        # deterministic training: bit-identical resume
    m_a, v_a = a.optimizer_state()
    m_r, v_r = r.optimizer_state()
    for k in m_a:
        np.testing.assert_array_equal(m_a[k], m_r[k])
        np.testing.assert_array_equal(v_a[k], v_r[k])
    assert all(np.abs(v).max() > 0 for v in m_a.values()), "gathered Adam state must be complete on every rank"
    if rank == 0:
        print(f"DP_SURFACE_OK checkpoint resume world={world}: resumed run bit-identical, optimizer state gathered", flush=True)
        os.remove(ckpt)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
