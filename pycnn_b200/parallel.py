"""Data-parallel host logic: one process per GPU (torchrun), NCCL owned by libpycnn_b200.so.

The reference has no parallelism of any kind.  The hot path shards per batch:
every rank holds the same device dataset, takes a contiguous 1/W slice of each batch of the shared
permutation, and the flat float32 gradient buffer is all-reduced with ncclSum once per step -- the
reference's gradients are batch SUMS (backward_pass.pyx:71-72,92-93), so no 1/W scaling is needed
and the reduced gradient equals the single-device gradient.  Feature extraction and prediction shard
per image with no collective.

torch.distributed is plumbing only: it carries the 128-byte ncclUniqueId from rank 0 to the other
ranks (any backend: nccl on the GPU box, gloo in the CPU tests) and the final scalar reductions of
embarrassingly-parallel jobs.
"""
from __future__ import annotations

import os

import numpy as np


def batch_slice(start: int, end: int, rank: int, world: int):
    """The slice of batch [start, end) that `rank` processes: contiguous ceil(len/W) chunks.
    Mirrors pycnn_b200_train_epoch (pycnn_b200/csrc/api.cu); ranks past the end get an empty slice."""
    length = end - start
    per = -(-length // world)
    lo, hi = min(length, per * rank), min(length, per * (rank + 1))
    return start + lo, start + hi


def image_shard(n: int, rank: int, world: int):
    """Contiguous N/W image range per rank for extraction / batched predict (no collective)."""
    per = -(-n // world)
    return min(n, per * rank), min(n, per * (rank + 1))


def dist_env():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def init_process_group(backend=None):
    """Initialise torch.distributed from the torchrun environment (idempotent)."""
    import torch.distributed as dist
    rank, world, local = dist_env()
    if world > 1 and not dist.is_initialized():
        if backend is None:
            import torch
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            import torch
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def broadcast_bytes(payload: bytes | None, src: int = 0) -> bytes:
    """Broadcast a small byte string from `src` over the default process group."""
    import torch.distributed as dist
    box = [payload]
    dist.broadcast_object_list(box, src=src)
    return box[0]


def attach(model, make_unique_id=None, peer_memory=True):
    """Join `model` (a cuda(True) PyCNN) to the job's NCCL communicator.  No-op for world == 1."""
    from . import _lib
    rank, world, _ = dist_env()
    model._rank, model._world = rank, world
    if world == 1:
        return rank, world
    import torch.distributed as dist
    if not dist.is_initialized():
        init_process_group()
    uid = (make_unique_id or _lib.comm_unique_id)() if rank == 0 else None
    uid = broadcast_bytes(uid, src=0)
    model._require().comm_init(uid, rank, world)
    if peer_memory and world <= 8 and getattr(model, "weights", None):
        model._bind_network()
        attach_exchange(model._require(), rank, world)
    return rank, world


# The multicast exchange ends in a REPLICATED update (every rank rewrites all of p, m, v: 28 bytes per parameter), which is
# what a small, latency-bound network wants; a large one is bandwidth-bound there and is better off with the peer-memory
# exchange, whose update is sharded 1/W per rank.  Measured on 2 GPUs: configs[1] (1.1 M parameters) 0.140 vs 0.148 ms per
# step in favour of multicast, configs[2] (18.9 M) 1.19 vs 1.02 ms in favour of peer memory.
NVLS_MAX_FLOATS = 4 << 20


def attach_exchange(ctx, rank, world, mode=None):
    """Pick the gradient exchange of the step.  mode (or PYCNN_B200_DP): "nvls" = NVSwitch multicast all-reduce
    (csrc/nvls.cuh), "p2p" = peer-memory reduce-scatter / update / all-gather (csrc/p2p.cuh), "nccl" = ncclAllReduce +
    replicated update; default: nvls for networks up to NVLS_MAX_FLOATS parameters where the box offers multicast
    objects, else p2p.  Collective.  -> the mode used"""
    mode = (mode or os.environ.get("PYCNN_B200_DP") or "auto").lower()
    if mode not in ("auto", "nvls", "p2p", "nccl"):
        raise ValueError(f"unknown exchange mode {mode!r}")
    if world <= 1 or mode == "nccl" or world > 8:
        return "nccl"
    if mode == "auto" and ctx.grad_buffer()[1] > NVLS_MAX_FLOATS:
        mode = "p2p"
    if mode in ("auto", "nvls"):
        if attach_multicast(ctx, rank, world):
            return "nvls"
        if mode == "nvls":
            raise RuntimeError("PYCNN_B200_DP=nvls: this box does not offer NVSwitch multicast objects to every rank")
    attach_peer_memory(ctx, rank, world)
    return "p2p"


def _all_agree(ok):
    import torch.distributed as dist
    flags = [None] * dist.get_world_size()
    dist.all_gather_object(flags, bool(ok))
    return all(flags)


def attach_multicast(ctx, rank, world):
    """After comm_init and set_network on every rank: rank 0 creates the multicast object and hands its file descriptor
    to the other ranks over a unix socket (SCM_RIGHTS), every rank adds its device, then binds its region.  Collective;
    every rank returns the same answer (False: nothing is attached, the caller falls back)."""
    import socket
    import tempfile
    import torch.distributed as dist
    if not _all_agree(ctx.comm_nvls_supported()):
        return False
    fd, srv, path, err = -1, None, None, None
    if rank == 0:
        try:
            fd = ctx.comm_nvls_create(world)
            path = os.path.join(tempfile.mkdtemp(prefix="pycnn_b200_nvls_"), "fd.sock")
            srv = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
            srv.bind(path)
            srv.listen(world)
            srv.settimeout(120.0)
        except Exception as e:            # an old driver, a container without the capability ...
            err, path = e, None
    box = [path]
    dist.broadcast_object_list(box, src=0)
    path = box[0]
    if path is None:
        if rank == 0:
            ctx.comm_nvls_detach()
        return False
    ok = True
    try:
        if rank == 0:
            for _ in range(world - 1):
                conn, _addr = srv.accept()
                with conn:
                    socket.send_fds(conn, [b"m"], [fd])
            ctx.comm_nvls_join(-1, rank, world)
        else:
            with socket.socket(socket.AF_UNIX, socket.SOCK_STREAM) as conn:
                conn.settimeout(120.0)
                conn.connect(path)
                _msg, fds, _flags, _addr = socket.recv_fds(conn, 1, 1)
            try:
                ctx.comm_nvls_join(fds[0], rank, world)
            finally:
                os.close(fds[0])
    except Exception as e:
        ok, err = False, e
    finally:
        if rank == 0:
            if srv is not None:
                srv.close()
            if fd >= 0:
                os.close(fd)
            try:
                os.unlink(path)
                os.rmdir(os.path.dirname(path))
            except OSError:
                pass
    if _all_agree(ok):                       # also the barrier: every device is added before anybody binds
        try:
            ctx.comm_nvls_bind()
        except Exception as e:
            ok, err = False, e
    else:
        ok = False
    if not _all_agree(ok):                   # and the barrier behind bind: nobody steps before everybody is bound
        ctx.comm_nvls_detach()
        if err is not None and os.environ.get("PYCNN_B200_DP", "").lower() == "nvls":
            raise err
        return False
    return True


def attach_peer_memory(ctx, rank, world):
    """After comm_init and set_network on every rank: exchange the CUDA IPC handles of the gradient / parameter /
    sync buffers and switch the step to the fused peer-memory exchange (csrc/p2p.cuh).  Collective."""
    import torch.distributed as dist
    blob = ctx.comm_ipc_export()                 # also zeroes this rank's sync block before anyone can see it
    blobs = [None] * world
    dist.all_gather_object(blobs, blob)
    ctx.comm_ipc_attach(b"".join(blobs), rank, world)
    dist.barrier()                               # nobody steps before everybody is attached


def allreduce_scalars(values, op="sum"):
    """Sum (or max) a few python floats over all ranks; identity for world == 1."""
    rank, world, _ = dist_env()
    if world == 1:
        return list(values)
    import torch
    import torch.distributed as dist
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor(list(values), dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM if op == "sum" else dist.ReduceOp.MAX)
    return t.cpu().tolist()


def gather_shards(local, n_total, rank, world):
    """Concatenate per-rank result arrays of an image-sharded job (image_shard order) on every rank.  Host-side:
    the data path itself has no collective, this only brings the (small) per-image results together."""
    if world == 1:
        return local
    import torch.distributed as dist
    parts = [None] * world
    dist.all_gather_object(parts, np.ascontiguousarray(local))
    out = np.concatenate([p for p in parts if len(p)]) if n_total else np.asarray(local)
    assert len(out) == n_total, (len(out), n_total)
    return out


def predict_sharded(ctx, images, rank, world):
    """Batched predict over `world` GPUs: every rank runs the fused extractor + forward on its contiguous N/W image
    range (pycnn/pycnn.py:737-773 per image; no collective), results are gathered on every rank.
    -> (classes int32 [N], confidences float32 [N])"""
    n = len(images)
    lo, hi = image_shard(n, rank, world)
    if hi > lo:
        cls, conf = ctx.predict_images(np.ascontiguousarray(images[lo:hi]))
    else:
        cls, conf = np.zeros(0, np.int32), np.zeros(0, np.float32)
    return gather_shards(cls, n, rank, world), gather_shards(conf, n, rank, world)


def evaluate_sharded(ctx, images, labels, batch_size, rank, world):
    """Evaluate._run_inference (pycnn/pycnn.py:873-914) over `world` GPUs: each rank evaluates its image range in
    chunks of batch_size; loss sums and correct counts are added on the host, predictions gathered.
    -> (loss_sum, correct, predicted int32 [N])"""
    n = len(images)
    lo, hi = image_shard(n, rank, world)
    if hi > lo:
        loss, correct, pred = ctx.evaluate_images(np.ascontiguousarray(images[lo:hi]),
                                                  np.ascontiguousarray(labels[lo:hi], dtype=np.int32), batch_size)
    else:
        loss, correct, pred = 0.0, 0, np.zeros(0, np.int32)
    loss, correct = allreduce_scalars([loss, float(correct)])
    return loss, int(round(correct)), gather_shards(pred, n, rank, world)


def extract_sharded(ctx, images, rank, world, row0=0):
    """Feature extraction over `world` GPUs: rank r extracts ITS image range into rows [row0, row0 + hi - lo) of its own
    device dataset (dataset_alloc first); features stay where they were produced.  -> (lo, hi)"""
    lo, hi = image_shard(len(images), rank, world)
    if hi > lo:
        ctx.dataset_extract(row0, np.ascontiguousarray(images[lo:hi]))
    return lo, hi


class ShardedStepper:
    """Reference implementation of the sharded step with a pluggable compute + all-reduce, used by the
    world_size-2 gloo tests to pin the host logic (slice arithmetic, sum-not-mean reduction, identical
    update on every rank) without a GPU.  `grad_fn(idx) -> (flat_grad, loss_sum, correct)`;
    `allreduce(array) -> array`; `update_fn(flat_grad)`."""

    def __init__(self, rank, world, grad_fn, allreduce, update_fn):
        self.rank, self.world = rank, world
        self.grad_fn, self.allreduce, self.update_fn = grad_fn, allreduce, update_fn

    def epoch(self, perm, batch_size):
        n = len(perm)
        totals = np.zeros(2)
        for start in range(0, n, batch_size):
            end = min(start + batch_size, n)
            lo, hi = batch_slice(start, end, self.rank, self.world)
            g, loss, correct = self.grad_fn(perm[lo:hi])
            g = self.allreduce(g)
            self.update_fn(g)
            totals += (loss, correct)
        totals = self.allreduce(totals)
        return float(totals[0]), int(round(totals[1]))
