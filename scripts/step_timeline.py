"""In-kernel timeline of one device-resident training step (BASELINE configs[1]) inside a 16-step graph.

    python scripts/step_timeline.py [steps_per_graph] > profiles/r1_step_timeline.txt

Every GEMM / gather / norm / update launch records {first CTA start, last CTA end} in %globaltimer ns
(pycnn_b200_trace_control / trace_read); the step printed is one from the middle of the graph, times relative to
the start of its layer-1 forward GEMM.  Streams: 0 main, 1 side (shallow dW, memset), 2 prefetch (next step's gather).
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pycnn_b200 import _lib as lib
from pycnn_b200.pycnn import PyCNN

S, C, LAYERS, BS, N = 32, 10, [256, 128, 64], 4096, 65536
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 16
rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
ctx = lib.Context(int(os.environ.get("LOCAL_RANK", "0")))
if world > 1:      # torchrun: weak scaling like bench.py, peer-memory gradient exchange
    from pycnn_b200 import parallel
    parallel.init_process_group()
    uid = parallel.broadcast_bytes(lib.comm_unique_id() if rank == 0 else None)
    ctx.comm_init(uid, rank, world)
    BS *= world
ctx.set_math_mode(lib.MATH_TF32)
ctx.set_filters(PyCNN().filters)
D = ctx.feature_count(S)
ctx.set_network(D, LAYERS, C)
rng = np.random.default_rng(0)
dims = [D] + LAYERS + [C]
for l in range(len(dims) - 1):
    ctx.set_param(2 * l, rng.standard_normal((dims[l + 1], dims[l])) * (2 / dims[l]) ** 0.5)
    ctx.set_param(2 * l + 1, rng.standard_normal(dims[l + 1]) * 0.01)
ctx.set_optimizer(lib.OPT_ADAM, 1e-4)
if world > 1:
    print("exchange:", parallel.attach_exchange(ctx, rank, world))      # needs the network's gradient / parameter buffers
ctx.dataset_alloc(N)
for s in range(0, N, 16384):
    ctx.dataset_extract(s, rng.integers(0, 256, (16384, S, S, 3), dtype=np.uint8))
ctx.dataset_set_labels(0, rng.integers(0, C, N).astype(np.int32))
idx = rng.integers(0, N, steps * BS)
ctx.trace_enable(8192)
for _ in range(3):
    ctx.train_epoch(idx, BS)          # captures the graph(s) with trace slots, warms up
ctx.trace_reset()
ctx.event_record(0)
ctx.train_epoch(idx, BS)
ctx.event_record(1)
ctx.synchronize()
ms = ctx.event_elapsed_ms(0, 1)
ev = ctx.trace_read(marks=True)
if rank != 0:
    ctx.close()
    sys.exit(0)
print(f"# {steps} steps in {ms * 1e3:.1f} us = {ms * 1e3 / steps:.2f} us/step (CUDA events around train_epoch); "
      f"{len(ev)} traced launches")
# a step starts at a layer-1 forward GEMM on the main stream
ev_sorted = sorted(ev, key=lambda e: e[2])
starts_t = sorted(e[2] for e in ev if e[0] == "gemm_forward_layer1")
if len(starts_t) >= 1:
    k = len(starts_t) // 2
    # (single-step graphs reuse their slots at every replay: only the last step is there)
    t0, t1 = (starts_t[k], starts_t[k + 1]) if len(starts_t) >= 3 else (starts_t[-1], 1 << 62)
    if len(starts_t) >= 3:
        print(f"# step {k}: {(t1 - t0) / 1e3:.2f} us from its layer-1 GEMM start to the next one's")
    print(f"{'kernel':24s} {'stream':>6s} {'start us':>9s} {'end us':>9s} {'dur us':>8s}   GEMM phase marks, us after the kernel's first CTA start (min..max over CTAs)")
    for name, sid, a, b, *marks in ev_sorted:
        if t0 <= a < t1:
            ph = "  ".join(f"{lbl} {(m[0] - a) / 1e3:.1f}..{(m[1] - a) / 1e3:.1f}" for lbl, m in
                           zip(("first-stage", "accum-done", "parked") if name.startswith("gemm") else ("m1", "m2", "m3"), marks) if m)
            print(f"{name:24s} {sid:6d} {(a - t0) / 1e3:9.2f} {(b - t0) / 1e3:9.2f} {(b - a) / 1e3:8.2f}   {ph}")
ctx.close()
