#!/usr/bin/env python
"""bench.py -- headline benchmark of the PyCNN hot path on B200 (contract: see the task's bench section).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...      (N > 1)

Workload (BASELINE.json configs[1], the configuration `metric` is quoted on): CIFAR-10-shaped synthetic
32x32 RGB, 10 classes, layers [256,128,64], Adam (lr 1e-4), batch 4096 per GPU; weak scaling for N > 1
(global batch 4096*N, one NCCL all-reduce of the flat gradient per step).  A "step" is one pass of the
reference's batch loop (pycnn/pycnn.py:629-642: gather, forward, CE/argmax, backward, update).

  value  device-resident: the feature matrix (65536 x 4275 fp32, 1.1 GB > L2) lives in HBM; each step
         gathers 4096 fresh random rows.  CUDA events on the library's compute stream, max over ranks.
  e2e    the same step through the C ABI with HOST buffers: pinned uint8 images + int32 labels are
         copied H2D, the fused extractor produces the features, the step runs, loss/#correct come back.
  roofline / cpu_baseline: see DESIGN.md "Measurement".
One JSON line on stdout (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

S, C, LAYERS, BS, N_RESIDENT, LR = 32, 10, [256, 128, 64], 4096, 65536, 1e-4
F = 19
D = ((S - 2) // 2) ** 2 * F                                                    # 4275
DIMS = [D] + LAYERS + [C]
FWD_FLOP_PER_IMG = 2 * sum(a * b for a, b in zip(DIMS[:-1], DIMS[1:]))          # 2.272 MFLOP
TRAIN_FLOP_PER_IMG = 3 * FWD_FLOP_PER_IMG - 2 * DIMS[0] * DIMS[1]               # 4.627 MFLOP (no dA for layer 1)
EXTRACT_BYTES_PER_IMG = S * S * 3 + 4 * D                                       # 20172 B (u8 in + f32 out)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        with open(p) as f:
            d = json.load(f)
        return dict(hbm_gbs=d["hbm_gbs"], bf16_tflops=d["bf16_tflops"],
                    bf16_tflops_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]), source="measured")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, source="fallback")


class ClockSampler:
    """SM clock + throttle reasons DURING the timed region (B200_PROFILING.md).  NVML is polled in-process
    (nvidia_ml_py) from a thread, initialised before the warm-up: spawning `nvidia-smi -lms` next to a
    10 ms timed region perturbs it (its NVML start-up stalled some runs by 3x); nvidia-smi is the fallback."""
    REASONS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, index, period=0.05):
        self.rows, self.stop_flag, self.nvml, self.proc = [], False, None, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.period = period
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
        except Exception:
            self.nvml = None
            self._start_smi(index)

    def _poll(self):
        n = self.nvml
        while not self.stop_flag:
            try:
                sm = n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM)
                rs = n.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(n, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else n.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                pw = n.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                self.rows.append((time.time(), sm, rs, pw))
            except Exception:
                pass
            time.sleep(self.period)

    def _start_smi(self, index):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read_smi, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read_smi(self):
        for line in self.proc.stdout:
            parts = [x.strip() for x in line.split(",")]
            try:
                rs = sum(bit for (nm, bit), v in zip(self.REASONS.items(), parts[3:7]) if v.lower().startswith("active"))
                self.max_sm = float(parts[1])
                self.rows.append((time.time(), float(parts[0]), rs, float(parts[2])))
            except Exception:
                continue

    def stop(self, t0=None, t1=None):
        self.stop_flag = True
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
        rows = [r for r in self.rows if (t0 is None or t0 <= r[0] <= t1)] or self.rows
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        reasons = sorted(nm for nm, bit in self.REASONS.items() if any(r[2] & bit for r in rows))
        return {"sm_mhz": float(np.median([r[1] for r in rows])), "sm_max_mhz": float(self.max_sm),
                "power_w_max": max(r[3] for r in rows), "samples": len(rows), "reasons": reasons,
                "source": "nvml in-process" if self.nvml else "nvidia-smi"}


def synth_model(lib, device):
    """Context with the default bank, the C2 network initialised from the reference's seeded stream."""
    from pycnn_b200.pycnn import PyCNN
    ctx = lib.Context(device)
    ctx.set_math_mode(lib.MATH_TF32)
    ctx.set_extract_mode(lib.EXTRACT_TC)
    ctx.set_filters(PyCNN().filters)                                 # the reference's default 19-filter bank
    ctx.set_network(D, LAYERS, C)
    np.random.seed(42)                                               # pycnn/pycnn.py:225, :303-311
    prev = D
    for l, size in enumerate(LAYERS + [C]):
        ctx.set_param(2 * l, np.random.randn(size, prev) * (2 / prev) ** 0.5)
        ctx.set_param(2 * l + 1, np.random.randn(size) * 0.01)
        prev = size
    ctx.set_optimizer(lib.OPT_ADAM, LR)
    return ctx


def run_ours(args):
    import torch                                                      # plumbing: pinned memory, process group
    from pycnn_b200 import _lib as lib
    from pycnn_b200 import parallel
    rank, world, local = parallel.dist_env()
    if world > 1:
        parallel.init_process_group("nccl")
    ctx = synth_model(lib, local)
    dp_mode = "single"
    if world > 1:
        uid = parallel.broadcast_bytes(lib.comm_unique_id() if rank == 0 else None)
        ctx.comm_init(uid, rank, world)
        dp_mode = os.environ.get("PYCNN_B200_DP", "p2p" if world <= 8 else "nccl")
        if dp_mode == "p2p":      # fused reduce-scatter -> norm -> sharded Adam -> all-gather over NVLink peer memory
            parallel.attach_peer_memory(ctx, rank, world)

    sampler = ClockSampler(local) if rank == 0 else None      # NVML up and polling well before the timed regions
    # ---------------- resident dataset: extract N_RESIDENT synthetic images on the device
    rng = np.random.default_rng(1000 + rank)
    ctx.dataset_alloc(N_RESIDENT)
    for s in range(0, N_RESIDENT, 16384):
        ctx.dataset_extract(s, rng.integers(0, 256, (16384, S, S, 3), dtype=np.uint8))
    ctx.dataset_set_labels(0, rng.integers(0, C, N_RESIDENT).astype(np.int32))

    K, W = args.steps, args.warmup
    gbs = BS * world                                                  # weak scaling: 4096 per GPU
    idx_rng = np.random.default_rng(7)                                # identical on every rank
    warm_idx = idx_rng.integers(0, N_RESIDENT, W * gbs)
    timed_idx = idx_rng.integers(0, N_RESIDENT, K * gbs)

    def barrier():
        ctx.synchronize()
        if world > 1:
            torch.distributed.barrier()
        ctx.synchronize()

    t_load0 = time.time()                                      # clock samples: warm-up .. end of the e2e loops
    if W:
        ctx.train_epoch(warm_idx, gbs)
    # the W contract warm-up steps are ~1 ms of GPU work: clocks, TLBs and the graph executable are not settled
    # yet (first timed regions measured 10-50 % slower than steady state), so run two more untimed K-step passes
    extra_warm = 2 * K
    for _ in range(2):
        ctx.train_epoch(timed_idx, gbs)
    barrier()
    launches0 = ctx.launch_count()
    t0 = time.time()
    ctx.event_record(0)
    loss_sum, correct = ctx.train_epoch(timed_idx, gbs)               # K steps, one host sync at the end
    ctx.event_record(1)
    barrier()
    t1 = time.time()
    ms = ctx.event_elapsed_ms(0, 1)
    launches = ctx.launch_count() - launches0
    if world > 1:
        ms = parallel.allreduce_scalars([ms], op="max")[0]
    value = K * gbs / (ms / 1e3)
    if os.environ.get("PYCNN_B200_BENCH_REPEATS"):       # diagnostics: run-to-run spread of the same K-step region
        reps = []
        for _ in range(int(os.environ["PYCNN_B200_BENCH_REPEATS"])):
            barrier()
            ctx.event_record(4)
            ctx.train_epoch(timed_idx, gbs)
            ctx.event_record(5)
            barrier()
            reps.append(round(ctx.event_elapsed_ms(4, 5) / K, 5))
        if rank == 0:
            print("repeat ms/step:", reps, file=sys.stderr, flush=True)

    # ---------------- per-kernel-class timing (events around every launch) for the roofline
    ctx.set_kernel_timing(True)
    ctx.kernel_stats(reset=True)
    ctx.train_epoch(timed_idx[:min(K, 10) * gbs], gbs)
    stats = ctx.kernel_stats(reset=True)
    ctx.set_kernel_timing(False)
    nsteps_prof = min(K, 10)
    per_class = {k: {"launches_per_step": n / nsteps_prof, "ms_per_step": t / nsteps_prof}
                 for k, (n, t) in stats.items() if n}
    pk = peaks()
    gemm_classes = ("gemm_forward_layer1", "gemm_dW_layer1", "gemm_forward_tail", "gemm_dA", "gemm_dW_tail")
    l1_flop = 2 * BS * DIMS[0] * DIMS[1]                                       # 8.96 GFLOP: one layer-1 GEMM
    class_flops = {"gemm_forward_layer1": l1_flop, "gemm_dW_layer1": l1_flop,
                   "gemm_forward_tail": BS * FWD_FLOP_PER_IMG - l1_flop, "gemm_dW_tail": BS * FWD_FLOP_PER_IMG - l1_flop,
                   "gemm_dA": BS * FWD_FLOP_PER_IMG - l1_flop}
    n_params = sum(a * b + b for a, b in zip(DIMS[:-1], DIMS[1:]))
    class_bytes = {"gather_rows": 2 * BS * (D + 1) * 4, "clip_update": 28 * n_params, "grad_sqnorm": 4 * n_params,
                   "extract_conv_relu_pool": BS * EXTRACT_BYTES_PER_IMG}
    tf32_peak = pk["bf16_tflops_sustained"] / 2.0

    # DRAM bytes per launch from the committed ncu --set full captures of this same command (profiles/, cold L2)
    ncu_names = {"gemm_forward_layer1": "gemm_tf32_kernel<256, 0, 0> (1, 32, 4)", "gemm_dW_layer1": "gemm_tf32_kernel<256, 1, 1> (17, 2, 4)",
                 "gather_rows": "gather_rows_kernel (296, 1, 1)", "clip_update": "clip_update_kernel (560, 1, 1)",
                 "grad_sqnorm": "grad_sqnorm_kernel (560, 1, 1)"}
    try:
        with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r1_traffic.json")) as f:
            ncu_traffic = json.load(f)["bytes_per_launch"]
    except (OSError, ValueError, KeyError):
        ncu_traffic = {}

    def traffic_of(k):
        return ncu_traffic.get(ncu_names.get(k, ""))

    def roofline_of(k):
        """algorithmic FLOPs (or bytes) of one launch / average CUDA-event duration of one launch of class k"""
        ms_launch = per_class[k]["ms_per_step"] / per_class[k]["launches_per_step"]
        if k in class_flops:
            ach = class_flops[k] / per_class[k]["ms_per_step"] / 1e9          # whole class per step (TFLOP/s)
            return {"kernel": k, "bound": "tensor", "achieved": ach, "peak": tf32_peak, "unit": "TFLOP/s",
                    "frac": ach / tf32_peak, "traffic": traffic_of(k), "us_per_launch": ms_launch * 1e3,
                    "peak_source": f"{pk['source']} bf16 sustained / 2 (tf32 runs at half the bf16 rate)"}
        ach = class_bytes.get(k, 0) / per_class[k]["ms_per_step"] / 1e6          # GB/s
        return {"kernel": k, "bound": "hbm", "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s",
                "frac": ach / pk["hbm_gbs"], "traffic": traffic_of(k), "us_per_launch": ms_launch * 1e3, "peak_source": pk["source"]}

    # dominant kernel: the single-kernel class with the largest share of the step (the "tail" classes are groups of
    # several tiny latency-bound GEMMs and are reported in rooflines_all); a split-K GEMM's zero-fill and fix-up
    # launches are charged to that GEMM
    single = [k for k in ("gemm_forward_layer1", "gemm_dW_layer1", "gather_rows", "clip_update") if k in per_class]
    dom = max(single, key=lambda k: per_class[k]["ms_per_step"])
    roofline = roofline_of(dom)
    roofline["note"] = "CUDA-event time around each launch of this kernel in an eager (non-graph) pass of the same steps"
    rooflines_all = {k: roofline_of(k) for k in per_class if k in class_flops or k in class_bytes}
    gemm_ms = sum(per_class[k]["ms_per_step"] for k in gemm_classes if k in per_class)
    mlp_tflops = BS * TRAIN_FLOP_PER_IMG / (gemm_ms * 1e-3) / 1e12 if gemm_ms else None

    # ---------------- e2e: host uint8 images -> H2D -> extractor -> step -> D2H stats, every step
    # pycnn_b200_train_epoch_images: every step's batch crosses PCIe from pinned host memory (copy stream),
    # overlapped with extractor + step of the previous batch (compute stream); the running loss/#correct are
    # copied back to the host after every step.
    pool_steps = max(1, min(K, 32, (1 << 30) // (gbs * S * S * 3)))   # <= 1 GB of pinned host memory per rank
    pool = pool_steps * gbs
    h_img = torch.empty((pool, S, S, 3), dtype=torch.uint8, pin_memory=True)
    h_lab = torch.empty((pool,), dtype=torch.int32, pin_memory=True)
    h_img.numpy()[:] = rng.integers(0, 256, (pool, S, S, 3), dtype=np.uint8)
    h_lab.numpy()[:] = rng.integers(0, C, pool).astype(np.int32)

    def e2e_run(nsteps, shuffled):
        """nsteps global batches (each rank trains on its 1/world slice of every batch, gradients all-reduce
        per step) streamed from this rank's pinned pool; returns (ms, launches)."""
        l0 = ctx.launch_count()
        ctx.event_record(2)
        w0 = time.perf_counter()
        done = 0
        while done < nsteps:
            k = min(pool_steps, nsteps - done)
            perm = np.random.default_rng(done).permutation(pool)[:k * gbs] if shuffled else None
            ctx.train_epoch_images(h_img.data_ptr(), h_lab.data_ptr(), k * gbs, S, gbs, perm=perm, num_images=pool)
            done += k
        ctx.event_record(3)
        ctx.synchronize()
        w1 = time.perf_counter()
        return max(ctx.event_elapsed_ms(2, 3), (w1 - w0) * 1e3), ctx.launch_count() - l0

    e2e_run(max(W, 3), False)
    e2e_run(max(W, 3), True)
    barrier()
    e2e_ms, e2e_launches = e2e_run(K, False)
    barrier()
    e2e_shuf_ms, _ = e2e_run(K, True)
    barrier()
    # extractor kernel alone, timed by events around its launches (eager, outside the graphs)
    ctx.set_kernel_timing(True)
    ctx.kernel_stats(reset=True)
    for i in range(5):
        ctx.train_step_images(None, None, True, h_img.data_ptr() + (i % pool_steps) * BS * S * S * 3,
                              h_lab.data_ptr() + (i % pool_steps) * BS * 4, BS, S)
    ex_stats = ctx.kernel_stats(reset=True)
    ctx.set_kernel_timing(False)
    if world > 1:
        e2e_ms, e2e_shuf_ms = parallel.allreduce_scalars([e2e_ms, e2e_shuf_ms], op="max")
    e2e_value = K * gbs / (e2e_ms / 1e3)
    e2e_shuf_value = K * gbs / (e2e_shuf_ms / 1e3)
    clocks = sampler.stop(t_load0, time.time()) if sampler else None
    ex_n, ex_ms = ex_stats.get("extract_conv_relu_pool", (0, 0.0))
    ex_gbs = (BS * EXTRACT_BYTES_PER_IMG) / (ex_ms / max(ex_n, 1) * 1e-3) / 1e9 if ex_n else None

    out = {
        "metric": "train_images_per_sec", "value": value, "unit": "images/s", "n_gpus": world, "steps": K,
        "warmup": W, "extra_untimed_warmup_steps": extra_warm, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "tf32", "data": "synthetic",
        "config": {"workload": "BASELINE configs[1]: 32x32x3 synthetic, 10 classes, layers [256,128,64], Adam lr 1e-4, "
                               "batch 4096 per GPU, CPU update rule (clip 5, frozen t)",
                   "global_batch": gbs, "resident_rows": N_RESIDENT, "feature_dim": D,
                   "parallelism": f"dp{world}" if world > 1 else "single",
                   "gradient_exchange": {"single": "none", "p2p": "fused peer-memory reduce-scatter/update/all-gather (NVLink, CUDA IPC)",
                                         "nccl": "ncclAllReduce(sum) of the flat gradient + replicated update"}[dp_mode],
                   "l2": "inputs larger than L2: 1.1 GB resident fp32 feature matrix, fresh random rows every step",
                   "math_mode": "tf32 tcgen05 (fp32 accumulate)", "extractor": "default 19-filter bank"},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": BS * S * S * 3 + BS * 4,
                "d2h_bytes_per_step": 16, "ms_per_step": e2e_ms / K, "gpu_launches": int(e2e_launches),
                "path": "pycnn_b200_train_epoch_images: per step, pinned uint8 batch -> H2D (copy stream) -> fused "
                        "extractor -> train step -> (loss,correct) D2H; copy of batch i+1 overlaps compute of batch i",
                "shuffled_rows_value": e2e_shuf_value,
                "shuffled_rows_path": "same, batch rows picked by a random permutation: zero-copy gather kernel over PCIe"},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "kernels_ms_per_step": {k: round(v["ms_per_step"], 5) for k, v in per_class.items()},
        "rooflines_all": {k: {"frac": round(v["frac"], 4), "achieved": round(v["achieved"], 1), "unit": v["unit"],
                              "us_per_launch": round(v["us_per_launch"], 2)} for k, v in rooflines_all.items()},
        "mlp_gemm_tflops": mlp_tflops,
        "roofline_extractor": None if ex_gbs is None else {
            "bound": "hbm", "achieved": ex_gbs, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": ex_gbs / pk["hbm_gbs"],
            "frac_of_8TBps_nominal": ex_gbs / 8000.0, "images_per_s": BS / (ex_ms / ex_n * 1e-3),
            "bytes_per_image": EXTRACT_BYTES_PER_IMG},
        "final_loss_per_image": loss_sum / (K * gbs),
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline(seconds=args.cpu_seconds)
    ctx.close()
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()
    if rank == 0:
        print(json.dumps(out))


def make_cpu_trainer(batch):
    """The reference's CPU path on the same workload: compiled Cython/C kernels from oracle/_ref when
    present (kind 'reference'), else the numpy restatement (kind 'port')."""
    from oracle import pycnn_oracle as O
    from oracle import ref_kernels
    rng = np.random.default_rng(1000)
    imgs = rng.integers(0, 256, (min(batch, 4096), S, S, 3), dtype=np.uint8)
    feats = O.extract_batch(imgs)
    if batch > len(feats):
        feats = np.concatenate([feats] * (batch // len(feats)))
    y = O.one_hot(rng.integers(0, C, batch), C)
    w, b = O.init_layers(D, LAYERS, C, seed=42)
    if ref_kernels.available():
        tr = ref_kernels.RefTrainer(w, b, LAYERS, LR, "adam")
        return "reference", (lambda: tr.step(feats, y))
    m, v = {}, {}
    for d in (w, b):
        for k in d:
            m[k], v[k] = np.zeros(d[k].shape), np.zeros(d[k].shape)
    return "port", (lambda: O.train_step(feats, y, w, b, LAYERS, LR, "adam", m, v))


def cpu_threads():
    try:
        from threadpoolctl import threadpool_info
        return max([p.get("num_threads", 1) for p in threadpool_info()] + [1])
    except Exception:
        return os.cpu_count() or 1


class all_host_threads:
    """Give the CPU arm every host core: torchrun exports OMP_NUM_THREADS=1 to its workers, which would pin the
    reference's BLAS to one thread and flatter the GPU arm."""

    def __enter__(self):
        self._ctl = None
        try:
            from threadpoolctl import threadpool_limits
            self._ctl = threadpool_limits(limits=os.cpu_count() or 1)
        except Exception:
            pass
        return self

    def __exit__(self, *exc):
        if self._ctl is not None:
            self._ctl.restore_original_limits()
        return False


def cpu_baseline(seconds=12.0):
    with all_host_threads():
        return _cpu_baseline(seconds)


def _cpu_baseline(seconds):
    kind, step = make_cpu_trainer(BS)
    step()
    n, t0 = 0, time.perf_counter()
    while True:
        step()
        n += 1
        el = time.perf_counter() - t0
        if (el >= seconds and n >= 3) or n >= 400:
            break
    return {"value": n * BS / el, "unit": "images/s", "cores": cpu_threads(), "kind": kind,
            "sample": f"{n} train steps of batch {BS} (features precomputed, as in the reference's train_model) "
                      f"in {el:.1f} s on the host; os.cpu_count()={os.cpu_count()}",
            "ms_per_step": el / n * 1e3}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    gbs = BS * world
    with all_host_threads():
        kind, step = make_cpu_trainer(gbs)
        for _ in range(max(1, min(args.warmup, 3))):
            step()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step()
        el = time.perf_counter() - t0
        threads = cpu_threads()
    value = args.steps * gbs / el
    cb = {"value": value, "unit": "images/s", "cores": threads, "kind": kind,
          "sample": f"{args.steps} CPU train steps of batch {gbs}; os.cpu_count()={os.cpu_count()}"}
    print(json.dumps({
        "impl": "reference", "metric": "train_images_per_sec", "value": value, "unit": "images/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": el / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "BASELINE configs[1] on the reference's CPU path (float64 Cython/NumPy): 32x32x3 synthetic, "
                               "10 classes, layers [256,128,64], Adam lr 1e-4", "global_batch": gbs, "feature_dim": D},
        "cpu_baseline": cb,
        "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
