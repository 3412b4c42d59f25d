"""Turn gpurun_out/*.ncu-rep + launches.csv into small text/CSV summaries under profiles/ (committed).
Usage: python scripts/summarize_profiles.py r1   (needs ncu on PATH; no GPU)"""
import csv
import io
import os
import re
import subprocess
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out_dir = os.path.join(root, "profiles")
os.makedirs(out_dir, exist_ok=True)
go = os.path.join(root, "gpurun_out")

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor_op_gmma.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.max", "lts__t_bytes.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__t_sector_hit_rate.pct"]

TRAFFIC = {}   # "<kernel> <grid>" -> DRAM bytes (read + write) of one launch; bench.py reports it as roofline.traffic


def short(name):
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*$", "", name)
    return name.replace("pb::", "").replace("(int)", "").replace("(bool)", "")


def launches():
    src = os.path.join(go, "launches.csv")
    if not os.path.isfile(src):
        return
    lines = [l for l in open(src) if not l.startswith("==")]
    rows = [(r["Kernel Name"], float(r["Metric Value"]), r.get("Grid Size", ""), r.get("Block Size", ""))
            for r in csv.DictReader(lines) if r.get("Metric Value")]
    names = [short(n) for n, *_ in rows]
    with open(os.path.join(out_dir, f"{tag}_launches.csv"), "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none python bench.py --steps 3 --warmup 3 --no-cpu-baseline\n")
        f.write("# per-launch device time (cold-cache, serialised): compare SHARES, not absolutes\n")
        f.write("id,kernel,grid,block,time_us\n")
        for i, (n, (_, t, g, b)) in enumerate(zip(names, rows)):
            f.write(f"{i},{n},\"{g}\",\"{b}\",{t / 1000:.3f}\n")
    # one steady-state resident step: between two consecutive layer-1 forward GEMM launches
    idx = [i for i, n in enumerate(names) if "gemm_tf32_kernel<256, 0, 0, 1>" in n]
    if len(idx) >= 6:
        s, e = idx[4], idx[5]
        tot = sum(rows[i][1] for i in range(s, e))
        with open(os.path.join(out_dir, f"{tag}_step_breakdown.txt"), "w") as f:
            f.write("One device-resident training step (BASELINE configs[1], batch 4096), ncu launch list\n")
            f.write(f"{'kernel':58s} {'grid':>14s} {'us':>8s} {'share':>7s}\n")
            for i in range(s, e):
                f.write(f"{names[i][:58]:58s} {rows[i][2]:>14s} {rows[i][1] / 1000:8.2f} {100 * rows[i][1] / tot:6.1f}%\n")
            f.write(f"{'TOTAL (serialised)':58s} {'':>14s} {tot / 1000:8.2f}\n")
        print(open(os.path.join(out_dir, f"{tag}_step_breakdown.txt")).read())


def report(rep, name):
    path = os.path.join(go, rep)
    if not os.path.isfile(path):
        return
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    if len(rows) < 3:
        return
    h, units = rows[0], rows[1]
    ix = {k: h.index(k) for k in KEYS if k in h}
    with open(os.path.join(out_dir, f"{tag}_{name}_ncu_full.csv"), "w") as f:
        f.write(f"# ncu --set full --clock-control none --import-source on ({rep}); one row per captured launch\n")
        f.write("kernel,grid," + ",".join(f"{k} [{units[i]}]" for k, i in ix.items()) + "\n")
        for r in rows[2:]:
            f.write(f"\"{short(r[h.index('Kernel Name')])}\",\"{r[h.index('Grid Size')]}\"," + ",".join(r[i] for i in ix.values()) + "\n")
            try:
                scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
                rd, wr = ix["dram__bytes_read.sum"], ix["dram__bytes_write.sum"]
                b = float(r[rd]) * scale[units[rd]] + float(r[wr]) * scale[units[wr]]
                TRAFFIC[f"{short(r[h.index('Kernel Name')])} {r[h.index('Grid Size')]}"] = round(b)
            except (KeyError, ValueError):
                pass
    print(open(os.path.join(out_dir, f"{tag}_{name}_ncu_full.csv")).read()[:3000])


launches()
report("prof_gemm.ncu-rep", "gemm")
report("prof_extract.ncu-rep", "extract")
report("prof_elem.ncu-rep", "elementwise")
if TRAFFIC:
    import json
    # bench.py's kernel classes -> the captured launch that class consists of (BASELINE configs[1] shapes)
    classes = {"gemm_forward_layer1": ("gemm_tf32_kernel<256, 0, 0, 1>", "(32, 1, 4)"), "gemm_dW_layer1": ("gemm_tf32_kernel<256, 1, 1, 1>", "(2, 17, 4)"),
               "gemm_forward_tail": ("tail_fused_kernel", "(32, 1, 1)"), "gather_rows": ("gather_rows_kernel", "(296, 1, 1)"),
               "clip_update": ("clip_update_kernel", ""), "grad_finalize": ("grad_finalize_kernel", ""),
               "extract_conv_relu_pool": ("extract_tc_kernel", "")}
    by_class = {}
    for cls, (kern, grid) in classes.items():
        hits = [v for k, v in TRAFFIC.items() if kern in k and (not grid or k.endswith(grid))]
        if hits:
            by_class[cls] = hits[-1]
    with open(os.path.join(out_dir, f"{tag}_traffic.json"), "w") as f:
        json.dump({"source": "dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full captures of bench.py (cold L2)",
                   "bytes_per_launch": TRAFFIC, "c1": by_class}, f, indent=1, sort_keys=True)
