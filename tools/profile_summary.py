"""Summarise gpurun_out/ ncu artefacts into tracked files under profiles/ (per-round tag).
usage: python tools/profile_summary.py r1c lq:lq_kernel ric:riccati_kernel ro:rollout_kernel"""
import collections
import csv
import json
import subprocess
import sys
from pathlib import Path

import os

ROOT = Path(__file__).resolve().parents[1]
LAUNCH_NOTE = os.environ.get("PROFILE_LAUNCH_NOTE", "bench.py --steps 2 --warmup 1, B=256, 115 nodes")
FULL_NOTE = os.environ.get("PROFILE_FULL_NOTE", "bench.py --batch 64, one launch")
GO, PR = ROOT / "gpurun_out", ROOT / "profiles"
tag = sys.argv[1]
kernels = [a.split(":") for a in sys.argv[2:]]
PR.mkdir(exist_ok=True)
out = [f"# ncu summary {tag}", ""]

# launch list (cold-cache, serialised: compare SHARES, not absolutes)
ll = GO / f"launches_{tag}.csv"
if ll.exists():
    (PR / ll.name).write_bytes(ll.read_bytes())
    rows = [r for r in csv.reader(open(ll)) if len(r) > 5]
    h = rows[0]
    ki, vi = h.index("Kernel Name"), h.index("Metric Value")
    agg = collections.defaultdict(list)
    for r in rows[1:]:
        try:
            agg[r[ki].split("(")[0]].append(float(r[vi].replace(",", "")))
        except ValueError:
            pass
    tot = sum(sum(v) for v in agg.values())
    out += [f"## launch list (`profiles/{ll.name}`; {LAUNCH_NOTE})", "",
            "| kernel | launches | total ms | avg ms | share |", "|---|---|---|---|---|"]
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        out.append(f"| {k} | {len(v)} | {sum(v)/1e6:.3f} | {sum(v)/len(v)/1e6:.3f} | {sum(v)/tot:.3f} |")
    out.append("")

WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio"]
for short, name in kernels:
    rep = GO / f"prof_{short}_{tag}.ncu-rep"
    rawcsv = GO / f"ncuraw_{short}_{tag}.csv"   # `ncu -i <rep> --page raw --csv` run on the GPU box (the .ncu-rep files are too big to bring back)
    if rawcsv.exists():
        raw = rawcsv.read_text()
    elif rep.exists():
        raw = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    else:
        continue
    rows = list(csv.reader(raw.splitlines()))
    h, units, v = rows[0], rows[1], rows[-1]
    out += [f"## {name}  (`ncu --set full`, {FULL_NOTE})", "", "| metric | unit | value |", "|---|---|---|"]
    vals = {}
    for n in WANT:
        if n in h:
            i = h.index(n)
            vals[n] = v[i]
            out.append(f"| {n} | {units[i]} | {v[i]} |")
    out.append("")
    rawd = {n: v[i] for i, n in enumerate(h)}
    # unit-normalised copies of the byte counters (ncu picks Kbyte / Mbyte / Gbyte per value)
    mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    for n in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        if n in h:
            i = h.index(n)
            rawd[n.replace("__bytes_", "_bytes_").replace(".sum", "") + "_B"] = float(v[i].replace(",", "")) * mult.get(units[i], 1.0)
    tmult = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}
    if "gpu__time_duration.sum" in h:
        i = h.index("gpu__time_duration.sum")
        rawd["gpu_time_ms"] = float(v[i].replace(",", "")) * tmult.get(units[i], 1.0)
    rawd["batch"] = int(os.environ.get("PROFILE_BATCH", "64"))   # instances per GPU of the captured launch (bench.py scales `traffic` by it)
    (PR / f"ncu_{short}_{tag}_raw.json").write_text(json.dumps(rawd, indent=0))
bench = GO / f"bench_{tag}.json"
if bench.exists():
    line = bench.read_text().strip().splitlines()[-1]
    (PR / f"bench_{tag}.json").write_text(line + "\n")
    out += [f"## bench line (`profiles/bench_{tag}.json`, python bench.py --steps 10 --warmup 3, not under a profiler)", "", "```json", line, "```", ""]
(PR / f"summary_{tag}.md").write_text("\n".join(out))
print("\n".join(out)[:6000])
