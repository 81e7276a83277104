#!/usr/bin/env python
"""Summarise one `ncu --set full` capture of the render megakernel for profiles/.

  NCU_WORKLOAD=C2 NCU_SAMPLES=48000000 python scripts/ncu_summary.py gpurun_out/prof_render.ncu-rep profiles/r02/v13 [librtiow_b200.so kernel-substring]

Writes <prefix>_render_kernel_ncu_full.json (selected raw metrics), <prefix>_by_line.txt (executed
warp instructions / stall samples per source function and line, when the .so is given) and
refreshes the NCU_WORKLOAD entry of profiles/ncu_summary.json (per-sample DRAM bytes and warp instructions, issue-slot
utilisation, lanes per instruction: read by bench.py's `roofline`; NCU_SAMPLES = pixel-samples of the captured launch).
"""
import csv
import io
import json
import os
import subprocess
import sys

rep, prefix = sys.argv[1], sys.argv[2]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
keep = ("Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.avg.per_cycle_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__inst_executed.sum", "sm__cycles_active.avg", "sm__cycles_elapsed.avg",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_cbu.avg.pct_of_peak_sustained_active", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__warps_eligible.avg.per_cycle_active")
out = {}
for i, h in enumerate(hdr):
    if h in keep or ("issue_stalled" in h and h.endswith("per_issue_active.ratio")):
        out[h] = {"unit": units[i], "value": vals[i]}
os.makedirs(os.path.dirname(prefix), exist_ok=True)
with open(prefix + "_render_kernel_ncu_full.json", "w") as f:
    json.dump(out, f, indent=1)


def to_bytes(m):
    v, u = float(out[m]["value"]), out[m]["unit"].lower()
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[u]


dram = to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum")
# profiles/ncu_summary.json feeds bench.py's `roofline`: one entry per BASELINE workload, refreshed by a capture of that workload
wl, n_samples = os.environ.get("NCU_WORKLOAD"), float(os.environ.get("NCU_SAMPLES", "0"))
if wl and n_samples > 0:
    path = os.path.join(ROOT, "profiles", "ncu_summary.json")
    table = json.load(open(path)) if os.path.exists(path) else {}
    fnum = lambda m: float(out[m]["value"].replace(",", ""))  # noqa: E731
    table[wl] = {"kernel": out["Kernel Name"]["value"], "samples_in_capture": n_samples,
                 "dram_bytes_per_sample": dram / n_samples,
                 "warp_inst_per_sample": fnum("smsp__inst_executed.sum") / n_samples,
                 "issue_slot_frac": fnum("smsp__issue_active.avg.pct_of_peak_sustained_active") / 100.0,
                 "lanes_per_inst": fnum("smsp__thread_inst_executed_per_inst_executed.ratio"),
                 "kernel_ms_under_ncu": fnum("gpu__time_duration.sum"),
                 "source": os.path.relpath(prefix + "_render_kernel_ncu_full.json", ROOT)}
    with open(path, "w") as f:
        json.dump(table, f, indent=1)
if len(sys.argv) > 4:
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    tmp = prefix + "_source.tmp.csv"
    with open(tmp, "w") as f:
        f.write(src)
    txt = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_by_line.py"), tmp, sys.argv[3], sys.argv[4], "40"],
                         capture_output=True, text=True).stdout
    hot = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_hot.py"), tmp, "25"], capture_output=True, text=True).stdout
    with open(prefix + "_by_line.txt", "w") as f:
        f.write(txt + "\n" + hot)
    os.remove(tmp)
print(json.dumps({k: v["value"] for k, v in out.items() if "stalled" not in k}, indent=1))
