#!/usr/bin/env python
"""Join an `ncu --page source --csv` SASS dump with the line table of the same cubin (nvdisasm -g)
and aggregate executed instructions / stall samples per source line and per source function.

  python scripts/ncu_by_line.py src.csv librtiow_b200.so <kernel-name-substring> [top_n]
"""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

src_csv, so, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, stdout=subprocess.DEVNULL)
dis = "".join(subprocess.run(["nvdisasm", "-g", os.path.join(tmp, f)], capture_output=True, text=True).stdout
              for f in sorted(os.listdir(tmp)) if f.endswith(".cubin"))  # one cubin per translation unit
lines, cur, on = [], None, False
for l in dis.splitlines():
    if l.startswith(".text."):
        on = kname in l
        continue
    if not on:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+[A-Z@]", l):
        lines.append(cur)
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
col = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
assert len(data) == len(lines), (len(data), len(lines), "profile and binary differ")


def num(r, name):
    try:
        return float(r[col[name]])
    except (ValueError, IndexError):
        return 0.0


# function table from the sources (start line of every function-like definition)
funcs = {}
root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "rtiow-rust_b200", "csrc", "device")
for f in os.listdir(root):
    starts = []
    for n, l in enumerate(open(os.path.join(root, f)), 1):
        if not l.startswith(("RT_HD", "__device__", "__global__")):
            continue
        l2 = re.sub(r"__launch_bounds__\([^)]*\)", "", l)
        m = re.search(r"(\w+)\s*\(", l2.replace("operator", "operator_"))
        if m:
            starts.append((n, m.group(1)))
    funcs[f] = starts


def func_of(key):
    if not key or key[0] not in funcs:
        return key[0] if key else "?"
    name = "?"
    for n, fn in funcs[key[0]]:
        if n <= key[1]:
            name = fn
    return f"{key[0]}:{name}"


by_line = collections.defaultdict(lambda: [0.0, 0.0, 0.0])
by_func = collections.defaultdict(lambda: [0.0, 0.0, 0.0])
for r, k in zip(data, lines):
    for agg, kk in ((by_line, k), (by_func, func_of(k))):
        agg[kk][0] += num(r, "Instructions Executed")
        agg[kk][1] += num(r, "Thread Instructions Executed")
        agg[kk][2] += num(r, "# Samples")
tot = sum(v[0] for v in by_func.values())
tots = sum(v[2] for v in by_func.values())
print(f"warp-inst {tot:.3e}  samples {tots:.0f}")
print("--- by function (warp-inst %, lanes/inst, stall-sample %)")
for k, v in sorted(by_func.items(), key=lambda kv: -kv[1][0]):
    if v[0] / tot > 0.002:
        print(f"  {k:48s} {100 * v[0] / tot:6.2f}%  {v[1] / max(v[0], 1):5.1f}  {100 * v[2] / tots:6.2f}%")
print("--- by line")
for k, v in sorted(by_line.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"  {str(k):38s} {100 * v[0] / tot:6.2f}%  {v[1] / max(v[0], 1):5.1f}  {100 * v[2] / tots:6.2f}%")
