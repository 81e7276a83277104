"""Summarise an `ncu --page source --csv` dump: instruction totals, hottest SASS by stall samples."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
col = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
def num(r, name):
    try: return float(r[col[name]])
    except (ValueError, IndexError): return 0.0
tot_samples = sum(num(r, "# Samples") for r in data)
tot_inst = sum(num(r, "Instructions Executed") for r in data)
tot_thr = sum(num(r, "Thread Instructions Executed") for r in data)
print(f"SASS instructions: {len(data)}  samples: {tot_samples:.0f}  warp-inst: {tot_inst:.3e}  thread-inst: {tot_thr:.3e}  avg threads/inst: {tot_thr/tot_inst:.2f}")
# opcode histogram weighted by executed warp instructions
from collections import Counter
ops = Counter(); ops_s = Counter()
for r in data:
    sass = r[col["Source"]].strip()
    op = sass.split()[0] if not sass.startswith("@") else sass.split()[1]
    op = op.split(".")[0]
    ops[op] += num(r, "Instructions Executed"); ops_s[op] += num(r, "# Samples")
print("opcode mix (warp-inst %, sample %):")
for op, n in ops.most_common(28):
    print(f"  {op:10s} {100*n/tot_inst:6.2f}%  {100*ops_s[op]/tot_samples:6.2f}%")
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
print(f"top {n} instructions by samples:")
for r in sorted(data, key=lambda r: -num(r, "# Samples"))[:n]:
    print(f"  {r[col['Address']][-5:]} {num(r,'# Samples'):8.0f} {100*num(r,'# Samples')/tot_samples:5.2f}%  exec {num(r,'Instructions Executed'):.2e} thr/inst {num(r,'Avg. Threads Executed'):5.1f}  {r[col['Source']].strip()[:90]}")
