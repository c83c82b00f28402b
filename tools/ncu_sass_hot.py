#!/usr/bin/env python
"""Dynamic SASS statistics from `ncu -i X.ncu-rep --page source --csv --launch-skip k --launch-count 1`:
executed warp-instructions per opcode and the instructions that collect the most stall samples."""
import csv
import collections
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
col = {h: i for i, h in enumerate(hdr)}
ops = collections.Counter()
samples = []
total = 0
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    src = r[col["Source"]].strip()
    if not r[col["Instructions Executed"]].isdigit():
        continue      # repeated header (several functions in one export)
    n = int(r[col["Instructions Executed"]] or 0)
    op = src.split()[0] if not src.startswith("@") else src.split()[1]
    ops[op.split(".")[0]] += n
    total += n
    samples.append((int(r[col["# Samples"]] or 0), n, src, {k: int(r[col[k]] or 0) for k in
                   ("stall_long_sb", "stall_short_sb", "stall_barrier", "stall_wait", "stall_mio", "stall_math", "stall_not_selected",
                    "stall_branch_resolving", "stall_lg", "stall_dispatch") if k in col}))
print("total warp instructions", total, " (print name:", rows[0][1][:80], ")")
print("by opcode:", ", ".join(f"{k} {v / total:.1%}" for k, v in ops.most_common(22)))
tot_s = sum(s[0] for s in samples)
print("top stall instructions (samples, % of all, executed, sass, dominant reasons):")
for i, (s, n, src, st) in enumerate(sorted(samples, key=lambda t: -t[0])[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]):
    dom = sorted(st.items(), key=lambda kv: -kv[1])[:2]
    print(f"  {s:6d} {s / tot_s:5.1%} {n:9d}  {src[:60]:60s} {dom}")
