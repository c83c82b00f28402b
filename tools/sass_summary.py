#!/usr/bin/env python
"""Static SASS evidence per kernel of libmdf_b200.so (cuobjdump -sass): counts of the mnemonics that prove the tile movement
is TMA (UTMALDG), the synchronisation is mbarrier based (SYNCS), the math is Blackwell packed fp32 (FFMA2 / FMUL2 / FADD2),
plus LDS.128 / RED / MUFU.  No tensor-core instruction (UTC*MMA / HMMA) is expected: the path is a gather + reduction.
    python tools/sass_summary.py [lib.so] > profiles/r02_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "mdf_net_b200", "libmdf_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
WANT = ["UTMALDG", "UBLKCP", "SYNCS", "FFMA2", "FMUL2", "FADD2", "LDS.128", "MUFU", "RED", "ATOMS", "STG", "LDG", "UTCHMMA", "UTCIMMA", "HMMA"]
kern, counts, arch = None, {}, set()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r"\(.*", "", kern)[:110]
        counts.setdefault(kern, collections.Counter())
        continue
    m = re.search(r"arch = (sm_\w+)", line)
    if m:
        arch.add(m.group(1))
    if kern and re.search(r"/\*[0-9a-f]{4,6}\*/", line):
        counts[kern]["total"] += 1
        for w in WANT:
            if re.search(r"\b" + re.escape(w), line):
                counts[kern][w] += 1
print(f"# {os.path.relpath(lib, ROOT)}: architectures {sorted(arch)}; static SASS instruction counts per kernel")
print("# kernel | total | " + " | ".join(WANT))
tot = collections.Counter()
for k, c in sorted(counts.items()):
    print(f"{k} | {c['total']} | " + " | ".join(str(c[w]) for w in WANT))
    tot.update(c)
print(f"ALL KERNELS | {tot['total']} | " + " | ".join(str(tot[w]) for w in WANT))
