import csv, collections, sys
path = sys.argv[1]
rows = list(csv.reader(open(path)))
hdr = rows[1]; col = {h: i for i, h in enumerate(hdr)}
reg = collections.defaultdict(lambda: [0, 0, 0])
tot_s = 0; tot_i = 0
seen = set()
for r in rows[2:]:
    if len(r) < len(hdr) or not r[col['Instructions Executed']].isdigit(): continue
    if r[col['Address']] in seen: continue
    seen.add(r[col['Address']])
    n = int(r[col['Instructions Executed']]); s = int(r[col['# Samples']] or 0)
    e = reg[n]; e[0] += 1; e[1] += n; e[2] += s
    tot_s += s; tot_i += n
print('total instr', tot_i, 'samples', tot_s)
for n, e in sorted(reg.items(), key=lambda kv: -kv[1][2])[:16]:
    print(f'exec {n:9d} static {e[0]:4d} dyn {100*e[1]/tot_i:5.1f}%  samples {100*e[2]/tot_s:5.1f}%  (samples per instr-issue {e[2]/max(1,e[1])*1000:.2f}e-3)')
