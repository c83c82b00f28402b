import csv, sys
for a in sys.argv[1:]:
    rows = list(csv.DictReader(open(f'gpurun_out/m_{a}.csv')))
    by = {}
    for r in rows:
        by.setdefault(r['ID'], {'k': r['Kernel Name'][5:70]})[r['Metric Name']] = float(r['Metric Value'].replace(',', ''))
    for i, d in by.items():
        print(f"algo {a:>4} {d['k']:66s} {d['gpu__time_duration.sum'] / 1e3:7.1f} us  inst {d['smsp__inst_executed.sum'] / 1e6:6.1f}M  issue {d['smsp__issue_active.avg.pct_of_peak_sustained_active']:5.1f}%  "
              f"wf {d['l1tex__data_pipe_lsu_wavefronts_mem_shared.sum'] / 1e6:5.1f}M  warps {d['sm__warps_active.avg.pct_of_peak_sustained_active']:4.1f}%  eligible {d['smsp__warps_eligible.avg.per_cycle_active']:.2f}")
