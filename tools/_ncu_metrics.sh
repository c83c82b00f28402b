M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__warps_eligible.avg.per_cycle_active
for a in $@; do
  ncu --metrics $M --clock-control none -k regex:cost_volume_ --csv python tools/run_cv.py $a dtu_1600x1152_n5 scene 1 2>/dev/null | grep '^"' > gpurun_out/m_$a.csv
done
