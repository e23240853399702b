#!/bin/bash
set -u
mkdir -p gpurun_out
python -c 'import torch' 2> /dev/null
NCU="ncu --set full --clock-control none --import-source on"
timeout 300 $NCU -k regex:splitk_reduce -s 3 -c 1 -f -o gpurun_out/prof_reduce_z16 python tools/prof_one.py wgrad 256 64 100352 > gpurun_out/ncu_reduce_z16.log 2>&1
timeout 300 $NCU -k regex:splitk_reduce -s 3 -c 1 -f -o gpurun_out/prof_reduce_z4 python tools/prof_one.py wgrad 1024 256 6272 > gpurun_out/ncu_reduce_z4.log 2>&1
timeout 300 $NCU -k regex:gemm_tcgen05_kernel -s 3 -c 1 -f -o gpurun_out/prof_wgrad_part python tools/prof_one.py wgrad 1024 256 6272 > gpurun_out/ncu_wgrad_part.log 2>&1
for f in prof_reduce_z16 prof_reduce_z4 prof_wgrad_part; do
  ncu -i gpurun_out/$f.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin))
h,u,r=rows[0],rows[1],rows[2]
want=['gpu__time_duration.sum','launch__grid_size','launch__block_size','launch__registers_per_thread','sm__cycles_active.avg','sm__cycles_elapsed.max','dram__bytes_read.sum','lts__t_sector_hit_rate.pct','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio','launch__waves_per_multiprocessor','sm__throughput.avg.pct_of_peak_sustained_elapsed','lts__throughput.avg.pct_of_peak_sustained_elapsed','smsp__inst_executed.sum']
print('== $f')
for w in want:
    if w in h: print(' ', w, r[h.index(w)], u[h.index(w)])
"
done
