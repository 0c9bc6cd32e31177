#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/c4_launches.csv python bench.py --quick --pods --steps 5 --warmup 3 > gpurun_out/c4_l.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/c4_launches.csv')) if len(r)>5]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
for r in rows[1:][-12:]: print(r[ki][:60], r[vi])
PY
timeout 400 ncu --set full --clock-control none --import-source on -k regex:ust_pod_summary --launch-skip 4 --launch-count 1 -o gpurun_out/c4_podsum python bench.py --quick --pods --steps 5 --warmup 3 > gpurun_out/c4_f.log 2>&1
ncu -i gpurun_out/c4_podsum.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]; v=rows[2] if len(rows)>2 else rows[1]
want=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','dram__throughput.avg.pct_of_peak_sustained_elapsed','sm__throughput.avg.pct_of_peak_sustained_elapsed','smsp__inst_executed.sum','l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum','l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum','sm__warps_active.avg.pct_of_peak_sustained_active','lts__t_sectors_srcunit_tex_op_read.sum','smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct','launch__registers_per_thread','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','smsp__issue_active.avg.pct_of_peak_sustained_active']
for w in want:
    if w in h: print(w, rows[1][h.index(w)], v[h.index(w)])
"
