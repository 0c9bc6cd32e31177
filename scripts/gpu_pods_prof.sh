#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:ust_ -c 8 --csv --log-file gpurun_out/pods_launches.csv \
   python bench.py --steps 3 --warmup 3 --quick --pods --nodes ${NODES:-4000000} > gpurun_out/pods_ncu.log 2>&1
grep ust_ gpurun_out/pods_launches.csv | awk -F, '{print $5, $NF}' | tail -6
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ust_pod_summary -s 2 -c 1 -f -o gpurun_out/podprof \
   python bench.py --steps 3 --warmup 3 --quick --pods --nodes ${NODES:-4000000} > gpurun_out/pods_ncu_full.log 2>&1
ls -la gpurun_out/podprof.ncu-rep
