#!/bin/bash
# bench + ncu launch list + one full capture of the dominant kernel. Results in gpurun_out/.
set -x
mkdir -p gpurun_out
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?"
cat gpurun_out/bench.log
timeout 600 python bench.py --steps 50 --warmup 5 --sets 2 > gpurun_out/bench_sets2.log 2>> gpurun_out/bench.err
cat gpurun_out/bench_sets2.log | cut -c1-400
# every launch with its device time (cold-cache, serialised: compare shares)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 6 --warmup 3 --e2e-steps 1 > gpurun_out/ncu_launches.log 2>&1
tail -5 gpurun_out/launches.csv
# the top kernel, full set, 3 launches after warm-up
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ust_fused -s 3 -c 3 -f -o gpurun_out/prof \
   python bench.py --steps 6 --warmup 3 --e2e-steps 1 > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out/
