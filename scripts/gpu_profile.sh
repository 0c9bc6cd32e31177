#!/bin/bash
# Round artefacts: bench line, ncu launch list of the same command, one full capture of the dominant kernel.
mkdir -p gpurun_out
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.log 2>> gpurun_out/bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 6 --warmup 3 --e2e-steps 2 > gpurun_out/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ust_fused -s 3 -c 2 -f -o gpurun_out/prof \
   python bench.py --steps 6 --warmup 3 --quick > gpurun_out/ncu_full.log 2>&1
UST_STAMPS=296 python bench.py --steps 30 --warmup 5 --quick 2>&1 | grep stamps > gpurun_out/stamps.log
cat gpurun_out/bench.log gpurun_out/bench_reference.log | cut -c1-300
# C4 (pod lists): launch list at full size, one full capture of the pod-summary kernel
timeout 900 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:ust_ -c 6 --csv \
   --log-file gpurun_out/pods_launches.csv python bench.py --steps 3 --warmup 3 --quick --pods > gpurun_out/pods_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ust_pod_summary -s 2 -c 1 -f -o gpurun_out/podprof \
   python bench.py --steps 3 --warmup 3 --quick --pods --nodes 4000000 > gpurun_out/pods_ncu_full.log 2>&1
timeout 600 python bench.py --steps 30 --warmup 5 --quick --pods 2>&1 | tail -1 > gpurun_out/bench_pods.log; cut -c1-200 gpurun_out/bench_pods.log
