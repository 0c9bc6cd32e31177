#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "c4 or C4 or pod or vector" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 900 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum --clock-control none -k regex:ust_ -c 6 --csv --log-file gpurun_out/pods_launches.csv \
   python bench.py --steps 3 --warmup 3 --quick --pods > gpurun_out/pods_ncu.log 2>&1
grep ust_ gpurun_out/pods_launches.csv | awk -F, '{print $5, $(NF-2), $NF}' | tail -6
timeout 600 python bench.py --steps 30 --warmup 5 --quick --pods 2>&1 | cut -c1-120 | tail -1
