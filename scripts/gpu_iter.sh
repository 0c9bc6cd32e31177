#!/bin/bash
# Iteration loop on the GPU box: parity tests, bench, launch list, one full ncu capture. Results in gpurun_out/.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?"
tail -3 gpurun_out/bench.err; cat gpurun_out/bench.log
if [ "$1" != "noprof" ]; then
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 6 --warmup 3 --e2e-steps 1 > gpurun_out/ncu_launches.log 2>&1
grep ust_ gpurun_out/launches.csv | tail -3
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ust_fused -s 3 -c 2 -f -o gpurun_out/prof \
   python bench.py --steps 6 --warmup 3 --e2e-steps 1 > gpurun_out/ncu_full.log 2>&1
fi
ls gpurun_out/
