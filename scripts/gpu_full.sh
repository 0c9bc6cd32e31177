#!/bin/bash
# everything a round needs from one box: full GPU suite, default bench line, reference arm, sanitizer
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench.err; echo "bench exit $?"; cut -c1-400 gpurun_out/bench_n1.json
bash scripts/gpu_sanitize.sh
