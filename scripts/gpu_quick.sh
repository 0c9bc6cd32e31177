#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
for i in 1 2; do UST_STAMPS=296 timeout 300 python bench.py --steps 50 --warmup 5 --quick 2>&1 | grep -v "^slowest" | cut -c1-330; done
UST_STAMPS=296 timeout 300 python bench.py --steps 50 --warmup 5 --quick --maxpar 0 --maxunav 30% 2>&1 | grep -v "^slowest\|^stamps" | cut -c1-100
