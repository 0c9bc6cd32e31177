#!/bin/bash
# Parity tests, then the kernel-only bench under the default policy and under policies whose slot budget cuts
# inside the array (the speculation is wrong on the first call, hinted on the following ones).
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
for args in "" "--maxpar 0 --maxunav 30%" "--maxpar 0 --maxunav 10%" "--maxpar 50000 --maxunav 100%"; do
  echo "== $args"
  timeout 300 python bench.py --steps 50 --warmup 5 --quick $args 2> gpurun_out/bench.err | python -c "
import json,sys
for l in sys.stdin:
    if not l.startswith('{'):
        print(l.strip()); continue
    d=json.loads(l); print('ms_per_step', d['ms_per_step'], 'frac', d['roofline']['frac'])"
  tail -2 gpurun_out/bench.err
done
