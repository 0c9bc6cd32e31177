#!/bin/bash
# parity suite + full bench line
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -12 gpurun_out/pytest_gpu.log
T0=$(date +%s); timeout 900 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench exit $? in $(( $(date +%s) - T0 )) s"
tail -3 gpurun_out/bench_full.err | cut -c1-400
python -c "
import json;d=json.load(open('gpurun_out/bench_full.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','verified_vs_oracle')}, d['roofline']['frac']); print(d['e2e']); print(d.get('e2e_delta'))
b=d['by_config']; print('C3_cut', b['C3_cut']['first_call_us'], b['C3_cut']['steady_us'], b['C3_cut']['verified_vs_oracle']); print('C2', b['C2']['us_per_call'], 'small', {k:v['us_per_call'] for k,v in b['small'].items()}); print('C4', b['C4']['us_per_call'], b['C4']['frac'], b['C4']['verified_vs_oracle'])"
