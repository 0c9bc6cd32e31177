#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_full.json') if l.startswith('{')][-1])
print('C3 us', d['ms_per_step']*1e3, 'frac', d['roofline']['frac'], 'e2e ms', d['e2e']['ms_per_step'])
print(json.dumps(d['by_config'], indent=1)[:3500])
print(json.dumps(d.get('e2e_delta'))[:600])
PY
tail -5 gpurun_out/bench_full.err
