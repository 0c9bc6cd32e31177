#!/bin/bash
# N GPUs of one box: the bench line only (fused exchange), optionally with extra bench flags
N=${1:-8}; shift
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 \
   bench.py --gpus $N --steps 50 --warmup 5 "$@" > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench exit $?"
tail -3 gpurun_out/bench_n$N.err | cut -c1-300
python -c "
import json
d=json.loads([l for l in open('gpurun_out/bench_n$N.json') if l.startswith('{')][-1])
print('N=$N fused: %.2f us/step, %.4g nodes/s' % (d['ms_per_step']*1e3, d['value']), 'parity', d.get('parity_checked'), d.get('mismatches'), 'exchange_us', d.get('exchange_us'), 'local', d.get('local_ms_per_step'), 'e2e ms', d.get('e2e',{}).get('ms_per_step'))"
