#!/bin/bash
# N GPUs of one box: world-N parity tests (both exchange modes; log kept), the full bench line at N (fused exchange:
# parity against the unsharded oracle, exchange_us), a quick line for the NCCL mode
N=${1:-2}
mkdir -p gpurun_out
UST_TEST_WORLD=$N timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -rA > gpurun_out/pytest_multi_n$N.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_multi_n$N.log
tail -6 gpurun_out/pytest_multi_n$N.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 \
   bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench exit $?"
tail -3 gpurun_out/bench_n$N.err | cut -c1-300
python -c "
import json
d=json.loads([l for l in open('gpurun_out/bench_n$N.json') if l.startswith('{')][-1])
print('N=$N fused: %.2f us/step, %.4g nodes/s' % (d['ms_per_step']*1e3, d['value']), 'parity', d.get('parity_checked'), d.get('mismatches'), 'exchange_us', d.get('exchange_us'), 'local', d.get('local_ms_per_step'), 'e2e ms', d['e2e']['ms_per_step'], d.get('parity'))"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29613 \
   bench.py --gpus $N --steps 40 --warmup 5 --quick --exchange nccl 2>&1 | grep '^{' | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print('N=$N nccl: %.1f us/step, %.3g nodes/s' % (d['ms_per_step']*1e3, d['value']))"
