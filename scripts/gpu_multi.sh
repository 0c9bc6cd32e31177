#!/bin/bash
# N GPUs of one box: world-N parity tests (both exchange modes), then the bench at N (fused and NCCL)
N=${1:-2}
mkdir -p gpurun_out
UST_TEST_WORLD=$N timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -3
for ex in ${EXCHANGES:-fused nccl}; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 \
     bench.py --gpus $N --steps 40 --warmup 5 --quick --exchange $ex 2>&1 | grep '^{' | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print('N=$N $ex: %.1f us/step, %.3g nodes/s' % (d['ms_per_step']*1e3, d['value']))"
done
