#!/bin/bash
mkdir -p gpurun_out
q() {
  timeout 180 env "$@" python bench.py --steps 50 --warmup 5 --quick $ARGS 2>&1 | python -c "
import json,sys
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print('   us/step %.2f frac %.3f' % (d['ms_per_step']*1e3, d['roofline']['frac']))
    elif l.startswith('stamps') or l.startswith('counters') or 'rror' in l: print('   '+l[:500])"
}
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
echo "== default (overlapped)"; ARGS="" q X=1
echo "== strict order"; ARGS="" q UST_OVERLAP=0
echo "== cut hinted"; ARGS="--maxpar 0 --maxunav 30%" q X=1
echo "== cut no hint"; ARGS="--maxpar 0 --maxunav 30%" q UST_NO_HINT=1 UST_STAMPS=148
echo "== 100k"; ARGS="--nodes 100000" q UST_STAMPS=148
echo "== C2"; ARGS="--nodes 1000000 --sets 32" q X=1
echo "== 100k strict"; ARGS="--nodes 100000 --sets 64" q UST_OVERLAP=0
