#!/bin/bash
# Round-2 tuning session: parity of the streaming / verification kernels, then timing of the geometry variants.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader > gpurun_out/gpu.txt 2>&1
q() {  # one --quick bench run: prints us/step and the stamps line
  timeout 180 env "$@" python bench.py --steps 50 --warmup 5 --quick $ARGS 2>&1 | python -c "
import json,sys
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print('   us/step %.2f frac %.3f' % (d['ms_per_step']*1e3, d['roofline']['frac']))
    elif l.startswith('stamps') or l.startswith('counters') or 'rror' in l: print('   '+l[:400])"
}
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
echo "== default C3"; ARGS="" q UST_STAMPS=148
echo "== static 100"; ARGS="" q UST_STATIC_PCT=100 UST_STAMPS=148
echo "== parity tests"
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
for so in build_variants/*.so; do echo "== $so"; ARGS="" q UST_LIB=$PWD/$so UST_STAMPS=148; done
echo "== cut mid-array, hinted"; ARGS="--maxpar 0 --maxunav 30%" q UST_STAMPS=148
echo "== cut mid-array, no hint"; ARGS="--maxpar 0 --maxunav 30%" q UST_NO_HINT=1 UST_STAMPS=148
echo "== C2 1M"; ARGS="--nodes 1000000" q UST_STAMPS=148
echo "== 100k"; ARGS="--nodes 100000" q UST_STAMPS=148
echo "== 10k"; ARGS="--nodes 10000" q UST_STAMPS=148
