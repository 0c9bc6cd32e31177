#!/bin/bash
q() {
  timeout 180 env "$@" python bench.py --steps 50 --warmup 5 --quick $ARGS 2>&1 | python -c "
import json,sys
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print('   us/step %.2f frac %.3f' % (d['ms_per_step']*1e3, d['roofline']['frac']))
    elif l.startswith('stamps') or l.startswith('counters') or 'rror' in l: print('   '+l[:600])"
}
echo "== overlapped, stamps"; ARGS="" q UST_STAMPS=148
echo "== overlapped"; ARGS="" q X=1
echo "== strict stamps"; ARGS="" q UST_OVERLAP=0 UST_STAMPS=148
echo "== strict"; ARGS="" q UST_OVERLAP=0
echo "== cut hinted"; ARGS="--maxpar 2000000 --maxunav 30%" q X=1
echo "== cut no hint"; ARGS="--maxpar 2000000 --maxunav 30%" q UST_NO_HINT=1
echo "== 100k"; ARGS="--nodes 100000" q X=1
echo "== C2"; ARGS="--nodes 1000000" q X=1
