#!/bin/bash
for i in 1 2; do
timeout 300 python bench.py --steps 10 --warmup 3 --e2e-steps 10 --no-by-config 2>&1 | python -c "
import json,sys
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print('   e2e ms %.4f  delta %.4f sparse %.4f' % (d['e2e']['ms_per_step'], d['e2e_delta']['ms_per_step'], d['e2e_delta']['sparse_outputs']['ms_per_step']))
    elif 'rror' in l: print('   '+l[:300])"
done
