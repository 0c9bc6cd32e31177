#!/bin/bash
timeout 600 python -m pytest tests -m gpu -x -q -k "pod or Pod or c4 or C4 or drain or overlapped" 2>&1 | tail -2
timeout 180 python bench.py --steps 20 --warmup 5 --quick --pods 2>&1 | python -c "
import json,sys
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print('   C4 us/step %.2f' % (d['ms_per_step']*1e3))
    elif 'rror' in l: print('   '+l[:300])"
