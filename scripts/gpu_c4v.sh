#!/bin/bash
for v in podbase podtick; do
echo "== $v"
UST_LIB=$PWD/build_variants/$v.so timeout 180 python bench.py --steps 20 --warmup 5 --quick --pods 2>&1 | python -c "
import json,sys
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print('   C4 us/step %.2f' % (d['ms_per_step']*1e3))
    elif 'rror' in l: print('   '+l[:300])"
done
