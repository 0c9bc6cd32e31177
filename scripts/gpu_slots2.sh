#!/bin/bash
# wrong-speculation cases in detail: hinted and unhinted, with per-CTA stamps
for env in "" "UST_NO_HINT=1"; do
for args in "--maxpar 0 --maxunav 30%" "--maxpar 1000000 --maxunav 100%"; do
  echo "== $env $args"
  env $env UST_STAMPS=296 timeout 300 python bench.py --steps 50 --warmup 5 --quick $args 2>&1 | python -c "
import json,sys
for l in sys.stdin:
    if not l.startswith('{'):
        print(l.strip()); continue
    d=json.loads(l); print('ms_per_step', d['ms_per_step'], 'frac', d['roofline']['frac'])"
done
done
