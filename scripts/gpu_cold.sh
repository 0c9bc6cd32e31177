#!/bin/bash
# tuning: wrong-speculation cases per build_variants/cold*.so
for so in build_variants/cold*.so; do
  for env in "UST_NO_HINT=1" "X=1"; do
    r=$(env $env UST_LIB=$PWD/$so timeout 300 python bench.py --steps 40 --warmup 5 --quick --maxpar 0 --maxunav 30% 2>&1 | grep '^{' | python -c "import json,sys; print(json.loads(sys.stdin.read())['ms_per_step']*1e3)")
    echo "$so $env: $r us"
  done
  r=$(UST_LIB=$PWD/$so timeout 300 python bench.py --steps 40 --warmup 5 --quick 2>&1 | grep '^{' | python -c "import json,sys; print(json.loads(sys.stdin.read())['ms_per_step']*1e3)")
  echo "$so default: $r us"
done
