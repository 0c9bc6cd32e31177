#!/bin/bash
# tuning: time each build_variants/*.so (device-resident timing only)
mkdir -p gpurun_out; : > gpurun_out/variants.log
for so in build_variants/*.so "$@"; do
  [ -f "$so" ] || continue
  echo "== $so" | tee -a gpurun_out/variants.log
  UST_LIB=$PWD/$so timeout 300 python bench.py --steps 40 --warmup 5 --quick 2>&1 | tail -1 | cut -c1-330 | tee -a gpurun_out/variants.log
done
