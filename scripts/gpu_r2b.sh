#!/bin/bash
# Round-2: parity suite, the full bench line, stamp overhead check, ncu launch list + one full capture of the streaming kernel.
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
echo "== parity tests"
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
echo "== quick, no stamps"; ARGS="" q X=1
echo "== quick, stamps"; ARGS="" q UST_STAMPS=148
echo "== cut hinted, no stamps"; ARGS="--maxpar 0 --maxunav 30%" q X=1
echo "== cut no hint, stamps"; ARGS="--maxpar 0 --maxunav 30%" q UST_NO_HINT=1 UST_STAMPS=148
echo "== 100k no stamps"; ARGS="--nodes 100000" q X=1
echo "== full bench"
T0=$(date +%s); timeout 900 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench exit $? in $(( $(date +%s) - T0 )) s"
tail -3 gpurun_out/bench_full.err | cut -c1-300
python -c "import json;d=json.load(open('gpurun_out/bench_full.json'));print({k:d[k] for k in ('value','ms_per_step','gpu_launches','verified_vs_oracle')}, d['roofline']['frac'], d['e2e'], d.get('e2e_delta')); print(json.dumps(d.get('by_config'), indent=1))"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2>> gpurun_out/bench_full.err
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 6 --warmup 3 --e2e-steps 1 --no-by-config > gpurun_out/ncu_launches.log 2>&1
grep -c ust_ gpurun_out/launches.csv
echo "== ncu full capture"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ust_stream -s 3 -c 2 -f -o gpurun_out/prof \
   python bench.py --steps 6 --warmup 3 --quick > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out/ | head -20
