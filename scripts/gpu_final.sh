#!/bin/bash
# Round-2 final evidence: GPU parity suite, the bench line (both arms), ncu launch list of the bench command, one
# --set full capture of the streaming kernel, compute-sanitizer memcheck over a slice of the parity tests.
mkdir -p gpurun_out
echo "== parity tests"
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
echo "== full bench"
T0=$(date +%s); timeout 900 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench exit $? in $(( $(date +%s) - T0 )) s"
tail -3 gpurun_out/bench_full.err | cut -c1-300
python -c "import json;d=json.loads([l for l in open('gpurun_out/bench_full.json') if l.startswith('{')][-1]);print({k:d[k] for k in ('value','ms_per_step','gpu_launches','verified_vs_oracle')}, d['roofline']['frac'], d['e2e']['ms_per_step'], d['clocks']); print({k:(v.get('us_per_call') or v.get('ms') or v.get('steady_us')) for k,v in d['by_config'].items() if isinstance(v,dict)})"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2>> gpurun_out/bench_full.err; tail -c 400 gpurun_out/bench_reference.json
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 6 --warmup 3 --e2e-steps 1 --no-by-config > gpurun_out/ncu_launches.log 2>&1
grep -c ust_ gpurun_out/launches.csv
echo "== ncu full capture"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ust_stream -s 3 -c 2 -f -o gpurun_out/prof \
   python bench.py --steps 6 --warmup 3 --quick > gpurun_out/ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:ust_verify -s 3 -c 1 -f -o gpurun_out/prof_verify \
   python bench.py --steps 6 --warmup 3 --quick > gpurun_out/ncu_full2.log 2>&1
echo "== C4 launch list + pod kernel capture"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/c4_launches.csv python bench.py --quick --pods --steps 5 --warmup 3 > gpurun_out/c4_l.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:ust_pod_summary --launch-skip 4 --launch-count 1 -f -o gpurun_out/c4_podsum python bench.py --quick --pods --steps 5 --warmup 3 > gpurun_out/c4_f.log 2>&1
echo "== memcheck"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q \
     -k "random_snapshot and (4097 or 8192 or 20000 or 127) or slot_budget_cut_positions and (0 or 1) or many_daemonsets or disabled or speculation_hint and 300000 or long_and_empty and 2 or build_state_vector or delta_updates and 5000 or simulated_rollout and 3000 or sparse or c4_pod_lists_sample" \
     > gpurun_out/sanitizer_memcheck.log 2>&1
echo "memcheck exit $?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitizer_memcheck.log | tail -3
ls -la gpurun_out/ | head -30
