#!/bin/bash
# compute-sanitizer over a representative slice of the GPU parity tests (memcheck + racecheck + synccheck)
mkdir -p gpurun_out
SEL='tests/test_gpu_parity.py -k "reference_vector or random_snapshot or slot_budget or disabled or many_daemonsets"'
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q \
     -k "random_snapshot and (4097 or 8192 or 20000 or 127) or slot_budget_cut_positions and (0 or 1) or many_daemonsets or disabled or c4_pod_lists_sample or speculation_hint and 300000 or long_and_empty and 2 or build_state_uid_join_random and 257 or build_state_vector or delta_updates and 5000 or simulated_rollout and 3000" \
     > gpurun_out/sanitizer_$tool.log 2>&1
  echo "$tool exit $?"; grep -E "ERROR SUMMARY|passed|failed|RACECHECK SUMMARY" gpurun_out/sanitizer_$tool.log | tail -3
done
