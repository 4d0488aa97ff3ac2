#!/bin/bash
# compute-sanitizer passes over small instances of every kernel family (memcheck + racecheck + synccheck)
mkdir -p gpurun_out
SEL='fused_uux_matches_numpy and (16-1 or 512-3 or 1024-3 or 4096-3 or 8192-3) or fused_nls_matches_numpy and (512-5 or 2048-5 or 8192-5) or test_adaptive_dt_sequence_and_final_state and kdv and fused or test_fixed_step_parity_per_step and ksb and fused or cfg2b_batched and (ETD35-nls512 or IF34-kdv256)'
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$SEL" > gpurun_out/sanitizer_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|passed|failed|RACECHECK SUMMARY" gpurun_out/sanitizer_$tool.log | tail -3
done
