#!/bin/bash
# compute-sanitizer passes over small instances of every kernel family (memcheck + racecheck + synccheck)
mkdir -p gpurun_out
SEL='fused_uux_matches_numpy and (1-16 or 3-512 or 3-1024 or 3-4096 or 40-8192) or fused_nls_matches_numpy and (5-512 or 5-2048 or 150-8192) or test_adaptive_dt_sequence_and_final_state and kdv and fused or test_fixed_step_parity_per_step and ksb and fused or cfg2b_batched and (ETD35-nls512 or IF34-kdv256) or axis_fft_kernels and (32 or 512 or 4096) or fused_nls_matches_numpy and (5-64 or 150-256) or fused_uux_matches_numpy and (3-128 or 301-256) or fused_new_models and (8192 or 128) or diagonalized_dense and IF34 or cfg5_nls_3d'
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$SEL" > gpurun_out/sanitizer_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|passed|failed|RACECHECK SUMMARY" gpurun_out/sanitizer_$tool.log | tail -3
done
