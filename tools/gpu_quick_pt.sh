#!/bin/bash
# Short gpurun call: parity of the pre-transforming pair, default bench line, sanitizer on the new kernels.
mkdir -p gpurun_out
O=gpurun_out
T=${1:-q}
echo "== pretransformed tests"; timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k pretransformed > $O/${T}_tests.log 2>&1; echo "rc=$?"; tail -3 $O/${T}_tests.log
echo "== bench"; timeout 300 python bench.py --no-cpu-baseline > $O/${T}_bench_cfg2.json 2> $O/${T}_bench_cfg2.err; echo "rc=$?"; tail -3 $O/${T}_bench_cfg2.err
python tools/show_bench.py $O/${T}_bench_cfg2.json
for tool in memcheck racecheck; do
  timeout 150 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "pretransformed_pair_equals_plain_pair and (ETD35 or IF4) and (512 or 8192)" > $O/${T}_sanitizer_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|passed|failed|RACECHECK SUMMARY" $O/${T}_sanitizer_$tool.log | tail -3
done
