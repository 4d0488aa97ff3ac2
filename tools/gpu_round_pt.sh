#!/bin/bash
# One gpurun call: parity of the pre-transforming kernel pair, bench with it on and off, ncu evidence, full GPU suite.
# Everything lands in gpurun_out/pt_*.  Each step has its own timeout so a hang cannot eat the whole call.
mkdir -p gpurun_out
O=gpurun_out
echo "== pretransformed tests"; timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k pretransformed > $O/pt_tests.log 2>&1; echo "rc=$?"; tail -3 $O/pt_tests.log
echo "== smoke"; timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/pt_smoke.log 2>&1; echo "rc=$?"; tail -2 $O/pt_smoke.log
echo "== bench (default: pre-transforming pair on)"; timeout 300 python bench.py > $O/pt_bench_cfg2.json 2> $O/pt_bench_cfg2.err; echo "rc=$?"
echo "== bench RKS_PT=0"; RKS_PT=0 timeout 200 python bench.py --no-cpu-baseline > $O/pt_bench_cfg2_off.json 2> $O/pt_bench_cfg2_off.err; echo "rc=$?"
python tools/show_bench.py $O/pt_bench_cfg2.json $O/pt_bench_cfg2_off.json 2>/dev/null | head -40
echo "== ncu launch list"; timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/pt_launches_cfg2.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/pt_ncu_list.log 2>&1; echo "rc=$?"
echo "== ncu --set full of the new kernels"; timeout 240 ncu --set full --clock-control none --import-source on -k regex:"nl_fast_pre_kernel|stage_pre_kernel" -c 7 -f -o $O/pt_full python tools/prof_kernels.py cfg2 1 > $O/pt_ncu_full.log 2>&1; echo "rc=$?"
echo "== reference arm"; timeout 200 python bench.py --impl reference > $O/pt_bench_reference.json 2> $O/pt_bench_reference.err; echo "rc=$?"
echo "== bench cfg3"; timeout 200 python bench.py --workload cfg3 --no-cpu-baseline > $O/pt_bench_cfg3.json 2> $O/pt_bench_cfg3.err; echo "rc=$?"
echo "== full GPU suite"; timeout 480 python -m pytest tests -m gpu -x -q > $O/pt_gpu_suite.log 2>&1; echo "rc=$?"; tail -3 $O/pt_gpu_suite.log
echo "== sanitizer (new kernels)"
for tool in memcheck racecheck; do
  timeout 150 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "pretransformed_pair_equals_plain_pair and ETD35 and (512 or 8192)" > $O/pt_sanitizer_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|passed|failed|RACECHECK SUMMARY" $O/pt_sanitizer_$tool.log | tail -3
done
