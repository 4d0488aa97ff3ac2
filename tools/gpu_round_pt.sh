#!/bin/bash
# One gpurun call: full GPU suite, smoke, bench lines (default, pair off, other methods/workloads, reference arm), ncu
# launch list and ncu --set full of the newest kernels.  Everything lands in gpurun_out/<prefix>_*.  Each step has its own
# timeout so a hang cannot eat the whole call.
mkdir -p gpurun_out
O=gpurun_out
T=${1:-fin}
echo "== full GPU suite"; timeout 480 python -m pytest tests -m gpu -x -q > $O/${T}_gpu_suite.log 2>&1; echo "rc=$?"; tail -3 $O/${T}_gpu_suite.log
echo "== smoke"; timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/${T}_smoke.log 2>&1; echo "rc=$?"; tail -2 $O/${T}_smoke.log
echo "== bench (default)"; timeout 300 python bench.py > $O/${T}_bench_cfg2.json 2> $O/${T}_bench_cfg2.err; echo "rc=$?"
echo "== bench RKS_PT=0"; RKS_PT=0 timeout 200 python bench.py --no-cpu-baseline > $O/${T}_bench_cfg2_pt_off.json 2> $O/${T}_bench_cfg2_pt_off.err; echo "rc=$?"
echo "== bench IF45DP"; timeout 200 python bench.py --no-cpu-baseline --method IF45DP > $O/${T}_bench_cfg2_if45dp.json 2> $O/${T}_bench_cfg2_if45dp.err; echo "rc=$?"
echo "== bench cfg3"; timeout 200 python bench.py --workload cfg3 > $O/${T}_bench_cfg3.json 2> $O/${T}_bench_cfg3.err; echo "rc=$?"
python tools/show_bench.py $O/${T}_bench_cfg2.json $O/${T}_bench_cfg2_pt_off.json $O/${T}_bench_cfg2_if45dp.json $O/${T}_bench_cfg3.json 2>/dev/null | grep -v "stage"
echo "== reference arm"; timeout 200 python bench.py --impl reference > $O/${T}_bench_reference.json 2> $O/${T}_bench_reference.err; echo "rc=$?"
echo "== ncu launch list"; timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${T}_launches_cfg2.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/${T}_ncu_list.log 2>&1; echo "rc=$?"
echo "== ncu --set full of the newest kernels"; timeout 300 ncu --set full --clock-control none -k regex:"nl_fast_pre_kernel|stage_pre_kernel|norm_kernel" -c 11 -f -o /tmp/${T}_full python tools/prof_kernels.py cfg2 1 > $O/${T}_ncu_full.log 2>&1; echo "rc=$?"
ncu -i /tmp/${T}_full.ncu-rep --page raw --csv > $O/${T}_ncu_full_raw.csv 2>/dev/null; ls -la $O/${T}_ncu_full_raw.csv
echo "== sanitizer (new kernels)"
for tool in memcheck racecheck synccheck; do
  timeout 200 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "pretransformed_pair_equals_plain_pair and (ETD35 or IF4) and (1024 or 8192) or pretransformed_pair_ragged and (ETD35-2048 or IF34) or pretransformed_adaptive and ETD35 and 2048" > $O/${T}_sanitizer_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|passed|failed|RACECHECK SUMMARY" $O/${T}_sanitizer_$tool.log | tail -3
done
