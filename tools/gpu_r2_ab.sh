#!/bin/bash
# round 2, call 1: measure the two opt-in K4 variants that round 1 left unmeasured, A/B against the defaults
mkdir -p gpurun_out; O=gpurun_out; T=r02a
echo "== new tests (coefficient storage)"; timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "grid_coefficient or indexed_records or separable_tables or non_separable or recognises or grid_model or non_power or cfg4 or cfg5 or batched_2d" > $O/${T}_new_tests.log 2>&1; echo "rc=$?"; tail -15 $O/${T}_new_tests.log
echo "== opt-in tests"; RKS_TEST_RFFT_HALF=1 RKS_TEST_K4_X2=1 timeout 400 python -m pytest tests -m gpu -q -k "rfft_half or k4_x2" > $O/${T}_optin_tests.log 2>&1; echo "rc=$?"; tail -5 $O/${T}_optin_tests.log
echo "== bench_nl default"; timeout 200 python tools/bench_nl.py > $O/${T}_bench_nl_default.txt 2>&1; cat $O/${T}_bench_nl_default.txt
echo "== bench_nl RFFT_HALF"; RKS_RFFT_HALF=1 timeout 200 python tools/bench_nl.py > $O/${T}_bench_nl_rffthalf.txt 2>&1; cat $O/${T}_bench_nl_rffthalf.txt
echo "== cfg3 default"; timeout 200 python bench.py --workload cfg3 --no-cpu-baseline > $O/${T}_bench_cfg3.json 2> $O/${T}_bench_cfg3.err; echo "rc=$?"
echo "== cfg3 RFFT_HALF"; RKS_RFFT_HALF=1 timeout 200 python bench.py --workload cfg3 --no-cpu-baseline > $O/${T}_bench_cfg3_rffthalf.json 2> $O/${T}_bench_cfg3_rffthalf.err; echo "rc=$?"
echo "== cfg2 default"; timeout 200 python bench.py --no-cpu-baseline > $O/${T}_bench_cfg2.json 2> $O/${T}_bench_cfg2.err; echo "rc=$?"
echo "== cfg2 K4_X2"; RKS_K4_X2=1 timeout 200 python bench.py --no-cpu-baseline > $O/${T}_bench_cfg2_x2.json 2> $O/${T}_bench_cfg2_x2.err; echo "rc=$?"
python tools/show_bench.py $O/${T}_bench_cfg3.json $O/${T}_bench_cfg3_rffthalf.json $O/${T}_bench_cfg2.json $O/${T}_bench_cfg2_x2.json 2>/dev/null
tail -3 $O/${T}_*.err
echo "== cfg4 arrays vs auto"; RKS_COEF_STORAGE=arrays timeout 300 python bench.py --workload cfg4 > $O/${T}_bench_cfg4_arrays.json 2> $O/${T}_bench_cfg4_arrays.err; echo "rc=$?"
timeout 300 python bench.py --workload cfg4 > $O/${T}_bench_cfg4_auto.json 2> $O/${T}_bench_cfg4_auto.err; echo "rc=$?"
echo "== cfg5 512 arrays vs auto"; RKS_COEF_STORAGE=arrays timeout 300 python bench.py --workload cfg5 --size 512 > $O/${T}_bench_cfg5_arrays.json 2> $O/${T}_bench_cfg5_arrays.err; echo "rc=$?"
timeout 300 python bench.py --workload cfg5 --size 512 > $O/${T}_bench_cfg5_auto.json 2> $O/${T}_bench_cfg5_auto.err; echo "rc=$?"
for f in cfg4_arrays cfg4_auto cfg5_arrays cfg5_auto; do python -c "import json,sys; d=json.load(open('$O/${T}_bench_$f.json')); print('$f', d['ms_per_step'], d['value'], d['steps'])"; done
echo "== full GPU suite"; timeout 900 python -m pytest tests -m gpu -x -q > $O/${T}_gpu_suite.log 2>&1; echo "rc=$?"; tail -5 $O/${T}_gpu_suite.log
