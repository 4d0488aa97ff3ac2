#!/bin/bash
# round 2, call 2: full suite on the refactored kernels; PT barrier restructure; N-D engine models; launch lists of cfg4/cfg5
mkdir -p gpurun_out; O=gpurun_out; T=r02b
echo "== full GPU suite"; timeout 1200 python -m pytest tests -m gpu -x -q > $O/${T}_gpu_suite.log 2>&1; echo "rc=$?"; tail -8 $O/${T}_gpu_suite.log
echo "== cfg2"; timeout 200 python bench.py --no-cpu-baseline > $O/${T}_bench_cfg2.json 2> $O/${T}_bench_cfg2.err; echo "rc=$?"
echo "== cfg3"; timeout 200 python bench.py --workload cfg3 --no-cpu-baseline > $O/${T}_bench_cfg3.json 2> $O/${T}_bench_cfg3.err; echo "rc=$?"
python tools/show_bench.py $O/${T}_bench_cfg2.json $O/${T}_bench_cfg3.json 2>/dev/null
echo "== cfg4 / cfg5 (engine N-D models)"
timeout 300 python bench.py --workload cfg4 > $O/${T}_bench_cfg4.json 2> $O/${T}_bench_cfg4.err; echo "rc=$?"
timeout 300 python bench.py --workload cfg5 --size 512 > $O/${T}_bench_cfg5.json 2> $O/${T}_bench_cfg5.err; echo "rc=$?"
RKS_COEF_STORAGE=arrays timeout 300 python bench.py --workload cfg5 --size 512 > $O/${T}_bench_cfg5_arrays.json 2> $O/${T}_bench_cfg5_arrays.err; echo "rc=$?"
for f in cfg4 cfg5 cfg5_arrays; do python -c "import json,sys; d=json.load(open('$O/${T}_bench_$f.json')); print('$f', d['ms_per_step'], d['value'], d['steps'])"; tail -n 3 $O/${T}_bench_$f.err; done
echo "== ncu launch lists"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/${T}_launches_cfg4.csv python bench.py --workload cfg4 > $O/${T}_ncu_cfg4.log 2>&1; echo "rc=$?"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file $O/${T}_launches_cfg5.csv python bench.py --workload cfg5 --size 512 > $O/${T}_ncu_cfg5.log 2>&1; echo "rc=$?"
python tools/summarize_launches.py $O/${T}_launches_cfg4.csv 2>/dev/null | head -30
python tools/summarize_launches.py $O/${T}_launches_cfg5.csv 2>/dev/null | head -30
