#!/bin/bash
# round 2, call 25 (1 GPU): 4096-point axis kernel with the next tile staged by cp.async (main) against the plain kernel
mkdir -p gpurun_out; O=gpurun_out; T=r02y
timeout 400 python -m pytest tests/test_gpu_size_classes.py tests/test_gpu_parity.py -x -q -k "4096 or axis or grid or nd" > $O/${T}_axis_tests.log 2>&1; echo "axis tests rc=$?"; tail -2 $O/${T}_axis_tests.log
for v in main ax4096_nostage main2 ax4096_nostage2; do
  case $v in main*) unset RKS_LIB;; *) export RKS_LIB=$PWD/rkstiff_b200/variants/ax4096_nostage.so;; esac
  timeout 150 python bench.py --workload cfg4 --no-cpu-baseline > $O/${T}_cfg4_$v.json 2> $O/${T}_cfg4_$v.err; echo "$v cfg4 rc=$?"
done
unset RKS_LIB
timeout 100 python tools/bench_axis.py > $O/${T}_bench_axis_main.txt 2>&1; echo "bench_axis rc=$?"
RKS_LIB=$PWD/rkstiff_b200/variants/ax4096_nostage.so timeout 100 python tools/bench_axis.py > $O/${T}_bench_axis_nostage.txt 2>&1
python - <<'PY'
import json, glob
for p in sorted(glob.glob("gpurun_out/r02y_cfg*.json")):
    try:
        d = json.load(open(p)); print(p.split("r02y_")[1], "ms/step %.3f value %.3e" % (d["ms_per_step"], d["value"]), d["clocks"]["reasons"])
    except Exception as e: print(p, "no line", e)
PY
grep -h 4096 $O/${T}_bench_axis_main.txt $O/${T}_bench_axis_nostage.txt
