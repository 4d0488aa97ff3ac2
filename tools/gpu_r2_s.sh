#!/bin/bash
# round 2, call 19 (1 GPU): resident-CTA variants of the separable / indexed stage kernels; final suite on the final tree
mkdir -p gpurun_out; O=gpurun_out; T=r02s
for v in main st3 st4; do
  if [ $v = main ]; then unset RKS_LIB; else export RKS_LIB=$PWD/rkstiff_b200/variants/$v.so; fi
  timeout 150 python bench.py --workload cfg4 --no-cpu-baseline > $O/${T}_cfg4_$v.json 2> $O/${T}_cfg4_$v.err; echo "$v cfg4 rc=$?"
  timeout 150 python bench.py --workload cfg5 --no-cpu-baseline > $O/${T}_cfg5_$v.json 2> $O/${T}_cfg5_$v.err; echo "$v cfg5 rc=$?"
done
unset RKS_LIB
python - <<'PY'
import json, glob
for p in sorted(glob.glob("gpurun_out/r02s_cfg*.json")):
    try:
        d = json.load(open(p)); print(p.split("r02s_")[1], "ms/step %.3f value %.3e" % (d["ms_per_step"], d["value"]), d.get("clocks"))
    except Exception as e: print(p, "no line", e)
PY
echo "== GPU suite (new tests)"; timeout 600 python -m pytest tests/test_gpu_reference_suite.py tests/test_gpu_size_classes.py -x -q > $O/${T}_suite.log 2>&1; echo "rc=$?"; tail -3 $O/${T}_suite.log
