#!/bin/bash
# round 2, call 36 (1 GPU): sqrt(1/2) of the pi/4 rotations folded into the next butterfly's additions (main) vs multiplied first
mkdir -p gpurun_out; O=gpurun_out; T=r02ag
for v in main nofold main2 nofold2; do
  case $v in main*) unset RKS_LIB;; *) export RKS_LIB=$PWD/rkstiff_b200/variants/nofold.so;; esac
  timeout 150 python bench.py --workload cfg2 --no-cpu-baseline > $O/${T}_cfg2_$v.json 2> $O/${T}_cfg2_$v.err; echo "$v cfg2 rc=$?"
  timeout 150 python bench.py --workload cfg3 --no-cpu-baseline > $O/${T}_cfg3_$v.json 2> $O/${T}_cfg3_$v.err; echo "$v cfg3 rc=$?"
done
unset RKS_LIB
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "fixed_step or pretransformed" > $O/${T}_tests.log 2>&1; echo "tests rc=$?"; tail -1 $O/${T}_tests.log
python - <<'PY'
import json, glob
for p in sorted(glob.glob("gpurun_out/r02ag_cfg*.json")):
    try:
        d = json.load(open(p)); print(p.split("r02ag_")[1], "ms/step %.3f value %.3e" % (d["ms_per_step"], d["value"]), {k[:8]: round(v["us"],1) for k, v in d["roofline"]["kernels"].items() if "nl" in k}, d["clocks"]["reasons"])
    except Exception as e: print(p, "no line", e)
PY
