#!/bin/bash
# round 2, call 24 (1 GPU): 4096-point axis tile of one column (two / three resident CTAs) against the two-column tile
mkdir -p gpurun_out; O=gpurun_out; T=r02x
for v in main ax4096_c1b2 ax4096_c1b3 main2; do
  if [ $v = main ] || [ $v = main2 ]; then unset RKS_LIB; else export RKS_LIB=$PWD/rkstiff_b200/variants/$v.so; fi
  timeout 150 python bench.py --workload cfg4 --no-cpu-baseline > $O/${T}_cfg4_$v.json 2> $O/${T}_cfg4_$v.err; echo "$v cfg4 rc=$?"
  timeout 300 python -m pytest tests/test_gpu_size_classes.py -x -q -k "4096" > $O/${T}_size_$v.log 2>&1; echo "$v size-class parity rc=$?"; tail -1 $O/${T}_size_$v.log
done
unset RKS_LIB
python - <<'PY'
import json, glob
for p in sorted(glob.glob("gpurun_out/r02x_cfg*.json")):
    try:
        d = json.load(open(p)); print(p.split("r02x_")[1], "ms/step %.3f value %.3e" % (d["ms_per_step"], d["value"]), d["clocks"]["reasons"])
        for k, v in (d["roofline"].get("kernels") or {}).items(): print("      %-60s %7.1f us frac %.3f x%d" % (k[:60], v["us"], v["frac"], v["launches_per_step"]))
    except Exception as e: print(p, "no line", e)
PY
