#!/bin/bash
# round 2, call 39 (1 GPU): stagger mode fixed at compile time (no per-mode code copies) vs run-time mode; suite on the variant
mkdir -p gpurun_out; O=gpurun_out; T=r02ah
V=$PWD/rkstiff_b200/variants/modefixed.so
for v in main fixed main2 fixed2; do
  case $v in main*) unset RKS_LIB;; *) export RKS_LIB=$V;; esac
  timeout 100 python bench.py --workload cfg2 --no-cpu-baseline --steps 30 > $O/${T}_cfg2_$v.json 2> $O/${T}_cfg2_$v.err; echo "$v rc=$?"
done
python - <<'PY'
import json, glob
for p in sorted(glob.glob("gpurun_out/r02ah_cfg*.json")):
    try:
        d = json.load(open(p)); print(p.split("r02ah_")[1], "ms/step %.3f" % d["ms_per_step"], {k[:8]: round(v["us"],1) for k, v in d["roofline"]["kernels"].items() if "nl" in k}, d["clocks"]["reasons"])
    except Exception as e: print(p, "no line", e)
PY
export RKS_LIB=$V
timeout 200 python -m pytest tests -m gpu -q > $O/${T}_suite_fixed.log 2>&1; echo "suite (fixed) rc=$?"; tail -2 $O/${T}_suite_fixed.log
