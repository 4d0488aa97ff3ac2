#!/bin/bash
# round 2, call 32 (1 GPU): pre-transformed rows: slice tails copied by TMA into the row slab (main) vs L2-prefetched global reads
mkdir -p gpurun_out; O=gpurun_out; T=r02ad
timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_size_classes.py -x -q -k "pretransformed or 8192 or cfg2" > $O/${T}_pt_tests.log 2>&1; echo "pt tests rc=$?"; tail -2 $O/${T}_pt_tests.log
for v in main notail main2 notail2 main3; do
  case $v in main*) unset RKS_LIB;; *) export RKS_LIB=$PWD/rkstiff_b200/variants/notail.so;; esac
  timeout 150 python bench.py --workload cfg2 --no-cpu-baseline > $O/${T}_cfg2_$v.json 2> $O/${T}_cfg2_$v.err; echo "$v cfg2 rc=$?"
done
unset RKS_LIB
python - <<'PY'
import json, glob
for p in sorted(glob.glob("gpurun_out/r02ad_cfg*.json")):
    try:
        d = json.load(open(p)); print(p.split("r02ad_")[1], "ms/step %.3f value %.3e" % (d["ms_per_step"], d["value"]), {k[:8]: round(v["us"],1) for k, v in d["roofline"]["kernels"].items() if "nl" in k}, d["clocks"]["reasons"])
    except Exception as e: print(p, "no line", e)
PY
