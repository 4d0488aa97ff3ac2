#!/bin/bash
# round 2, call 34 (1 GPU): pre-transformed rows: global (tail) loads of both first-pass butterflies issued first (main) vs in order
mkdir -p gpurun_out; O=gpurun_out; T=r02ae
for v in main notailfirst main2 notailfirst2; do
  case $v in main*) unset RKS_LIB;; *) export RKS_LIB=$PWD/rkstiff_b200/variants/notailfirst.so;; esac
  timeout 150 python bench.py --workload cfg2 --no-cpu-baseline > $O/${T}_cfg2_$v.json 2> $O/${T}_cfg2_$v.err; echo "$v cfg2 rc=$?"
done
unset RKS_LIB
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "pretransformed and 8192" > $O/${T}_pt_tests.log 2>&1; echo "pt tests rc=$?"; tail -1 $O/${T}_pt_tests.log
python - <<'PY'
import json, glob
for p in sorted(glob.glob("gpurun_out/r02ae_cfg*.json")):
    try:
        d = json.load(open(p)); print(p.split("r02ae_")[1], "ms/step %.3f value %.3e" % (d["ms_per_step"], d["value"]), {k[:8]: round(v["us"],1) for k, v in d["roofline"]["kernels"].items() if "nl" in k}, d["clocks"]["reasons"])
    except Exception as e: print(p, "no line", e)
PY
