#!/bin/bash
# round 2, call 30 (1 GPU): pre-transformed rows: first 384 points of every slice staged (main) vs head of the row staged
mkdir -p gpurun_out; O=gpurun_out; T=r02ab
timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_size_classes.py -x -q -k "pretransformed or 8192 or cfg2" > $O/${T}_pt_tests.log 2>&1; echo "pt tests rc=$?"; tail -2 $O/${T}_pt_tests.log
for v in main headstage main2 headstage2; do
  case $v in main*) unset RKS_LIB;; *) export RKS_LIB=$PWD/rkstiff_b200/variants/headstage.so;; esac
  timeout 150 python bench.py --workload cfg2 --no-cpu-baseline > $O/${T}_cfg2_$v.json 2> $O/${T}_cfg2_$v.err; echo "$v cfg2 rc=$?"
done
unset RKS_LIB
RKS_ROW_STAGGER_CYC=800 timeout 150 python bench.py --workload cfg2 --no-cpu-baseline > $O/${T}_cfg2_main_d800.json 2> $O/${T}_cfg2_main_d800.err
RKS_ROW_STAGGER_CYC=1200 timeout 150 python bench.py --workload cfg2 --no-cpu-baseline > $O/${T}_cfg2_main_d1200.json 2> $O/${T}_cfg2_main_d1200.err
python - <<'PY'
import json, glob
for p in sorted(glob.glob("gpurun_out/r02ab_cfg*.json")):
    try:
        d = json.load(open(p)); print(p.split("r02ab_")[1], "ms/step %.3f value %.3e" % (d["ms_per_step"], d["value"]), {k[:8]: round(v["us"],1) for k, v in d["roofline"]["kernels"].items() if "nl" in k}, d["clocks"]["reasons"])
    except Exception as e: print(p, "no line", e)
PY
