#!/bin/bash
# round 2, call 31 (1 GPU): uneven slice staging of the pre-transformed rows (more staged points for the slots that start later / earlier)
mkdir -p gpurun_out; O=gpurun_out; T=r02ac
for v in main sl_256_512 sl_512_256 main2 sl_256_512b; do
  case $v in main*) unset RKS_LIB;; sl_256_512*) export RKS_LIB=$PWD/rkstiff_b200/variants/sl_256_512.so;; *) export RKS_LIB=$PWD/rkstiff_b200/variants/sl_512_256.so;; esac
  timeout 150 python bench.py --workload cfg2 --no-cpu-baseline > $O/${T}_cfg2_$v.json 2> $O/${T}_cfg2_$v.err; echo "$v cfg2 rc=$?"
done
export RKS_LIB=$PWD/rkstiff_b200/variants/sl_256_512.so
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "pretransformed and 8192" > $O/${T}_pt_tests.log 2>&1; echo "pt tests (sl_256_512) rc=$?"; tail -1 $O/${T}_pt_tests.log
unset RKS_LIB
python - <<'PY'
import json, glob
for p in sorted(glob.glob("gpurun_out/r02ac_cfg*.json")):
    try:
        d = json.load(open(p)); print(p.split("r02ac_")[1], "ms/step %.3f value %.3e" % (d["ms_per_step"], d["value"]), {k[:8]: round(v["us"],1) for k, v in d["roofline"]["kernels"].items() if "nl" in k}, d["clocks"]["reasons"])
    except Exception as e: print(p, "no line", e)
PY
