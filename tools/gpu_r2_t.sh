#!/bin/bash
# round 2, call 20 (1 GPU): start-delay (stagger) experiment on the K4 kernels
mkdir -p gpurun_out; O=gpurun_out; T=r02t
for D in 0 300 700 1400 2800; do
  RKS_STAGGER_CYC=$D timeout 120 python tools/bench_nl.py 25 > $O/${T}_nl_stagger_$D.txt 2>&1; echo "stagger $D rc=$?"
done
for D in 0 400 800 1600; do
  RKS_STAGGER8_CYC=$D timeout 120 python bench.py --workload cfg2 --no-cpu-baseline > $O/${T}_cfg2_stagger8_$D.json 2> $O/${T}_cfg2_stagger8_$D.err; echo "stagger8 $D rc=$?"
done
for D in 0 700 1400; do
  RKS_STAGGER_CYC=$D timeout 120 python bench.py --workload cfg3 --no-cpu-baseline > $O/${T}_cfg3_stagger_$D.json 2> $O/${T}_cfg3_stagger_$D.err; echo "cfg3 stagger $D rc=$?"
done
python - <<'PY'
import json, glob
for p in sorted(glob.glob("gpurun_out/r02t_cfg*.json")):
    try:
        d = json.load(open(p)); print(p.split("r02t_")[1], "ms/step %.3f value %.3e" % (d["ms_per_step"], d["value"]), {k: round(v["us"],1) for k, v in d["roofline"]["kernels"].items() if "nl" in k})
    except Exception as e: print(p, "no line", e)
PY
