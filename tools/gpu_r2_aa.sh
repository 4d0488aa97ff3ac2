#!/bin/bash
# round 2, call 29 (1 GPU): reversed stagger order (warps that read the unstaged tail of the row first)
mkdir -p gpurun_out; O=gpurun_out; T=r02aa
run() { tag=$1; shift; env "$@" timeout 120 python bench.py --workload cfg2 --no-cpu-baseline > $O/${T}_cfg2_$tag.json 2> $O/${T}_cfg2_$tag.err; echo "$tag rc=$?"; }
run m1a X=1
run m3a RKS_ROW_STAGGER_MODE=3
run m1b X=1
run m3b RKS_ROW_STAGGER_MODE=3
run m3d800 RKS_ROW_STAGGER_MODE=3 RKS_ROW_STAGGER_CYC=800
run m3d1200 RKS_ROW_STAGGER_MODE=3 RKS_ROW_STAGGER_CYC=1200
python - <<'PY'
import json, glob
for p in sorted(glob.glob("gpurun_out/r02aa_cfg*.json")):
    try:
        d = json.load(open(p)); print(p.split("r02aa_")[1], "ms/step %.3f value %.3e" % (d["ms_per_step"], d["value"]), {k[:8]: round(v["us"],1) for k, v in d["roofline"]["kernels"].items() if "nl" in k}, d["clocks"]["reasons"])
    except Exception as e: print(p, "no line", e)
PY
