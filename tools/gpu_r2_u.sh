#!/bin/bash
# round 2, call 21 (1 GPU): per-row warp stagger (RKS_ROW_STAGGER_CYC) A/B, interleaved with the baseline
mkdir -p gpurun_out; O=gpurun_out; T=r02u
run() {  # tag workload env...
  tag=$1; wl=$2; shift 2
  env "$@" timeout 120 python bench.py --workload $wl --no-cpu-baseline > $O/${T}_${wl}_$tag.json 2> $O/${T}_${wl}_$tag.err; echo "$wl $tag rc=$?"
}
run d0a cfg2 RKS_ROW_STAGGER_CYC=0
run d600 cfg2 RKS_ROW_STAGGER_CYC=600
run d800 cfg2 RKS_ROW_STAGGER_CYC=800
run d1000 cfg2 RKS_ROW_STAGGER_CYC=1000
run d0b cfg2 RKS_ROW_STAGGER_CYC=0
run d1200 cfg2 RKS_ROW_STAGGER_CYC=1200
run g2d1200 cfg2 RKS_ROW_STAGGER_CYC=1200 RKS_ROW_STAGGER_GROUPS=2
run g2d2000 cfg2 RKS_ROW_STAGGER_CYC=2000 RKS_ROW_STAGGER_GROUPS=2
run d800b cfg2 RKS_ROW_STAGGER_CYC=800
run d0c cfg2 RKS_ROW_STAGGER_CYC=0
run d0 cfg4 RKS_ROW_STAGGER_CYC=0
run d500 cfg4 RKS_ROW_STAGGER_CYC=500
run d1000 cfg4 RKS_ROW_STAGGER_CYC=1000
for D in 0 500 1000 1500; do
  BENCH_NL_N=4096,8192 RKS_ROW_STAGGER_CYC=$D timeout 120 python tools/bench_nl.py 25 > $O/${T}_nl_$D.txt 2>&1; echo "nl $D rc=$?"
done
python - <<'PY'
import json, glob
for p in sorted(glob.glob("gpurun_out/r02u_cfg*.json")):
    try:
        d = json.load(open(p)); print(p.split("r02u_")[1], "ms/step %.3f value %.3e" % (d["ms_per_step"], d["value"]), {k[:8]: round(v["us"],1) for k, v in d["roofline"]["kernels"].items() if "nl" in k or "K4" in k})
    except Exception as e: print(p, "no line", e)
PY
tail -n 7 $O/${T}_nl_*.txt
