#!/bin/bash
# round 2, call 22 (1 GPU): derivative rows on the engine kernels (GPU test); row-stagger modes on cfg 2; n = 4096 PT rows
mkdir -p gpurun_out; O=gpurun_out; T=r02v
timeout 300 python -m pytest tests/test_gpu_reference_suite.py -x -q -k derivatives > $O/${T}_deriv_test.log 2>&1; echo "deriv test rc=$?"; tail -5 $O/${T}_deriv_test.log
run() {  # tag workload env...
  tag=$1; wl=$2; shift 2
  env "$@" timeout 120 python bench.py --workload $wl --no-cpu-baseline > $O/${T}_${wl}_$tag.json 2> $O/${T}_${wl}_$tag.err; echo "$wl $tag rc=$?"
}
run m0a cfg2 RKS_ROW_STAGGER_MODE=0
run m1d1100 cfg2 RKS_ROW_STAGGER_MODE=1 RKS_ROW_STAGGER_CYC=1100
run m1d1300 cfg2 RKS_ROW_STAGGER_MODE=1 RKS_ROW_STAGGER_CYC=1300
run m1d1500 cfg2 RKS_ROW_STAGGER_MODE=1 RKS_ROW_STAGGER_CYC=1500
run m2 cfg2 RKS_ROW_STAGGER_MODE=2
run m0b cfg2 RKS_ROW_STAGGER_MODE=0
run m2np cfg2 RKS_ROW_STAGGER_MODE=2 RKS_ROW_STAGGER_NP=1 RKS_ROW_STAGGER_CYC=1100
run m2b cfg2 RKS_ROW_STAGGER_MODE=2
RKS_ROW_STAGGER_MODE=2 timeout 120 python bench.py --workload cfg2 --method IF45DP --no-cpu-baseline > $O/${T}_cfg2_m2if45.json 2> $O/${T}_cfg2_m2if45.err; echo "if45 m2 rc=$?"
RKS_ROW_STAGGER_MODE=0 timeout 120 python bench.py --workload cfg2 --method IF45DP --no-cpu-baseline > $O/${T}_cfg2_m0if45.json 2> $O/${T}_cfg2_m0if45.err; echo "if45 m0 rc=$?"
for M in 0 1 2; do
  PT_SWEEP_N=2048,4096,8192 RKS_ROW_STAGGER_MODE=$M timeout 200 python tools/pt_sweep.py ETD35 12 > $O/${T}_ptsweep_m$M.txt 2>&1; echo "ptsweep mode $M rc=$?"
done
python - <<'PY'
import json, glob
for p in sorted(glob.glob("gpurun_out/r02v_cfg*.json")):
    try:
        d = json.load(open(p)); print(p.split("r02v_")[1], "ms/step %.3f value %.3e" % (d["ms_per_step"], d["value"]), {k[:8]: round(v["us"],1) for k, v in d["roofline"]["kernels"].items() if "nl" in k}, d["clocks"]["reasons"])
    except Exception as e: print(p, "no line", e)
PY
tail -n 5 $O/${T}_ptsweep_m*.txt
