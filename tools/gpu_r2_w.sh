#!/bin/bash
# round 2, call 23 (1 GPU): derivative rows + rks_gemv (GPU tests), row stagger defaults on cfg 2, NP spin sweep, n = 8192 rows of every model
mkdir -p gpurun_out; O=gpurun_out; T=r02w
timeout 400 python -m pytest tests/test_gpu_reference_suite.py -x -q > $O/${T}_refsuite.log 2>&1; echo "reference suite rc=$?"; tail -3 $O/${T}_refsuite.log
timeout 400 python -m pytest tests/test_gpu_parity.py -x -q -k "diag or pretransformed or 8192" > $O/${T}_parity_sel.log 2>&1; echo "parity selection rc=$?"; tail -3 $O/${T}_parity_sel.log
run() {  # tag workload env...
  tag=$1; wl=$2; shift 2
  env "$@" timeout 120 python bench.py --workload $wl --no-cpu-baseline > $O/${T}_${wl}_$tag.json 2> $O/${T}_${wl}_$tag.err; echo "$wl $tag rc=$?"
}
run off cfg2 RKS_ROW_STAGGER_MODE=0 RKS_ROW_STAGGER_NP=0
run def cfg2 X=1
run np900 cfg2 RKS_ROW_STAGGER_CYC=900
run np1300 cfg2 RKS_ROW_STAGGER_CYC=1300
run off2 cfg2 RKS_ROW_STAGGER_MODE=0 RKS_ROW_STAGGER_NP=0
run def2 cfg2 X=1
BENCH_NL_N=8192 RKS_ROW_STAGGER_NP=0 timeout 120 python tools/bench_nl.py 25 > $O/${T}_nl8192_np0.txt 2>&1
BENCH_NL_N=8192 timeout 120 python tools/bench_nl.py 25 > $O/${T}_nl8192_np1.txt 2>&1
python - <<'PY'
import json, glob
for p in sorted(glob.glob("gpurun_out/r02w_cfg*.json")):
    try:
        d = json.load(open(p)); print(p.split("r02w_")[1], "ms/step %.3f value %.3e" % (d["ms_per_step"], d["value"]), {k[:8]: round(v["us"],1) for k, v in d["roofline"]["kernels"].items() if "nl" in k}, d["clocks"]["reasons"])
    except Exception as e: print(p, "no line", e)
PY
tail -n 4 $O/${T}_nl8192_np*.txt
