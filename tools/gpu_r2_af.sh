#!/bin/bash
# round 2, call 35 (1 GPU): every method on the two batched 1-D workloads with the final kernels
mkdir -p gpurun_out/r02_matrix
for wl in cfg2 cfg3; do
  for m in IF4 ETD4 ETD5 IF34 ETD34 ETD35 IF45DP; do
    timeout 120 python bench.py --workload $wl --method $m --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r02_matrix/${wl}_${m}.json 2> gpurun_out/r02_matrix/${wl}_${m}.err || tail -3 gpurun_out/r02_matrix/${wl}_${m}.err
  done
done
python - <<'PY'
import json, glob
for p in sorted(glob.glob("gpurun_out/r02_matrix/*.json")):
    try:
        d = json.load(open(p)); r = d["roofline"]
        print(p.split("/")[-1], "ms/step %.3f value %.3e whole %.3f" % (d["ms_per_step"], d["value"], (r.get("whole_step") or {}).get("frac", float("nan"))), d["clocks"]["reasons"])
    except Exception as e: print(p, "no line", e)
PY
