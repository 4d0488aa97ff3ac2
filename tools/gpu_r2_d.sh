#!/bin/bash
# round 2, call 4 (1 GPU): quick parity after the removal of the fused K1+K4 path, the new default bench line
# (cfg2 + secondary cfg3/4/5), ncu --set full inventory of every kernel family at the bench geometry
mkdir -p gpurun_out; O=gpurun_out; T=r02d
echo "== parity (subset)"; timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_size_classes.py -x -q -k "fused or pretransformed or size or cfg or adaptive_dt" > $O/${T}_parity.log 2>&1; echo "rc=$?"; tail -3 $O/${T}_parity.log
echo "== default bench line"; SECONDS=0; timeout 900 python bench.py > $O/${T}_bench_default.json 2> $O/${T}_bench_default.err; echo "rc=$? wall=${SECONDS}s"; tail -5 $O/${T}_bench_default.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r02d_bench_default.json"))
except Exception as e:
    print("no bench line:", e); raise SystemExit
def show(tag, x):
    if "error" in x:
        print(tag, "ERROR", x["error"]); return
    r = x.get("roofline", {})
    print(tag, "ms/step %.3f value %.3e e2e %.3e frac %.3f whole %s cpu %s" % (x["ms_per_step"], x["value"], x["e2e"]["value"], r.get("frac"),
          (r.get("whole_step") or {}).get("frac"), (x.get("cpu_baseline") or {}).get("value")))
    print("   parity:", (x.get("cpu_baseline") or {}).get("parity"), "clocks:", x.get("clocks"), "launches:", x.get("gpu_launches"))
show("cfg2", d)
for k, v in d.get("secondary", {}).items():
    show(k, v)
PY
echo "== ncu inventory"
for w in cfg2 cfg3 cfg4 cfg5 cfg2b; do
  timeout 900 ncu --set full --clock-control none --profile-from-start off -k regex:kernel -f -o /tmp/${T}_inv_$w python tools/prof_all.py $w > $O/${T}_inv_$w.log 2>&1; echo "$w rc=$?"
  ncu -i /tmp/${T}_inv_$w.ncu-rep --page raw --csv > $O/${T}_inv_${w}_raw.csv 2>/dev/null
done
python tools/ncu_summary.py cfg2=$O/${T}_inv_cfg2_raw.csv cfg3=$O/${T}_inv_cfg3_raw.csv cfg4=$O/${T}_inv_cfg4_raw.csv cfg5=$O/${T}_inv_cfg5_raw.csv cfg2b=$O/${T}_inv_cfg2b_raw.csv > $O/${T}_ncu_all_kernels.csv
wc -l $O/${T}_ncu_all_kernels.csv; cut -d, -f1,2,7,8,9,10,11,12,13 $O/${T}_ncu_all_kernels.csv | cut -c1-200
cp /tmp/${T}_inv_cfg2.ncu-rep /tmp/${T}_inv_cfg3.ncu-rep $O/ 2>/dev/null; ls -la $O/*.ncu-rep
