#!/bin/bash
# round 2, final 1-GPU call on the final tree: full suite, smoke, default bench line, reference arm,
# ncu --set full inventory of the workloads whose kernels changed (cfg 2, cfg 4), launch list of cfg 2
mkdir -p gpurun_out; O=gpurun_out; T=r02H
echo "== full GPU suite"; timeout 900 python -m pytest tests -m gpu -x -q > $O/${T}_gpu_suite.log 2>&1; echo "rc=$?"; tail -3 $O/${T}_gpu_suite.log
echo "== smoke"; timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/${T}_smoke.log 2>&1; echo "rc=$?"; tail -2 $O/${T}_smoke.log
echo "== default bench line"; SECONDS=0; timeout 600 python bench.py > $O/${T}_bench_default.json 2> $O/${T}_bench_default.err; echo "rc=$? wall=${SECONDS}s"; tail -3 $O/${T}_bench_default.err
echo "== reference arm"; SECONDS=0; timeout 300 python bench.py --impl reference > $O/${T}_bench_reference.json 2> $O/${T}_bench_reference.err; echo "rc=$? wall=${SECONDS}s"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02H_bench_default.json"))
def show(tag, x):
    if "error" in x: print(tag, "ERROR", x["error"]); return
    r = x.get("roofline", {})
    print(tag, "ms/step %.3f value %.3e e2e %s frac %.3f whole %s cpu %s" % (x["ms_per_step"], x["value"], ("%.3e" % x["e2e"]["value"]) if x.get("e2e") else None, r.get("frac"),
          (r.get("whole_step") or {}).get("frac"), (x.get("cpu_baseline") or {}).get("value")))
    print("   parity:", (x.get("cpu_baseline") or {}).get("parity"), "clocks:", x.get("clocks"), "launches:", x.get("gpu_launches"), "traffic:", r.get("traffic"))
    for k, v in (r.get("kernels") or {}).items(): print("      %-48s %7.1f us  frac %.3f x%d" % (k, v["us"], v["frac"], v["launches_per_step"]))
show("cfg2", d)
for k, v in d.get("secondary", {}).items(): show(k, v)
r = json.load(open("gpurun_out/r02H_bench_reference.json"))
print("reference arm: value %.3e cores %s kind %s" % (r["value"], r["cpu_baseline"]["cores"], r["cpu_baseline"]["kind"]))
PY
echo "== ncu inventory"
for w in cfg2; do
  SECONDS=0; timeout 420 ncu --set full --clock-control none --profile-from-start off -k regex:kernel -f -o /tmp/${T}_inv_$w python tools/prof_all.py $w > $O/${T}_inv_$w.log 2>&1; echo "$w rc=$? ${SECONDS}s"
  ncu -i /tmp/${T}_inv_$w.ncu-rep --page raw --csv > $O/${T}_inv_${w}_raw.csv 2>/dev/null
done
python tools/ncu_summary.py cfg2=$O/${T}_inv_cfg2_raw.csv > $O/${T}_ncu_cfg2.csv
python - <<'PY'
import csv
for r in csv.DictReader(open("gpurun_out/r02H_ncu_cfg2.csv")):
    try: print("%-6s %-62s %8.1f us dram %6.0f GB/s %5.1f%% lsu %5.1f%% fp64 %5.1f%% regs %s" % (r["workload"], r["kernel"][:62], float(r["us"]), float(r["dram_gbps"] or 0), float(r["dram_pct"] or 0), float(r["lsu_pct"] or 0), float(r["fp64_pct"] or 0), r["regs"]))
    except Exception as e: print(r.get("kernel"), e)
PY
echo "== launch list cfg2"; timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${T}_launches_cfg2.csv python bench.py --workload cfg2 --steps 2 --warmup 1 --no-cpu-baseline > $O/${T}_ncu_list.log 2>&1; echo "rc=$?"
python tools/summarize_launches.py $O/${T}_launches_cfg2.csv | head -16
gzip -f $O/${T}_inv_*_raw.csv; ls -la $O | grep ${T} | awk '{print $5, $9}'
