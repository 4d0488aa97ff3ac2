#!/bin/bash
# round 2, call 17 (1 GPU): inventory rows of cfg2 (plain evaluation kernel) and cfg2b (running rows) again; axis tile variants
mkdir -p gpurun_out; O=gpurun_out; T=r02q
for w in cfg2 cfg2b; do
  SECONDS=0; timeout 300 ncu --set full --clock-control none --profile-from-start off -k regex:kernel -f -o /tmp/${T}_inv_$w python tools/prof_all.py $w > $O/${T}_inv_$w.log 2>&1; echo "$w rc=$? ${SECONDS}s"
  ncu -i /tmp/${T}_inv_$w.ncu-rep --page raw --csv > $O/${T}_inv_${w}_raw.csv 2>/dev/null
done
python tools/ncu_summary.py cfg2=$O/${T}_inv_cfg2_raw.csv cfg2b=$O/${T}_inv_cfg2b_raw.csv > $O/${T}_ncu_cfg2_cfg2b.csv
python - <<'PY'
import csv
for r in csv.DictReader(open("gpurun_out/r02q_ncu_cfg2_cfg2b.csv")):
    try: print("%-6s %-62s %8.1f us dram %6.0f GB/s %5.1f%% lsu %5.1f%% fp64 %5.1f%% regs %s" % (r["workload"], r["kernel"][:62], float(r["us"]), float(r["dram_gbps"] or 0), float(r["dram_pct"] or 0), float(r["lsu_pct"] or 0), float(r["fp64_pct"] or 0), r["regs"]))
    except Exception as e: print(r.get("kernel"), e)
PY
gzip -f $O/${T}_inv_*_raw.csv
echo "== axis variants"
for v in main ax_u2b2 ax_t512b2 ax_c4t128b6 ax_c16; do
  if [ $v = main ]; then unset RKS_LIB; else export RKS_LIB=$PWD/rkstiff_b200/variants/$v.so; fi
  timeout 120 python tools/bench_axis.py > $O/${T}_axis_$v.txt 2>&1; echo "$v rc=$?"
done
unset RKS_LIB
paste -d'|' $O/${T}_axis_main.txt $O/${T}_axis_ax_u2b2.txt | cut -c1-160
echo "-- t512b2 | c4t128b6"; paste -d'|' $O/${T}_axis_ax_t512b2.txt $O/${T}_axis_ax_c4t128b6.txt | cut -c1-160
echo "-- c16"; cat $O/${T}_axis_ax_c16.txt
