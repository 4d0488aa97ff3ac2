#!/bin/bash
# round 2, call 5 (1 GPU): parity of the paired-row cubic kernel and of the twiddle-table passes; K4 variants A/B
mkdir -p gpurun_out; O=gpurun_out; T=r02e
echo "== paired rows + full GPU suite"; timeout 1500 python -m pytest tests -m gpu -x -q > $O/${T}_gpu_suite.log 2>&1; echo "rc=$?"; tail -4 $O/${T}_gpu_suite.log
echo "== K4 alone, variants"
for v in main tw0 tw64 twsq0; do
  if [ $v = main ]; then unset RKS_LIB; else export RKS_LIB=$PWD/rkstiff_b200/variants/$v.so; fi
  timeout 300 python tools/bench_nl.py > $O/${T}_nl_$v.txt 2>&1; echo "$v rc=$?"
done
unset RKS_LIB
RKS_PAIR_ROWS=0 timeout 300 python tools/bench_nl.py > $O/${T}_nl_main_nopair.txt 2>&1
paste -d'|' $O/${T}_nl_main.txt $O/${T}_nl_tw0.txt | cut -c1-170
echo "-- tw64 | twsq0"; paste -d'|' $O/${T}_nl_tw64.txt $O/${T}_nl_twsq0.txt | cut -c1-170
echo "-- no pair"; grep cubic $O/${T}_nl_main_nopair.txt
echo "== bench lines main vs tw0"
for v in main tw0 tw64; do
  if [ $v = main ]; then unset RKS_LIB; else export RKS_LIB=$PWD/rkstiff_b200/variants/$v.so; fi
  for w in cfg2 cfg3 cfg4; do
    timeout 300 python bench.py --workload $w --no-cpu-baseline > $O/${T}_bench_${w}_$v.json 2> $O/${T}_bench_${w}_$v.err; echo "$v $w rc=$?"
  done
done
unset RKS_LIB
python - <<'PY'
import json
for v in ("main", "tw0", "tw64"):
    for w in ("cfg2", "cfg3", "cfg4"):
        try:
            d = json.load(open(f"gpurun_out/r02e_bench_{w}_{v}.json"))
        except Exception as e:
            print(v, w, "no line", e); continue
        r = d["roofline"]
        ks = {k.split(" ")[0]: round(x["us"], 1) for k, x in (r.get("kernels") or {}).items()}
        print(v, w, "ms/step %.3f e2e %.3e frac %.3f" % (d["ms_per_step"], d["e2e"]["value"], r["frac"]), "whole", (r.get("whole_step") or {}).get("frac"), ks, d.get("clocks"))
PY
