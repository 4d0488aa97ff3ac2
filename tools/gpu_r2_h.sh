#!/bin/bash
# round 2, call 8 (1 GPU): axis tiles with more resident CTAs for short axes; two-kernel long axis A/B again
mkdir -p gpurun_out; O=gpurun_out; T=r02h
echo "== parity subset"; timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_size_classes.py -x -q -k "axis or grid or cfg4 or cfg5 or separable or indexed or batched_2d" > $O/${T}_parity.log 2>&1; echo "rc=$?"; tail -3 $O/${T}_parity.log
for v in split nosplit; do
  if [ $v = nosplit ]; then export RKS_AXIS_SPLIT=0; else unset RKS_AXIS_SPLIT; fi
  timeout 200 python bench.py --workload cfg4 --no-cpu-baseline > $O/${T}_bench_cfg4_$v.json 2> $O/${T}_bench_cfg4_$v.err; echo "$v rc=$?"
  timeout 200 python bench.py --workload cfg4 --size 2048 --no-cpu-baseline > $O/${T}_bench_cfg4_2048_$v.json 2> $O/${T}_bench_cfg4_2048_$v.err; echo "$v 2048 rc=$?"
done
unset RKS_AXIS_SPLIT
timeout 200 python bench.py --workload cfg5 --size 256 --no-cpu-baseline > $O/${T}_bench_cfg5_256.json 2> $O/${T}_bench_cfg5_256.err; echo "cfg5 256 rc=$?"
timeout 200 python bench.py --workload cfg5 --no-cpu-baseline > $O/${T}_bench_cfg5_512.json 2> $O/${T}_bench_cfg5_512.err; echo "cfg5 512 rc=$?"
python - <<'PY'
import json, glob
for p in sorted(glob.glob("gpurun_out/r02h_bench_*.json")):
    try:
        d = json.load(open(p))
        print(p.split("r02h_bench_")[1], "ms/step %.3f e2e %.3e frac %.3f launches %s" % (d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["gpu_launches"]), d.get("clocks"))
    except Exception as e:
        print(p, "no line", e)
PY
echo "== launch list cfg4 (split)"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $O/${T}_launches_cfg4.csv python bench.py --workload cfg4 --no-cpu-baseline > $O/${T}_ncu_cfg4.log 2>&1; echo "rc=$?"
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/r02h_launches_cfg4.csv')) if len(r)>5]
hdr=next(r for r in rows if "Kernel Name" in r); i0=rows.index(hdr)
ki,vi,ui=hdr.index("Kernel Name"),hdr.index("Metric Value"),hdr.index("Metric Unit")
d=collections.defaultdict(list)
for r in rows[i0+1:]:
    v=float(r[vi].replace(",","")); v = v/1e3 if r[ui]=="ns" else v*1e3 if r[ui]=="ms" else v
    d[r[ki].split("(")[0][:64]].append(v)
for k,v in d.items():
    if 'rks' in k:
        big=[x for x in v if x>20]
        if big: print("%-66s n=%4d real=%4d avg_real=%7.1f max=%7.1f" % (k, len(v), len(big), sum(big)/len(big), max(big)))
PY
