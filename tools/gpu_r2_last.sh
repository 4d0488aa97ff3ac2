#!/bin/bash
# round 2, last call (1 GPU): full suite, smoke, default bench line and reference arm on the last tree
mkdir -p gpurun_out; O=gpurun_out; T=r02M
echo "== full GPU suite"; timeout 900 python -m pytest tests -m gpu -x -q > $O/${T}_gpu_suite.log 2>&1; echo "rc=$?"; tail -3 $O/${T}_gpu_suite.log
echo "== smoke"; timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/${T}_smoke.log 2>&1; echo "rc=$?"; tail -2 $O/${T}_smoke.log
echo "== default bench line"; SECONDS=0; timeout 600 python bench.py > $O/${T}_bench_default.json 2> $O/${T}_bench_default.err; echo "rc=$? wall=${SECONDS}s"; tail -3 $O/${T}_bench_default.err
echo "== reference arm"; SECONDS=0; timeout 300 python bench.py --impl reference > $O/${T}_bench_reference.json 2> $O/${T}_bench_reference.err; echo "rc=$? wall=${SECONDS}s"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02M_bench_default.json"))
def show(tag, x):
    if "error" in x: print(tag, "ERROR", x["error"]); return
    r = x.get("roofline", {})
    print(tag, "ms/step %.3f value %.3e e2e %s frac %.3f whole %s" % (x["ms_per_step"], x["value"], ("%.3e" % x["e2e"]["value"]) if x.get("e2e") else None, r.get("frac"), (r.get("whole_step") or {}).get("frac")))
    print("   parity:", (x.get("cpu_baseline") or {}).get("parity"), "clocks:", x.get("clocks"), "launches:", x.get("gpu_launches"))
    for k, v in (r.get("kernels") or {}).items(): print("      %-48s %7.1f us  frac %.3f x%d" % (k, v["us"], v["frac"], v["launches_per_step"]))
show("cfg2", d)
for k, v in d.get("secondary", {}).items(): show(k, v)
PY
