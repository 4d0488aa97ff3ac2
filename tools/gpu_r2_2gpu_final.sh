#!/bin/bash
# round 2, final 2-GPU sanity on the final tree: multi-GPU parity tests, default bench line at N = 2 (parity block)
mkdir -p gpurun_out; O=gpurun_out; T=r02G
timeout 240 python -m pytest tests/test_gpu_multi.py -x -q > $O/${T}_multi.log 2>&1; echo "multi rc=$?"; tail -3 $O/${T}_multi.log | cut -c1-200
timeout 120 python -m pytest tests/test_gpu_reference_suite.py -x -q -k "gemv or derivatives" > $O/${T}_new_tests.log 2>&1; echo "new tests rc=$?"; tail -2 $O/${T}_new_tests.log
SECONDS=0; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 2 > $O/${T}_bench_default_2gpu.json 2> $O/${T}_bench_default_2gpu.err; echo "default rc=$? wall=${SECONDS}s"
python - <<'PY'
import json
def load(p):
    txt = open(p).read(); i = txt.find('{"metric"')
    return json.loads(txt[i:txt.rfind('}') + 1])
try:
    d = load("gpurun_out/r02G_bench_default_2gpu.json")
    print("cfg2 x2: ms/step %.3f value %.3e e2e %.3e" % (d["ms_per_step"], d["value"], d["e2e"]["value"]))
    for k, v in d.get("secondary", {}).items():
        print(" ", k, v.get("error") or "ms/step %.3f value %.3e frac %.3f" % (v["ms_per_step"], v["value"], v["roofline"]["frac"]))
    print("  parity:", json.dumps(d.get("parity")))
except Exception as e: print("default: no line", e)
PY
