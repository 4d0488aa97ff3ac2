#!/bin/bash
# round 2, call 10 (2 GPUs): own block of the slab exchange as a device copy; default 2-GPU bench line
mkdir -p gpurun_out; O=gpurun_out; T=r02j
run() { timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 "${@:2}"; }
run 29541 --workload cfg5 --no-cpu-baseline > $O/${T}_cfg5.json 2> $O/${T}_cfg5.err; echo "cfg5 rc=$?"
RKS_SLAB_CHUNKS=4 run 29542 --workload cfg5 --no-cpu-baseline > $O/${T}_cfg5_c4.json 2> $O/${T}_cfg5_c4.err; echo "cfg5 c4 rc=$?"
SECONDS=0; run 29543 > $O/${T}_bench_default_2gpu.json 2> $O/${T}_bench_default_2gpu.err; echo "default rc=$? wall=${SECONDS}s"
timeout 200 python -m pytest tests/test_gpu_multi.py -x -q > $O/${T}_multi.log 2>&1; echo "multi rc=$?"; tail -2 $O/${T}_multi.log
python - <<'PY'
import json, glob
def load(p):
    txt = open(p).read(); i = txt.find('{"metric"')
    return json.loads(txt[i:txt.rfind('}') + 1])
for p in ("gpurun_out/r02j_cfg5.json", "gpurun_out/r02j_cfg5_c4.json"):
    try:
        x = load(p); print(p.split("/")[-1], "ms/step %.3f value %.3e e2e %.3e" % (x["ms_per_step"], x["value"], x["e2e"]["value"]))
    except Exception as e: print(p, "no line", e)
try:
    d = load("gpurun_out/r02j_bench_default_2gpu.json")
    print("cfg2 x2: ms/step %.3f value %.3e e2e %.3e launches %s" % (d["ms_per_step"], d["value"], d["e2e"]["value"], d["gpu_launches"]))
    for k, v in d.get("secondary", {}).items():
        print(" ", k, v.get("error") or "ms/step %.3f value %.3e e2e %.3e frac %.3f" % (v["ms_per_step"], v["value"], v["e2e"]["value"], v["roofline"]["frac"]))
    print("  parity ok:", {k: v.get("ok") for k, v in (d.get("parity") or {}).items()})
except Exception as e: print("default: no line", e)
PY
