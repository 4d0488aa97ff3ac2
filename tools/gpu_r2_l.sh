#!/bin/bash
# round 2, call 12 (2 GPUs): slab exchange through peer memory (scattering transforms) vs NCCL all-to-all
mkdir -p gpurun_out; O=gpurun_out; T=r02l
timeout 240 python -m pytest tests/test_gpu_multi.py -x -q > $O/${T}_multi.log 2>&1; echo "multi rc=$?"; tail -15 $O/${T}_multi.log | cut -c1-200
run() { timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 "${@:2}"; }
run 29561 --workload cfg5 --no-cpu-baseline > $O/${T}_cfg5_p2p.json 2> $O/${T}_cfg5_p2p.err; echo "cfg5 p2p rc=$?"; grep -i "warn\|error\|Traceback" $O/${T}_cfg5_p2p.err | head -5
RKS_SLAB_P2P=0 run 29562 --workload cfg5 --no-cpu-baseline > $O/${T}_cfg5_nccl.json 2> $O/${T}_cfg5_nccl.err; echo "cfg5 nccl rc=$?"
SECONDS=0; run 29563 > $O/${T}_bench_default_2gpu.json 2> $O/${T}_bench_default_2gpu.err; echo "default rc=$? wall=${SECONDS}s"
python - <<'PY'
import json
def load(p):
    txt = open(p).read(); i = txt.find('{"metric"')
    return json.loads(txt[i:txt.rfind('}') + 1])
for p in ("gpurun_out/r02l_cfg5_p2p.json", "gpurun_out/r02l_cfg5_nccl.json"):
    try:
        x = load(p); print(p.split("/")[-1], "ms/step %.3f value %.3e e2e %.3e trials %s" % (x["ms_per_step"], x["value"], x["e2e"]["value"], x["steps"]))
    except Exception as e: print(p, "no line", e)
try:
    d = load("gpurun_out/r02l_bench_default_2gpu.json")
    print("cfg2 x2: ms/step %.3f value %.3e e2e %.3e" % (d["ms_per_step"], d["value"], d["e2e"]["value"]))
    for k, v in d.get("secondary", {}).items():
        print(" ", k, v.get("error") or "ms/step %.3f value %.3e e2e %.3e frac %.3f" % (v["ms_per_step"], v["value"], v["e2e"]["value"], v["roofline"]["frac"]))
    print("  parity:", json.dumps(d.get("parity")))
except Exception as e: print("default: no line", e)
PY
