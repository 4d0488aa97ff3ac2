#!/bin/bash
# round 2, call 15 (2 GPUs): the engine's own peer barrier vs the symmetric-memory handle's
mkdir -p gpurun_out; O=gpurun_out; T=r02o
timeout 200 python -m pytest tests/test_gpu_multi.py -x -q > $O/${T}_multi.log 2>&1; echo "multi rc=$?"; tail -3 $O/${T}_multi.log | cut -c1-200
run() { timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 "${@:2}"; }
run 29581 --workload cfg5 --no-cpu-baseline > $O/${T}_cfg5_ownbar.json 2> $O/${T}_cfg5_ownbar.err; echo "own barrier rc=$?"
RKS_PEER_BARRIER=0 run 29582 --workload cfg5 --no-cpu-baseline > $O/${T}_cfg5_hdlbar.json 2> $O/${T}_cfg5_hdlbar.err; echo "handle barrier rc=$?"
run 29583 --workload cfg5 --size 128 --no-cpu-baseline > $O/${T}_cfg5_128_ownbar.json 2> $O/${T}_cfg5_128_ownbar.err; echo "128 own rc=$?"
RKS_PEER_BARRIER=0 run 29584 --workload cfg5 --size 128 --no-cpu-baseline > $O/${T}_cfg5_128_hdlbar.json 2> $O/${T}_cfg5_128_hdlbar.err; echo "128 hdl rc=$?"
python - <<'PY'
import json, glob
def load(p):
    txt = open(p).read(); i = txt.find('{"metric"')
    return json.loads(txt[i:txt.rfind('}') + 1])
for p in sorted(glob.glob("gpurun_out/r02o_cfg5_*.json")):
    try:
        x = load(p); print(p.split("r02o_")[1], "ms/step %.3f value %.3e trials %s" % (x["ms_per_step"], x["value"], x["steps"]))
    except Exception as e: print(p, "no line", e)
PY
