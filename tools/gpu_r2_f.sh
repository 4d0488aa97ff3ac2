#!/bin/bash
# round 2, call 6 (2 GPUs): NCCL parity tests, the default 2-GPU bench line (cfg2 + cfg3 + slab cfg5 + parity block),
# A/B of the graph-captured sharded trial and of the pipelined slab exchange
mkdir -p gpurun_out; O=gpurun_out; T=r02f
echo "== multi-GPU parity tests"; timeout 900 python -m pytest tests/test_gpu_multi.py -x -q > $O/${T}_multi.log 2>&1; echo "rc=$?"; tail -4 $O/${T}_multi.log
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 "${@:2}"; }
echo "== default 2-GPU line"; SECONDS=0; timeout 900 bash -c "$(declare -f run); run 29511" > $O/${T}_bench_default_2gpu.json 2> $O/${T}_bench_default_2gpu.err; echo "rc=$? wall=${SECONDS}s"; tail -3 $O/${T}_bench_default_2gpu.err
echo "== sharded trial: graph vs eager loop"
RKS_GROUP_GRAPH=1 timeout 300 bash -c "$(declare -f run); run 29512 --workload cfg2 --no-cpu-baseline" > $O/${T}_cfg2_graph.json 2> $O/${T}_cfg2_graph.err; echo "rc=$?"
RKS_GROUP_GRAPH=0 timeout 300 bash -c "$(declare -f run); run 29513 --workload cfg2 --no-cpu-baseline" > $O/${T}_cfg2_eager.json 2> $O/${T}_cfg2_eager.err; echo "rc=$?"
echo "== slab exchange chunks"
for c in 1 2 4 8; do
  RKS_SLAB_CHUNKS=$c timeout 300 bash -c "$(declare -f run); run $((29520 + c)) --workload cfg5 --no-cpu-baseline" > $O/${T}_cfg5_chunks$c.json 2> $O/${T}_cfg5_chunks$c.err; echo "chunks=$c rc=$?"
done
python - <<'PY'
import json, glob
def load(p):
    try: return json.load(open(p))
    except Exception as e: return {"error": str(e)}
d = load("gpurun_out/r02f_bench_default_2gpu.json")
if "error" in d: print("default:", d)
else:
    print("cfg2 x2: ms/step %.3f value %.3e e2e %.3e launches %s" % (d["ms_per_step"], d["value"], d["e2e"]["value"], d["gpu_launches"]))
    for k, v in d.get("secondary", {}).items():
        print(" ", k, v.get("error") or "ms/step %.3f value %.3e e2e %.3e frac %.3f" % (v["ms_per_step"], v["value"], v["e2e"]["value"], v["roofline"]["frac"]))
    print("  parity:", json.dumps(d.get("parity"), indent=1))
for p in sorted(glob.glob("gpurun_out/r02f_cfg*.json")):
    x = load(p)
    print(p.split("/")[-1], x.get("error") or "ms/step %.3f value %.3e" % (x["ms_per_step"], x["value"]))
PY
