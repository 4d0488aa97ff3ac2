#!/bin/bash
# round 2, call 14 (8 GPUs): the default bench line at N = 8 and N = 4 (cfg2 + cfg3 + cfg2b + slab cfg5 through peer memory + parity)
mkdir -p gpurun_out; O=gpurun_out; T=r02n
run() { timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $2 bench.py --gpus $1 "${@:3}"; }
SECONDS=0; run 8 29571 > $O/${T}_bench_default_8gpu.json 2> $O/${T}_bench_default_8gpu.err; echo "N=8 rc=$? wall=${SECONDS}s"
SECONDS=0; run 4 29572 > $O/${T}_bench_default_4gpu.json 2> $O/${T}_bench_default_4gpu.err; echo "N=4 rc=$? wall=${SECONDS}s"
SECONDS=0; RKS_SLAB_P2P=0 run 8 29573 --workload cfg5 --no-cpu-baseline > $O/${T}_cfg5_nccl_8gpu.json 2> $O/${T}_cfg5_nccl_8gpu.err; echo "N=8 cfg5 nccl rc=$? wall=${SECONDS}s"
grep -i "warn\|error\|Traceback" $O/${T}_bench_default_8gpu.err | head -5
python - <<'PY'
import json
def load(p):
    txt = open(p).read(); i = txt.find('{"metric"')
    return json.loads(txt[i:txt.rfind('}') + 1])
for n in (8, 4):
    try:
        d = load(f"gpurun_out/r02n_bench_default_{n}gpu.json")
        print(f"N={n} cfg2: ms/step %.3f value %.3e e2e %.3e" % (d["ms_per_step"], d["value"], d["e2e"]["value"]))
        for k, v in d.get("secondary", {}).items():
            print("   ", k, v.get("error") or "ms/step %.3f value %.3e frac %.3f" % (v["ms_per_step"], v["value"], v["roofline"]["frac"]), "e2e %.3e" % v["e2e"]["value"] if "e2e" in v else "")
        print("    parity ok:", {k: v.get("ok") for k, v in (d.get("parity") or {}).items()}, (d.get("parity") or {}).get("error"))
    except Exception as e: print(f"N={n}: no line", e)
try:
    x = load("gpurun_out/r02n_cfg5_nccl_8gpu.json"); print("N=8 cfg5 NCCL route: ms/step %.3f value %.3e" % (x["ms_per_step"], x["value"]))
except Exception as e: print("nccl: no line", e)
PY
