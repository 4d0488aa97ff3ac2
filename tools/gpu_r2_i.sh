#!/bin/bash
# round 2, call 9 (2 GPUs): NCCL parity tests after the teardown fix; slab exchange overlap: chunks x SM margin x NCCL stream priority
mkdir -p gpurun_out; O=gpurun_out; T=r02i
echo "== multi-GPU parity tests"; timeout 300 python -m pytest tests/test_gpu_multi.py -x -q > $O/${T}_multi.log 2>&1; echo "rc=$?"; tail -4 $O/${T}_multi.log
run() { timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 "${@:2}"; }
i=0
for cfg in "1 0 0" "4 0 1" "4 132 1" "4 116 1" "2 132 1" "2 132 0"; do
  set -- $cfg; i=$((i+1))
  export RKS_SLAB_CHUNKS=$1
  if [ $2 != 0 ]; then export RKS_SM_LIMIT=$2; else unset RKS_SM_LIMIT; fi
  if [ $3 = 1 ]; then export TORCH_NCCL_HIGH_PRIORITY=1; else unset TORCH_NCCL_HIGH_PRIORITY; fi
  run $((29530 + i)) --workload cfg5 --no-cpu-baseline > $O/${T}_cfg5_c$1_sm$2_hp$3.json 2> $O/${T}_cfg5_c$1_sm$2_hp$3.err; echo "chunks=$1 sm=$2 hp=$3 rc=$?"
done
unset RKS_SLAB_CHUNKS RKS_SM_LIMIT TORCH_NCCL_HIGH_PRIORITY
python - <<'PY'
import json, glob
for p in sorted(glob.glob("gpurun_out/r02i_cfg5_*.json")):
    txt = open(p).read()
    i = txt.find('{"metric"')
    try:
        x = json.loads(txt[i:txt.rfind('}') + 1])
        print(p.split("r02i_")[1], "ms/step %.3f value %.3e e2e %.3e" % (x["ms_per_step"], x["value"], x["e2e"]["value"]))
    except Exception as e:
        print(p, "no line", e)
PY
