#!/bin/bash
# every method on the two batched 1-D workloads (device-resident numbers + per-kernel roofline)
mkdir -p gpurun_out/matrix
for wl in cfg2 cfg3; do
  for m in IF4 ETD4 ETD5 IF34 ETD34 ETD35 IF45DP; do
    timeout 200 python bench.py --workload $wl --method $m --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/matrix/${wl}_${m}.json 2> gpurun_out/matrix/${wl}_${m}.err || tail -3 gpurun_out/matrix/${wl}_${m}.err
  done
done
