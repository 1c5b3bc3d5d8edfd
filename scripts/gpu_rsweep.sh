#!/bin/bash
# restart-batch sweep of the bench workload (R in {1, 32, 256} on one GPU)
mkdir -p gpurun_out
for cfg in "1 1" "256 8" "32 16"; do
  set -- $cfg
  python bench.py --restarts $1 --nsplit $2 --no-backward --no-cpu-baseline --steps 5 2>/dev/null > gpurun_out/bench_R$1_ns$2.json
  python -c "
import json
l=json.load(open('gpurun_out/bench_R$1_ns$2.json')); r=l['roofline']; print('R=$1 nsplit=$2 value=%.0f e2e=%.0f ms_per_step=%.3f tile_ms=%.4f'%(l['value'],l['e2e']['value'],l['ms_per_step'],r['tile_kernel_ms']))"
done
