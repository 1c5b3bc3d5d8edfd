#!/bin/bash
mkdir -p gpurun_out
PILCO_B200_LIB=$PWD/pilco_b200/build_timing/libpilco_b200_timing.so python scripts/tile_phases.py 32 > gpurun_out/tile_phases.log 2>&1
cat gpurun_out/tile_phases.log | tail -40
for cfg in "32 16" "64 16" "128 8" "128 16" "32 4"; do
  set -- $cfg
  python bench.py --restarts $1 --nsplit $2 --no-backward --no-cpu-baseline --steps 10 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read()); print('R=$1 nsplit=$2 value=%.0f e2e=%.0f tile_ms=%.4f'%(l['value'],l['e2e']['value'],l['roofline']['tile_kernel_ms']))" | tee -a gpurun_out/exp1_bench.txt
done
