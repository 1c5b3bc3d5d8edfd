#!/bin/bash
# tuning sweep (not a bench line): forward and forward+backward throughput vs kernel variants / sub-batch count
mkdir -p gpurun_out
for cfg in "0 0 32 8" "1 0 32 8" "0 1 32 8" "0 3 32 8" "0 1 32 16" "0 0 32 16" "0 1 64 8" "0 0 64 16"; do
  set -- $cfg
  PILCO_NO_PRIORITY=$1 PILCO_TAPE_VARIANT=$2 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-api-line --restarts $3 --nsplit $4 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read())
print('noprio=$1 tape_variant=$2 R=$3 nsplit=$4 fwd %.0f e2e %.0f fwd_bwd %.0f fb_e2e %.0f tile %.3f ttile %.3f'%(l['value'], l['e2e']['value'], l['fwd_bwd']['value'], l['fwd_bwd']['e2e']['value'], l['roofline']['kernel_ms'], l['fwd_bwd']['roofline']['kernel_ms']))
"
done
