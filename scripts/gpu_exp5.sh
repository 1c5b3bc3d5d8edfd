#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x -p no:cacheprovider --timeout 600 2>&1 | tail -3
python bench.py --restarts 32 --no-cpu-baseline --steps 10 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read()); r=l['roofline']; print('R=32 value=%.0f e2e=%.0f fwdbwd=%.0f tile_ms=%.4f setup_ms=%.4f frac=%.3f'%(l['value'],l['e2e']['value'],l['fwd_bwd']['value'],r['tile_kernel_ms'],r['setup_kernel_ms'],r['frac']))" | tee -a gpurun_out/exp5_bench.txt
python bench.py --restarts 64 --no-backward --no-cpu-baseline --steps 10 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read()); r=l['roofline']; print('R=64 value=%.0f e2e=%.0f tile_ms=%.4f setup_ms=%.4f frac=%.3f'%(l['value'],l['e2e']['value'],r['tile_kernel_ms'],r['setup_kernel_ms'],r['frac']))" | tee -a gpurun_out/exp5_bench.txt
