#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_forward.py tests/test_golden.py tests/test_gpu_configs.py -m gpu -q -x -p no:cacheprovider --timeout 600 2>&1 | tail -3
PILCO_TILE_RPC=2 python -m pytest tests/test_gpu_forward.py tests/test_golden.py -m gpu -q -x -p no:cacheprovider --timeout 600 2>&1 | tail -3
for rpc in 1 2 5; do
  PILCO_TILE_RPC=$rpc python bench.py --restarts 32 --nsplit 8 --no-backward --no-cpu-baseline --steps 10 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read()); r=l['roofline']; print('rpc=$rpc R=32 value=%.0f e2e=%.0f tile_ms=%.4f setup_ms=%.4f frac=%.3f'%(l['value'],l['e2e']['value'],r['tile_kernel_ms'],r['setup_kernel_ms'],r['frac']))" | tee -a gpurun_out/exp4_bench.txt
done
python bench.py --restarts 64 --nsplit 8 --no-backward --no-cpu-baseline --steps 10 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read()); r=l['roofline']; print('auto R=64 value=%.0f e2e=%.0f tile_ms=%.4f frac=%.3f'%(l['value'],l['e2e']['value'],r['tile_kernel_ms'],r['frac']))" | tee -a gpurun_out/exp4_bench.txt
PILCO_B200_LIB=$PWD/pilco_b200/build_timing/libpilco_b200_timing.so python scripts/tile_phases.py 32 > gpurun_out/tile_phases.log 2>&1
python - <<'PY'
import json; r=json.load(open('gpurun_out/tile_phases.json')); print({k:(round(v) if isinstance(v,float) else v) for k,v in r['offdiag'].items()})
PY
