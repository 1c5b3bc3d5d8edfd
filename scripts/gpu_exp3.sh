#!/bin/bash
mkdir -p gpurun_out
for v in 1 2 0; do
  PILCO_TILE_VARIANT=$v python bench.py --restarts 32 --nsplit 8 --no-backward --no-cpu-baseline --steps 10 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read()); r=l['roofline']; print('variant=$v R=32 value=%.0f e2e=%.0f tile_ms=%.4f setup_ms=%.4f frac=%.3f'%(l['value'],l['e2e']['value'],r['tile_kernel_ms'],r['setup_kernel_ms'],r['frac']))" | tee -a gpurun_out/exp3_bench.txt
done
PILCO_TILE_VARIANT=2 PILCO_B200_LIB=$PWD/pilco_b200/build_timing/libpilco_b200_timing.so python scripts/tile_phases.py 32 > gpurun_out/tile_phases.log 2>&1
python - <<'PY'
import json; r=json.load(open('gpurun_out/tile_phases.json')); print({k:(round(v) if isinstance(v,float) else v) for k,v in r['offdiag'].items()})
PY
