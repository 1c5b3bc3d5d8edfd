#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_forward.py tests/test_golden.py tests/test_gpu_configs.py -m gpu -q -x -p no:cacheprovider --timeout 600 2>&1 | tail -4
PILCO_B200_LIB=$PWD/pilco_b200/build_timing/libpilco_b200_timing.so python scripts/tile_phases.py 32 > gpurun_out/tile_phases.log 2>&1
python - <<'PY'
import json; r=json.load(open('gpurun_out/tile_phases.json')); print({k:(round(v) if isinstance(v,float) else v) for k,v in r['offdiag'].items()}); print({k:(round(v) if isinstance(v,float) else v) for k,v in r['diag'].items()})
PY
for cfg in "32 8" "64 8"; do
  set -- $cfg
  python bench.py --restarts $1 --nsplit $2 --no-backward --no-cpu-baseline --steps 10 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read()); r=l['roofline']; print('R=$1 nsplit=$2 value=%.0f e2e=%.0f tile_ms=%.4f setup_ms=%.4f frac=%.3f'%(l['value'],l['e2e']['value'],r['tile_kernel_ms'],r['setup_kernel_ms'],r['frac']))" | tee -a gpurun_out/exp2_bench.txt
done
