#!/bin/bash
# One consolidated GPU-box visit: GPU parity tests, then the bench line(s).  Everything lands in gpurun_out/.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 600 --durations=25 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err
echo "bench exit $?"; cat gpurun_out/bench_1gpu.json
for extra in "$@"; do
  python bench.py $extra --no-backward --steps 10 > "gpurun_out/bench_${extra// /_}.json" 2>> gpurun_out/bench_1gpu.err
  cat "gpurun_out/bench_${extra// /_}.json"
done
