#!/bin/bash
# Round-end GPU visit: full GPU parity suite, the bench lines, the ncu launch list of the bench command and one
# `--set full` capture of the dominant kernel.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 600 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; echo "bench exit $?"; cat gpurun_out/bench_1gpu.json
python bench.py --restarts 64 --no-cpu-baseline --steps 10 > gpurun_out/bench_1gpu_R64.json 2>> gpurun_out/bench_1gpu.err; cat gpurun_out/bench_1gpu_R64.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 1 --no-backward --no-cpu-baseline --nsplit 1 > gpurun_out/ncu_bench.log 2>&1
echo "ncu launches exit $?"
ncu --set full --clock-control none --import-source on -k regex:mm_tile -s 1 -c 1 -o gpurun_out/mm_tile_full -f \
    python scripts/prof_mm.py 32 > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?"
ncu -i gpurun_out/mm_tile_full.ncu-rep --page raw --csv > gpurun_out/mm_tile_full_raw.csv 2>/dev/null
ls -la gpurun_out | head -20
