#!/bin/bash
# One GPU call that produces everything under profiles/ for a round: bench line (metric config), launch list of one
# forward+backward step, ncu --set full of the three dominant kernels.  Usage: bash scripts/profile_round.sh r02
TAG=${1:-r02}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > $OUT/clocks.csv &
SMI=$!
python bench.py > $OUT/bench_metric.json 2> $OUT/bench_metric.err
python bench.py --impl reference --steps 20 --warmup 3 > $OUT/bench_metric_reference.json 2>> $OUT/bench_metric.err
kill $SMI
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_fwd_bwd.csv python scripts/profile_step.py --H 3 > $OUT/p_launch.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_fwd_r1.csv python scripts/profile_step.py --R 1 --H 3 --no-backward > $OUT/p_launch_r1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:mm_tile_kernel -s 5 -c 1 -o $OUT/mm_tile python scripts/profile_step.py --H 2 --no-backward > $OUT/p_tile.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:mm_tape_tile -s 3 -c 1 -o $OUT/mm_tape_tile python scripts/profile_step.py --H 2 > $OUT/p_tape.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rb_dyn_finish -s 2 -c 1 -o $OUT/rb_dyn_finish python scripts/profile_step.py --H 2 > $OUT/p_fin.log 2>&1
ls -la $OUT
