#!/bin/bash
# ncu evidence for profiles/: launch list of one cycle + --set full capture of the hand-written kernels.
# bash tools/gpu_profile.sh <tag> [bench args]
TAG=${1:-prof}; shift
OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu.txt 2>&1
timeout 900 python bench.py "$@" > $OUT/bench.json 2> $OUT/bench.err; tail -c 600 $OUT/bench.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e "$@" > $OUT/ncu_launch.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on \
    -k regex:'paint_kernel|xline_kernel|readout_gather_kernel|readout_kernel|count_kernel|scatter_kernel|plane_r2c|plane_c2r' -s 21 -c 7 \
    -o $OUT/prof python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e "$@" > $OUT/ncu_full.log 2>&1
tail -2 $OUT/ncu_full.log | cut -c1-200
ls -la $OUT
