#!/bin/bash
# bash tools/gpu_ncu.sh <tag> <kernel-regex> [skip] [count] [bench args...]
TAG=$1; RE=$2; SKIP=${3:-3}; CNT=${4:-1}; shift 4
OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv \
    --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e "$@" > $OUT/ncu_launch.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"$RE" -s $SKIP -c $CNT \
    -o $OUT/prof python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e "$@" > $OUT/ncu_full.log 2>&1
tail -3 $OUT/ncu_full.log | cut -c1-300
ls -la $OUT
