#!/bin/bash
OUT=gpurun_out/r1t; mkdir -p $OUT
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 60 --csv \
    --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/ncu_launch.log 2>&1
tail -2 $OUT/ncu_launch.log | cut -c1-200
