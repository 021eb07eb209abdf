#!/bin/bash
OUT=gpurun_out/r1h; mkdir -p $OUT
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'plane_c2r_kernel|xline_kernel' -s 2 -c 2 \
    -o $OUT/prof python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --n 2000000 > $OUT/ncu_full.log 2>&1
tail -2 $OUT/ncu_full.log | cut -c1-200
