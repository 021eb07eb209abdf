#!/bin/bash
# C5 on one GPU (512^2 planes, L2-scratch plane kernels): how many resident CTAs (scratch planes) fit the L2?
TAG=${1:-r3g}
OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 1500 python tools/variants.py --workload C5 --steps 6 "-" "HYMD_B200_PLANE_GRID=112" "HYMD_B200_PLANE_GRID=96" "HYMD_B200_PLANE_GRID=80" "HYMD_B200_PLANE_GRID=64" > $OUT/variants_C5.log 2>&1; tail -6 $OUT/variants_C5.log | cut -c1-300
