#!/bin/bash
TAG=${1:-r3r}
OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 python tools/variants.py --workload C5 --steps 6 "-" "HYMD_B200_PLANE_THREADS=256" > $OUT/variants_C5.log 2>&1; tail -3 $OUT/variants_C5.log | cut -c1-330
