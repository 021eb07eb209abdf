#!/bin/bash
# row-walker paint: parity + A/B against the flattened kernel, ncu of the new kernel
TAG=${1:-r3b}
OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "paint or bitwise or field_forces_match or pme_matches" > $OUT/pytest_paint.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_paint.log
tail -8 $OUT/pytest_paint.log
timeout 600 python tools/variants.py --steps 20 "-" "HYMD_B200_PAINT=flat" "-" > $OUT/variants.log 2>&1; tail -8 $OUT/variants.log | cut -c1-400
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"paint_" -s 3 -c 1 \
    -o $OUT/prof python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/ncu_full.log 2>&1
tail -2 $OUT/ncu_full.log | cut -c1-200
