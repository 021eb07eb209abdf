#!/bin/bash
# pipelined exchange on virtual slabs (correctness), sort two-per-thread A/B, paint A/B
TAG=${1:-r3c}
OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_virtual_slabs.py tests/test_mgpu.py -x -q --durations=5 > $OUT/pytest_slabs.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_slabs.log
tail -12 $OUT/pytest_slabs.log
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_configs.py -x -q > $OUT/pytest_parity.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_parity.log
tail -4 $OUT/pytest_parity.log
timeout 600 python tools/variants.py --steps 20 "HYMD_B200_PAINT=flat" "HYMD_B200_PAINT=flat,HYMD_B200_SCATTER=1" "HYMD_B200_PAINT=flat" > $OUT/variants.log 2>&1; tail -8 $OUT/variants.log | cut -c1-330
