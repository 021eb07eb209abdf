#!/bin/bash
OUT=gpurun_out/${1:-r2x}; mkdir -p $OUT
export HYMD_B200_LOCAL_TIMEOUT_S=40
for i in 1 2 3 4 5 6; do
timeout 300 python -m pytest tests/test_gpu_virtual_slabs.py -q -k "one_empty" 2>&1 | tail -1
done
timeout 900 python -m pytest tests/test_gpu_virtual_slabs.py tests/test_gpu_nve.py -q --durations=5 > $OUT/pytest_virtual.log 2>&1; echo "virtual exit $?"; tail -12 $OUT/pytest_virtual.log
