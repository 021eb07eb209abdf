#!/bin/bash
# virtual-slab tests first (new code), then the whole GPU suite
TAG=${1:-r2b}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export HYMD_B200_LOCAL_TIMEOUT_S=40
timeout 900 python -m pytest tests/test_gpu_virtual_slabs.py -x -q --durations=10 > $OUT/pytest_virtual.log 2>&1; echo "virtual exit $?" >> $OUT/pytest_virtual.log
tail -40 $OUT/pytest_virtual.log
timeout 1500 python -m pytest tests -m gpu -q --durations=10 --deselect tests/test_gpu_virtual_slabs.py > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log
tail -15 $OUT/pytest_gpu.log
