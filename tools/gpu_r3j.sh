#!/bin/bash
# grouped pipeline: virtual-slab correctness on one GPU
TAG=${1:-r3j}
OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_virtual_slabs.py tests/test_gpu_nve.py -x -q > $OUT/pytest_slabs.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_slabs.log
tail -4 $OUT/pytest_slabs.log
