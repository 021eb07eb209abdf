#!/bin/bash
TAG=${1:-r2v}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export HYMD_B200_LOCAL_TIMEOUT_S=40
timeout 900 python -m pytest tests/test_gpu_virtual_slabs.py -q --durations=5 > $OUT/pytest_virtual.log 2>&1; echo "virtual exit $?" >> $OUT/pytest_virtual.log
tail -30 $OUT/pytest_virtual.log
HYMD_B200_EXCHANGE=kernels timeout 900 python -m pytest tests/test_gpu_virtual_slabs.py -q -k "match_oracle" > $OUT/pytest_virtual_nofuse.log 2>&1; echo "nofuse exit $?" >> $OUT/pytest_virtual_nofuse.log
tail -5 $OUT/pytest_virtual_nofuse.log
