#!/bin/bash
# last checks of the round: second PME call, slab tests with the final planner, compute-sanitizer on the new kernels
TAG=${1:-r3l}
OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "second_pme or pme_matches or paint_kernels" > $OUT/pytest_pme2.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_pme2.log
tail -4 $OUT/pytest_pme2.log
timeout 600 python -m pytest tests/test_gpu_virtual_slabs.py tests/test_mgpu.py -x -q > $OUT/pytest_slabs.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_slabs.log
tail -3 $OUT/pytest_slabs.log
timeout 420 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -k "paint_kernels_are_bitwise_identical and mesh1 or test_field_forces_match_oracle and mesh8 or second_pme and mesh0 and float32" > $OUT/sanitizer_memcheck.log 2>&1; echo "memcheck exit $?" >> $OUT/sanitizer_memcheck.log
tail -6 $OUT/sanitizer_memcheck.log
HYMD_B200_XPIPE=2 timeout 420 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_virtual_slabs.py -x -q -k "test_exchange_modes and blocked and 4-mesh0" > $OUT/sanitizer_slabs.log 2>&1; echo "memcheck slabs exit $?" >> $OUT/sanitizer_slabs.log
tail -4 $OUT/sanitizer_slabs.log
