#!/bin/bash
# dih_type 1 on the GPU + regression of every bonded / MD test, memcheck of the new kernels
TAG=${1:-r3n}
OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 python -m pytest tests/test_zgpu_md.py tests/test_zzgpu_gpe.py -x -q --durations=5 > $OUT/pytest_md.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_md.log
tail -10 $OUT/pytest_md.log
timeout 420 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_zgpu_md.py -x -q -k "cbt" > $OUT/sanitizer_cbt.log 2>&1; echo "memcheck exit $?" >> $OUT/sanitizer_cbt.log
tail -4 $OUT/sanitizer_cbt.log
