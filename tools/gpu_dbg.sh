#!/bin/bash
OUT=gpurun_out/r2dbg; mkdir -p $OUT
export HYMD_B200_LOCAL_TIMEOUT_S=40
for i in 1 2 3 4; do
timeout 300 python -m pytest tests/test_gpu_virtual_slabs.py -q -k "one_empty" 2>&1 | tail -3
done
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_virtual_slabs.py -q -k "one_empty and float32" > $OUT/memcheck.log 2>&1
grep -E "Invalid|ERROR SUMMARY|at .*\+0x|by hymd|in hymd" $OUT/memcheck.log | head -40
tail -5 $OUT/memcheck.log
