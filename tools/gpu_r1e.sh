#!/bin/bash
# round-1 session e: parity of the new kernel variants + phase timings of each variant at C4
OUT=gpurun_out/r1e; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "variants or forces_match or pme_matches or consecutive" > $OUT/pytest.log 2>&1; tail -5 $OUT/pytest.log
timeout 900 python tools/variants.py --steps 10 --out $OUT/variants.jsonl \
   "-" "HYMD_B200_PLANE_NT=512" "HYMD_B200_GRAD2=0" "HYMD_B200_PLANE_NT=512,HYMD_B200_GRAD2=0" \
   "HYMD_B200_PLANE_GRID=444" 2> $OUT/variants.err | cut -c1-600
tail -3 $OUT/variants.err
