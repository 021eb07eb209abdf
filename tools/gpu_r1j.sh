#!/bin/bash
OUT=gpurun_out/r1j; mkdir -p $OUT
timeout 600 python tools/variants.py --steps 10 --out $OUT/variants.jsonl "HYMD_B200_ROW_TMA=0" "HYMD_B200_ROW_TMA=1" "HYMD_B200_ROW_TMA=2" "HYMD_B200_ROW_TMA=3" 2> $OUT/variants.err | cut -c1-300
tail -3 $OUT/variants.err
