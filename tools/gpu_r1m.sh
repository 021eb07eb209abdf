#!/bin/bash
OUT=gpurun_out/r1m; mkdir -p $OUT
timeout 600 python tools/variants.py --steps 10 --out $OUT/variants.jsonl "HYMD_B200_PLANE_TILES=1,HYMD_B200_ROW_TMA=0" "HYMD_B200_PLANE_TILES=1,HYMD_B200_ROW_TMA=2" "HYMD_B200_PLANE_TILES=2,HYMD_B200_ROW_TMA=1" 2> $OUT/variants.err | cut -c1-300
tail -3 $OUT/variants.err
