#!/bin/bash
OUT=gpurun_out/r1k; mkdir -p $OUT
timeout 600 python tools/variants.py --steps 10 --out $OUT/variants.jsonl "-" "HYMD_B200_ROW_TMA=0" "HYMD_B200_ROW_TMA=0,HYMD_B200_PLANE_SMEM_PAD=32" "HYMD_B200_ROW_TMA=0,HYMD_B200_PLANE_SMEM_PAD=64" "HYMD_B200_PLANE_SMEM_PAD=64" 2> $OUT/variants.err | cut -c1-300
tail -3 $OUT/variants.err
