#!/bin/bash
OUT=gpurun_out/r1n; mkdir -p $OUT
timeout 600 python tools/variants.py --steps 10 --out $OUT/variants.jsonl "-" "HYMD_B200_XLINE_CARVEOUT=58" "HYMD_B200_XLINE_CH=4" "HYMD_B200_XLINE_CH=4,HYMD_B200_XLINE_CARVEOUT=72" "HYMD_B200_XLINE_CH=4,HYMD_B200_XLINE_CARVEOUT=58" "HYMD_B200_XLINE_CH=4,HYMD_B200_XLINE_CARVEOUT=86" 2> $OUT/variants.err | cut -c1-330
tail -3 $OUT/variants.err
