#!/bin/bash
OUT=gpurun_out/r1p; mkdir -p $OUT
timeout 600 python tools/variants.py --steps 10 --out $OUT/variants.jsonl "-" "HYMD_B200_SCR_ALT=1" 2> $OUT/variants.err | cut -c1-300
tail -3 $OUT/variants.err
