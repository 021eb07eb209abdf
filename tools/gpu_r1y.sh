#!/bin/bash
OUT=gpurun_out/r1y; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "variants" > $OUT/pytest.log 2>&1; tail -4 $OUT/pytest.log
timeout 600 python tools/variants.py --steps 10 --out $OUT/variants.jsonl "-" "HYMD_B200_PLANE_TILES=3" 2> $OUT/variants.err | cut -c1-300
tail -3 $OUT/variants.err
