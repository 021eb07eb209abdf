#!/bin/bash
# parity tests + phase timings of the default build: bash tools/gpu_ab.sh  (A/B pass used while tuning)
OUT=gpurun_out/ab; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > $OUT/pytest.log 2>&1; tail -5 $OUT/pytest.log
timeout 600 python tools/variants.py --steps 10 --out $OUT/variants.jsonl "-" 2> $OUT/variants.err | cut -c1-300
tail -3 $OUT/variants.err
