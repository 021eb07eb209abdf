#!/bin/bash
OUT=gpurun_out/r1w; mkdir -p $OUT
timeout 900 python -m pytest tests/test_mgpu.py -x -q -m gpu -k "single_gpu" > $OUT/pytest.log 2>&1; tail -15 $OUT/pytest.log
