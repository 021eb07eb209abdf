#!/bin/bash
OUT=gpurun_out/r1z; mkdir -p $OUT
timeout 900 python -m pytest tests/test_mgpu.py -x -q -m gpu > $OUT/pytest_mgpu.log 2>&1; tail -5 $OUT/pytest_mgpu.log
