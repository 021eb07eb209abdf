#!/bin/bash
OUT=gpurun_out/r1o; mkdir -p $OUT
timeout 1200 python -m pytest tests -x -q -m gpu > $OUT/pytest.log 2>&1; tail -6 $OUT/pytest.log
