#!/bin/bash
# multi-GPU correctness pass: bash tools/gpu_mgpu.sh <tag> <ngpus>
TAG=${1:-mg}; NG=${2:-2}
OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi -L > $OUT/gpus.txt
timeout 900 python -m pytest tests/test_mgpu.py -x -q -m gpu > $OUT/pytest_mgpu.log 2>&1; echo "exit $?" >> $OUT/pytest_mgpu.log
tail -30 $OUT/pytest_mgpu.log
