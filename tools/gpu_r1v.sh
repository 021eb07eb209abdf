#!/bin/bash
OUT=gpurun_out/r1v; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "laplacian or energies or variants" > $OUT/pytest.log 2>&1; tail -15 $OUT/pytest.log
