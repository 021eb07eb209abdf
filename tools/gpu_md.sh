#!/bin/bash
# row f2 on the GPU: parity tests of the bonded / integrator / thermostat kernels + their timing at C4 size
OUT=gpurun_out/r1md3; mkdir -p $OUT
timeout 50 python -m pytest tests/test_zgpu_md.py -q -m gpu --tb=short > $OUT/pytest.log 2>&1; tail -25 $OUT/pytest.log
timeout 40 python tools/bench_md.py --iters 10 --out $OUT/md_bench.json > $OUT/md_bench.log 2>&1; tail -3 $OUT/md_bench.log
