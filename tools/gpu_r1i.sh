#!/bin/bash
OUT=gpurun_out/r1i; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "variants or forces_match or pme_matches or consecutive or energies" > $OUT/pytest.log 2>&1; tail -5 $OUT/pytest.log
timeout 300 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "forces_match and 32-64-64" > $OUT/sanitizer.log 2>&1; tail -4 $OUT/sanitizer.log
timeout 600 python tools/variants.py --steps 10 --out $OUT/variants.jsonl "-" 2> $OUT/variants.err | cut -c1-420
tail -3 $OUT/variants.err
