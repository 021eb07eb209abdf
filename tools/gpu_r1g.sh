#!/bin/bash
OUT=gpurun_out/r1g; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "variants or forces_match or pme_matches or consecutive or energies" > $OUT/pytest.log 2>&1; tail -5 $OUT/pytest.log
timeout 900 python tools/variants.py --steps 10 --out $OUT/variants.jsonl "-" "HYMD_B200_GRAD2=0" 2> $OUT/variants.err | cut -c1-420
tail -3 $OUT/variants.err
