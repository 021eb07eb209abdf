#!/bin/bash
# 64^2 / 128^2 planes: one 512-thread CTA per SM vs two of 256 (C2, C3), parity first
TAG=${1:-r3p}
OUT=gpurun_out/$TAG; mkdir -p $OUT
HYMD_B200_PLANE_THREADS=256 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_configs.py -x -q -k "field_forces_match or baseline_config or pme_matches" > $OUT/pytest.log 2>&1; echo "pytest exit $?" >> $OUT/pytest.log
tail -3 $OUT/pytest.log
for w in C3 C2; do
timeout 600 python tools/variants.py --workload $w --steps 40 "-" "HYMD_B200_PLANE_THREADS=256" "-" > $OUT/variants_$w.log 2>&1; tail -3 $OUT/variants_$w.log | cut -c1-420
done
