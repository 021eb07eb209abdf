#!/bin/bash
# single GPU: x-line in-place combine, persistent readout, vectorised c2r output: parity + A/B in one process
TAG=${1:-r3f}
OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_configs.py tests/test_zzgpu_gpe.py -x -q > $OUT/pytest_parity.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_parity.log
tail -4 $OUT/pytest_parity.log
HYMD_B200_C2R_VEC=1 timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "tensor_memory or field_forces_match" > $OUT/pytest_vec.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_vec.log
tail -2 $OUT/pytest_vec.log
timeout 600 python tools/variants.py --steps 20 "-" "HYMD_B200_XLINE_INPLACE=0" "HYMD_B200_READOUT_PERSIST=0" "HYMD_B200_C2R_VEC=1" "-" > $OUT/variants.log 2>&1; tail -8 $OUT/variants.log | cut -c1-330
timeout 600 python tools/variants.py --workload C3 --steps 20 "-" "HYMD_B200_XLINE_INPLACE=0" > $OUT/variants_C3.log 2>&1; tail -3 $OUT/variants_C3.log | cut -c1-330
