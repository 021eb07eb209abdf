#!/bin/bash
# plane-transform bring-up: parity tests, then bench with and without the plane kernels
OUT=gpurun_out/${1:-pl}; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > $OUT/pytest.log 2>&1; tail -15 $OUT/pytest.log
for np in 0 1; do
  HYMD_B200_NO_PLANE=$np timeout 600 python bench.py --steps 20 --no-cpu-baseline --no-e2e 2>$OUT/bench_$np.err | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print('NO_PLANE=$np', 'ms/step', round(d['ms_per_step'],3), {k:round(x['ms_per_step'],3) for k,x in d['phases'].items()})"
  tail -2 $OUT/bench_$np.err
done
