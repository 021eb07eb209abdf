#!/bin/bash
# quick iteration pass: parity tests + one bench line with phases.  bash tools/gpu_iter.sh <tag> [pytest args]
OUT=gpurun_out/${1:-it}; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > $OUT/pytest.log 2>&1; tail -8 $OUT/pytest.log
timeout 600 python bench.py --steps 20 --no-cpu-baseline 2>$OUT/bench.err > $OUT/bench.json; python - <<PY
import json
d=json.loads([l for l in open('$OUT/bench.json') if l.startswith('{')][-1])
print('ms/step', round(d['ms_per_step'],3), 'e2e', d['e2e'] and round(d['e2e']['ms_per_step'],3), {k:round(x['ms_per_step'],3) for k,x in d['phases'].items()})
PY
tail -2 $OUT/bench.err
