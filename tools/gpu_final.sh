#!/bin/bash
# what the driver runs at round end, in one call: GPU tests, smoke, default bench line (+ C2 line)
OUT=gpurun_out/final; mkdir -p $OUT
timeout 900 python -m pytest tests -x -q -m gpu > $OUT/pytest.log 2>&1; tail -4 $OUT/pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py > $OUT/bench.json 2> $OUT/bench.err; tail -c 400 $OUT/bench.json; tail -2 $OUT/bench.err
timeout 300 python bench.py --workload C2 --no-cpu-baseline > $OUT/bench_C2.json 2> $OUT/bench_C2.err
python - <<'PY'
import json
for f in ['bench.json','bench_C2.json']:
    d=json.loads([l for l in open('gpurun_out/final/'+f) if l.startswith('{')][-1])
    print(f, 'ms/step %.4f value %.4g e2e %.4g (%.3f ms)'%(d['ms_per_step'], d['value'], d['e2e']['value'], d['e2e']['ms_per_step']), d['roofline']['kernel'], round(d['roofline']['frac'],3), d['roofline'].get('traffic'), d['clocks'])
PY
