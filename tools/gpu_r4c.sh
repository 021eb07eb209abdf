#!/bin/bash
# closing pass of the third session: every GPU test with the final tree (graph replay automatic), smoke, default bench
TAG=${1:-r4c}
OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 170 python -m pytest tests -x -q -m gpu --durations=5 > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log
tail -12 $OUT/pytest_gpu.log
timeout 40 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke exit $?"; tail -1 $OUT/smoke.log
timeout 100 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"; tail -3 $OUT/bench.err
python - <<PY
import json
d = json.loads([l for l in open("$OUT/bench.json") if l.startswith("{")][-1])
print("C4 ms/step", round(d["ms_per_step"], 4), "value %.4g" % d["value"], "e2e %.4g" % d["e2e"]["value"], d["parity"]["rel_err"], d["parity"]["ok"], d["roofline"]["kernel"], round(d["roofline"]["frac"], 3), d["graph"])
PY
