#!/bin/bash
# word-wise configuration digest of the graph replay: parity tests again, C1 bench line
TAG=${1:-r4f}
OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 25 python -m pytest tests/test_gpu_graph.py -x -q > $OUT/pytest_graph.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_graph.log
tail -3 $OUT/pytest_graph.log
timeout 20 python bench.py --workload C1 --steps 200 --warmup 10 --no-cpu-baseline > $OUT/bench_C1.json 2> $OUT/bench_C1.err; echo "bench exit $?"
python - <<PY
import json
d = json.loads([l for l in open("$OUT/bench_C1.json") if l.startswith("{")][-1])
print("C1 ms/step", round(d["ms_per_step"], 4), "value %.4g" % d["value"], "e2e %.4g" % d["e2e"]["value"], d["parity"]["rel_err"], d["parity"]["ok"], d["graph"])
PY
