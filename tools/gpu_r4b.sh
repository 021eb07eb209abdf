#!/bin/bash
# graph replay after the staging-copy fix: parity tests, then C1 / C2 / C3 bench lines with replay on
TAG=${1:-r4b}
OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 150 python -m pytest tests/test_gpu_graph.py -x -q --durations=5 > $OUT/pytest_graph.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_graph.log
tail -12 $OUT/pytest_graph.log
for w in C1 C2 C3; do
g=1
HYMD_B200_GRAPH=$g timeout 60 python bench.py --workload $w --steps 200 --warmup 10 --no-cpu-baseline > $OUT/bench_${w}_g$g.json 2> $OUT/bench_${w}_g$g.err; echo "bench $w graph=$g exit $?"
python - <<PY
import json
try:
    d = json.loads([l for l in open("$OUT/bench_${w}_g$g.json") if l.startswith("{")][-1])
    print("  $w g=$g ms/step", round(d["ms_per_step"], 4), "value %.4g" % d["value"], "e2e %.4g" % d["e2e"]["value"], d["parity"]["rel_err"], d["parity"]["ok"], d["gpu_launches"], d["graph"])
except Exception as e:
    print("  no line:", e)
PY
done
for f in $OUT/*.err; do tail -n 2 $f | cut -c1-300; done
