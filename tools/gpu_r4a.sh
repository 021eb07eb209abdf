#!/bin/bash
# graph replay of the per-step field update: parity tests, then C1 / C2 / C3 bench lines with and without it
TAG=${1:-r4a}
OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 150 python -m pytest tests/test_gpu_graph.py -x -q --durations=5 > $OUT/pytest_graph.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_graph.log
tail -15 $OUT/pytest_graph.log
for w in C1 C2 C3; do
for g in 0 1; do
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
done
tail -5 $OUT/*.err | cut -c1-300
