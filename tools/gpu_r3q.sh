#!/bin/bash
TAG=${1:-r3q}
OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_virtual_slabs.py -x -q -k "half_size or pipelined" > $OUT/pytest.log 2>&1; echo "pytest exit $?" >> $OUT/pytest.log
tail -3 $OUT/pytest.log
timeout 600 python bench.py --workload C3 --steps 50 --warmup 5 --no-cpu-baseline > $OUT/bench_C3.json 2> $OUT/bench_C3.err; echo "bench C3 exit $?"
python - <<PY
import json
d = json.loads([l for l in open("$OUT/bench_C3.json") if l.startswith("{")][-1])
print("C3 ms/step", round(d["ms_per_step"], 4), "value %.4g" % d["value"], d["parity"]["rel_err"], d["parity"]["ok"])
print("   ", {k: round(v["ms_per_step"], 4) for k, v in d["phases"].items()})
PY
