#!/bin/bash
# Round-2 first pass: GPU tests, smoke, bench (parity key) at C4 / C2 / C3.
TAG=${1:-r2a}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu.txt 2>&1
python -c "import os; print('cpus', os.cpu_count())" >> $OUT/gpu.txt
free -g >> $OUT/gpu.txt
timeout 1500 python -m pytest tests -m gpu -x -q --durations=15 > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log
tail -25 $OUT/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; tail -2 $OUT/smoke.log
timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"; tail -c 1500 $OUT/bench.json; tail -5 $OUT/bench.err
timeout 600 python bench.py --workload C2 --steps 50 > $OUT/bench_C2.json 2>> $OUT/bench.err; echo "C2 exit $?"
timeout 600 python bench.py --workload C3 --steps 20 > $OUT/bench_C3.json 2>> $OUT/bench.err; echo "C3 exit $?"
timeout 600 python bench.py --workload C1 --steps 100 > $OUT/bench_C1.json 2>> $OUT/bench.err; echo "C1 exit $?"
python - <<PY
import json
for w in ("", "_C1", "_C2", "_C3"):
    try:
        d = json.load(open("$OUT/bench%s.json" % w))
        print(w or "C4", d["ms_per_step"], d["value"], d["parity"], d["e2e"]["value"], d["e2e"].get("numpy_pageable", {}).get("ms_per_step"))
    except Exception as e:
        print(w, "ERR", e)
PY
