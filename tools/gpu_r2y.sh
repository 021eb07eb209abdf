#!/bin/bash
# single GPU: C5 (1e8 particles, 512^3) with every particle's force compared with the oracle
OUT=gpurun_out/${1:-r2y}; mkdir -p $OUT
free -g | head -2
timeout 1500 python bench.py --workload C5 --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --parity-max-n 200000000 > $OUT/bench_C5.json 2> $OUT/bench_C5.err; echo "C5 exit $?"; tail -3 $OUT/bench_C5.err
python - <<PY
import json
try:
    d = json.load(open("$OUT/bench_C5.json"))
    print("C5", d["ms_per_step"], "%.3e" % d["value"], d["parity"])
    print("   ", {k: (round(v["ms_per_step"], 4), round(v.get("frac_of_peak") or 0, 3)) for k, v in d["phases"].items()})
except Exception as e:
    print("ERR", e)
PY
