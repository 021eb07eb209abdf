#!/bin/bash
TAG=${1:-r2e}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export HYMD_B200_LOCAL_TIMEOUT_S=40
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "tensor_memory or field_forces_match" > $OUT/pytest_tmem.log 2>&1; echo "tmem exit $?" >> $OUT/pytest_tmem.log
tail -25 $OUT/pytest_tmem.log
timeout 900 python bench.py --no-e2e > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"; tail -3 $OUT/bench.err
HYMD_B200_PLANE_TMEM=0 timeout 900 python bench.py --no-e2e --no-cpu-baseline > $OUT/bench_notmem.json 2>> $OUT/bench.err; echo "bench exit $?"
python - <<PY
import json
for w in ("", "_notmem"):
    try:
        d = json.load(open("$OUT/bench%s.json" % w))
        print(w or "tmem", d["ms_per_step"], d["parity"]["rel_err"], d["parity"]["ok"])
        print("   ", {k: round(v["ms_per_step"], 4) for k, v in d["phases"].items()})
    except Exception as e:
        print(w, "ERR", e)
PY
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log
tail -8 $OUT/pytest_gpu.log
