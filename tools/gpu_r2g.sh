#!/bin/bash
TAG=${1:-r2g}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "tensor_memory or field_forces_match" > $OUT/pytest_tmem.log 2>&1; echo "tmem exit $?" >> $OUT/pytest_tmem.log
tail -5 $OUT/pytest_tmem.log
HYMD_B200_TMEM_DBUF=0 timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "tensor_memory" > $OUT/pytest_tmem0.log 2>&1; echo "tmem0 exit $?" >> $OUT/pytest_tmem0.log
tail -3 $OUT/pytest_tmem0.log
for v in 1 0; do
HYMD_B200_TMEM_DBUF=$v timeout 900 python bench.py --no-e2e --no-cpu-baseline > $OUT/bench_dbuf$v.json 2> $OUT/bench.err; echo "bench exit $?"; tail -3 $OUT/bench.err
done
python - <<PY
import json
for w in ("_dbuf1", "_dbuf0"):
    try:
        d = json.load(open("$OUT/bench%s.json" % w))
        print(w, d["ms_per_step"], d["parity"]["rel_err"], d["parity"]["ok"])
        print("   ", {k: round(v["ms_per_step"], 4) for k, v in d["phases"].items()})
    except Exception as e:
        print(w, "ERR", e)
PY
