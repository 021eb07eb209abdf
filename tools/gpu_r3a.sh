#!/bin/bash
# round 2, session 2, first pass: full GPU tests after the paint / c2r-addressing / binning changes, C4 bench,
# launch list + ncu --set full of the hand-written kernels
TAG=${1:-r3a}
OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 python -m pytest tests -x -q -m gpu --durations=8 > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log
tail -14 $OUT/pytest_gpu.log
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"; tail -3 $OUT/bench.err
python - <<PY
import json
try:
    d = json.loads([l for l in open("$OUT/bench.json") if l.startswith("{")][-1])
    print("ms/step", round(d["ms_per_step"], 4), "value %.4g" % d["value"], "e2e %.4g" % d["e2e"]["value"], d["parity"]["rel_err"], d["parity"]["ok"])
    for k, v in d["phases"].items(): print("   ", k, round(v["ms_per_step"], 4), v.get("frac_of_peak"))
except Exception as e:
    print("ERR", e)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv \
    --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"paint_kernel|count_kernel|scatter_kernel|plane_c2r|plane_r2c|xline_kernel|readout_gather" -s 21 -c 7 \
    -o $OUT/prof python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/ncu_full.log 2>&1
tail -2 $OUT/ncu_full.log | cut -c1-200
ls -la $OUT
