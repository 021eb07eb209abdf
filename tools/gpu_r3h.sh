#!/bin/bash
# final single-GPU pass of the round: every GPU test, smoke, the default bench (as the driver runs it), the other
# workloads, launch list + ncu --set full of the hand-written kernels
TAG=${1:-r3h}
OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 1200 python -m pytest tests -x -q -m gpu --durations=8 > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log
tail -14 $OUT/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke exit $?"; tail -2 $OUT/smoke.log
timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"; tail -3 $OUT/bench.err
for w in C1 C2 C3; do
timeout 600 python bench.py --workload $w --steps 50 --warmup 5 --no-cpu-baseline > $OUT/bench_$w.json 2> $OUT/bench_$w.err; echo "bench $w exit $?"
done
timeout 600 python bench.py --dtype f64 --steps 20 --warmup 3 --no-cpu-baseline > $OUT/bench_C4_f64.json 2> $OUT/bench_C4_f64.err; echo "bench f64 exit $?"
python - <<PY
import json
for w in ("", "_C1", "_C2", "_C3", "_C4_f64"):
    try:
        d = json.loads([l for l in open("$OUT/bench%s.json" % w) if l.startswith("{")][-1])
        print(w or "C4", "ms/step", round(d["ms_per_step"], 4), "value %.4g" % d["value"], "e2e %.4g" % d["e2e"]["value"], d["parity"]["rel_err"], d["parity"]["ok"],
              "roofline", d["roofline"]["kernel"], round(d["roofline"]["frac"], 3))
        print("   ", {k: (round(v["ms_per_step"], 4), round(v.get("frac_of_peak") or 0, 3)) for k, v in d["phases"].items()})
    except Exception as e:
        print(w, "ERR", e)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv \
    --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"paint_kernel|count2_kernel|scatter2_kernel|plane_c2r|plane_r2c|xline_kernel|readout_gather" -s 21 -c 7 \
    -o $OUT/prof python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/ncu_full.log 2>&1
tail -2 $OUT/ncu_full.log | cut -c1-200
ls -la $OUT
