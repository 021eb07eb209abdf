#!/bin/bash
bash tools/gpu_profile.sh r1e_prof
OUT=gpurun_out/r1e_prof
timeout 600 python bench.py --workload C3 --no-cpu-baseline > $OUT/bench_C3.json 2> $OUT/bench_C3.err; tail -c 300 $OUT/bench_C3.json
timeout 600 python bench.py --workload C2 --no-cpu-baseline > $OUT/bench_C2.json 2> $OUT/bench_C2.err; tail -c 300 $OUT/bench_C2.json
timeout 600 python bench.py --dtype f64 --no-cpu-baseline > $OUT/bench_C4_f64.json 2> $OUT/bench_C4_f64.err; tail -c 300 $OUT/bench_C4_f64.json
