#!/bin/bash
# One GPU-box pass: parity tests, smoke, bench, ncu launch list and full capture of our kernels.
# Usage (from the repo root, under gpurun): bash tools/gpu_check.sh [tag]
TAG=${1:-r1}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu.txt 2>&1
python -c "import os; print('cpus', os.cpu_count())" >> $OUT/gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log
tail -5 $OUT/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; tail -2 $OUT/smoke.log
timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err; tail -c 3000 $OUT/bench.json; tail -5 $OUT/bench.err
timeout 600 python bench.py --workload C2 --steps 50 > $OUT/bench_C2.json 2>> $OUT/bench.err
timeout 600 python bench.py --workload C3 --steps 20 > $OUT/bench_C3.json 2>> $OUT/bench.err
timeout 600 python bench.py --dtype f64 --steps 10 --no-cpu-baseline > $OUT/bench_C4_f64.json 2>> $OUT/bench.err
# launch list (cold-cache, serialised: shares only)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/ncu_launch.log 2>&1
# full capture of the hand-written kernels (skip the warm-up launches)
timeout 1200 ncu --set full --clock-control none --import-source on \
    -k regex:'paint_kernel|kspace_force_kernel|readout_kernel|count_kernel|scatter_kernel|fill_ghost' -s 18 -c 6 \
    -o $OUT/prof python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/ncu_full.log 2>&1
ls -la $OUT
