#!/bin/bash
# ncu launch list of the C1 bench with the step replayed as a CUDA graph (kernel nodes are profiled one by one)
TAG=${1:-r4d}
OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 80 ncu --metrics gpu__time_duration.sum --clock-control none --graph-profiling node -c 700 --csv \
    --log-file $OUT/launches_C1_graph.csv python bench.py --workload C1 --steps 4 --warmup 8 --no-cpu-baseline --no-e2e > $OUT/ncu_launch.log 2>&1
echo "ncu exit $?"; tail -2 $OUT/ncu_launch.log | cut -c1-300; wc -l $OUT/launches_C1_graph.csv
