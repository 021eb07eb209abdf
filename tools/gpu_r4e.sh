#!/bin/bash
# two ranks over NVLink with the final tree: C3 bench line with the all-particle parity check
TAG=${1:-r4e}
OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 45 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --workload C3 --steps 20 --warmup 3 --no-cpu-baseline > $OUT/bench_2gpu_C3.json 2> $OUT/bench_2gpu_C3.err; echo "bench exit $?"
tail -c 1500 $OUT/bench_2gpu_C3.json | cut -c1-1500; tail -3 $OUT/bench_2gpu_C3.err | cut -c1-300
