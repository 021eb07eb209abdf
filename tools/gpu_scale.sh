#!/bin/bash
# bash tools/gpu_scale.sh <tag> <ngpus> [extra bench args]
TAG=${1:-sc}; NG=${2:-2}; shift 2
OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi -L > $OUT/gpus.txt
nvidia-smi topo -m > $OUT/topo.txt 2>&1
for mode in strong weak; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29533 \
      bench.py --gpus $NG --steps 20 --warmup 3 --scaling $mode "$@" > $OUT/bench_${NG}gpu_$mode.json 2> $OUT/bench_${NG}gpu_$mode.err
  tail -c 2500 $OUT/bench_${NG}gpu_$mode.json; tail -3 $OUT/bench_${NG}gpu_$mode.err
done
