#!/bin/bash
# bash tools/gpu_scale.sh <tag> <ngpus> <modes: "strong weak"> [extra bench args]
TAG=${1:-sc}; NG=${2:-2}; MODES=${3:-strong}; shift 3
OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi -L > $OUT/gpus.txt
nvidia-smi topo -m > $OUT/topo.txt 2>&1
for mode in $MODES; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29533 \
      bench.py --gpus $NG --steps 20 --warmup 3 --scaling $mode "$@" > $OUT/bench_${NG}gpu_$mode.json 2> $OUT/bench_${NG}gpu_$mode.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open('$OUT/bench_${NG}gpu_$mode.json') if l.startswith('{')][-1])
    print('$NG GPUs $mode: ms/step', round(d['ms_per_step'],3), 'value %.3g'%d['value'], 'e2e', d['e2e'] and round(d['e2e']['ms_per_step'],3))
    print('   ', {k:round(x['ms_per_step'],3) for k,x in d['phases'].items()})
except Exception as e:
    print('no bench line', e)
PY
  tail -3 $OUT/bench_${NG}gpu_$mode.err
done
