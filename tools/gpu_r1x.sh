#!/bin/bash
# 2-GPU: slab tests (incl. pressure/laplacian) + strong-scaling bench line
OUT=gpurun_out/r1x; mkdir -p $OUT
timeout 900 python -m pytest tests/test_mgpu.py -x -q -m gpu > $OUT/pytest_mgpu.log 2>&1; tail -4 $OUT/pytest_mgpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 3 > $OUT/bench_2gpu.json 2> $OUT/bench_2gpu.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r1x/bench_2gpu.json') if l.startswith('{')][-1])
print('2gpu ms/step', round(d['ms_per_step'],4), 'value %.4g'%d['value'], {k:round(v['ms_per_step'],4) for k,v in d['phases'].items()})
PY
tail -2 $OUT/bench_2gpu.err
