#!/bin/bash
# multi-GPU pass: real NVLink tests + bench at N ranks.  usage: gpu_r2c.sh <tag> <N>
TAG=${1:-r2c}; N=${2:-2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export HYMD_B200_LOCAL_TIMEOUT_S=40
nvidia-smi -L > $OUT/gpus.txt
timeout 600 python -m pytest tests/test_gpu_virtual_slabs.py -x -q > $OUT/pytest_virtual.log 2>&1; echo "virtual exit $?" >> $OUT/pytest_virtual.log
tail -4 $OUT/pytest_virtual.log
timeout 1200 python -m pytest tests/test_mgpu.py -x -q --durations=10 > $OUT/pytest_mgpu.log 2>&1; echo "mgpu exit $?" >> $OUT/pytest_mgpu.log
tail -25 $OUT/pytest_mgpu.log
for n in $N; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $n --steps 20 --warmup 5 > $OUT/bench_$n.json 2> $OUT/bench_$n.err; echo "bench $n exit $?"
tail -c 600 $OUT/bench_$n.err
python - <<PY
import json
try:
    d = json.load(open("$OUT/bench_$n.json"))
    print($n, d["ms_per_step"], d["value"], d["parity"], d["e2e"]["value"] if d["e2e"] else None)
    print({k: round(v["ms_per_step"], 4) for k, v in d["phases"].items()})
except Exception as e:
    print("ERR", e)
PY
done
