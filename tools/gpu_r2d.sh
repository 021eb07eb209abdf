#!/bin/bash
# usage: gpu_r2d.sh <tag> <N>: mgpu tests, then C4 strong at N ranks with and without the fused pushes
TAG=${1:-r2d}; N=${2:-4}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi -L > $OUT/gpus.txt
timeout 1200 python -m pytest tests/test_mgpu.py -x -q -k "not single_gpu" > $OUT/pytest_mgpu.log 2>&1; echo "mgpu exit $?" >> $OUT/pytest_mgpu.log
tail -5 $OUT/pytest_mgpu.log
run() {  # name, env..., -- bench args
  name=$1; shift
  env "$@" timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $N --steps 30 --warmup 5 --no-e2e > $OUT/bench_$name.json 2> $OUT/bench_$name.err; echo "bench $name exit $?"
  python - <<PY
import json
try:
    t = open("$OUT/bench_$name.json").read(); d = json.loads(t[t.index('{"metric"'):].splitlines()[0])
    print("$name", round(d["ms_per_step"], 4), "%.3e" % d["value"], d["parity"]["rel_err"], d["parity"]["ok"])
    print("   ", {k: round(v["ms_per_step"], 4) for k, v in d["phases"].items()})
except Exception as e:
    print("$name ERR", e); print(open("$OUT/bench_$name.err").read()[-1500:])
PY
}
run fused HYMD_B200_FUSED_PUSH=1
run unfused HYMD_B200_FUSED_PUSH=0
run ncclbar HYMD_B200_NCCL_BARRIER=1
