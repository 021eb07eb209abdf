#!/bin/bash
# N ranks, C4: grouped pipelined exchange (default planner) vs off
TAG=${1:-r3k}; N=${2:-4}
OUT=gpurun_out/$TAG; mkdir -p $OUT
run() {
name=$1; shift
env "$@" timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 \
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
run pipe HYMD_B200_XPIPE=1
run off HYMD_B200_XPIPE=0
