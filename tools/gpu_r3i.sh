#!/bin/bash
# 8 ranks: C5 strong (pipelined exchange), C4 strong, C4-per-GPU weak
TAG=${1:-r3i}; N=${2:-8}
OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi -L > $OUT/gpus.txt
run() {
name=$1; shift
envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
env "${envs[@]}" timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29571 \
    bench.py --gpus $N "$@" > $OUT/bench_$name.json 2> $OUT/bench_$name.err; echo "bench $name exit $?"
python - <<PY
import json
try:
    t = open("$OUT/bench_$name.json").read(); d = json.loads(t[t.index('{"metric"'):].splitlines()[0])
    print("$name", round(d["ms_per_step"], 4), "%.3e" % d["value"], d["parity"].get("rel_err"), d["parity"]["ok"])
    print("   ", {k: round(v["ms_per_step"], 4) for k, v in d["phases"].items()})
except Exception as e:
    print("$name ERR", e); print(open("$OUT/bench_$name.err").read()[-1500:])
PY
}
run c5_strong X=1 -- --steps 10 --warmup 3 --no-e2e --workload C5
run c5_strong_nopipe HYMD_B200_XPIPE=0 -- --steps 10 --warmup 3 --no-e2e --workload C5
run c4_strong X=1 -- --steps 30 --warmup 5 --no-e2e
