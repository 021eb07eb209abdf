#!/bin/bash
# N ranks, final tree, the command the driver's scaling run uses (with e2e)
TAG=${1:-r3o}; N=${2:-2}
OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29547 \
    bench.py --gpus $N --steps 20 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"
python - <<PY
import json
t = open("$OUT/bench.json").read(); d = json.loads(t[t.index('{"metric"'):].splitlines()[0])
print(round(d["ms_per_step"], 4), "%.3e" % d["value"], "e2e %.3e" % d["e2e"]["value"], d["parity"]["rel_err"], d["parity"]["ok"])
print("   ", {k: round(v["ms_per_step"], 4) for k, v in d["phases"].items()})
print(d["config"]["parallelism"])
PY
