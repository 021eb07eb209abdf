#!/bin/bash
# exchange modes at N ranks (C4 strong): blocked (copy kernel), blockedm (memcpy), fused, kernels
TAG=${1:-r2t}; N=${2:-2}
OUT=gpurun_out/$TAG; mkdir -p $OUT
for mode in blocked blockedm fused kernels; do
HYMD_B200_EXCHANGE=$mode timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $N --steps 30 --warmup 5 --no-e2e > $OUT/bench_$mode.json 2> $OUT/bench_$mode.err; echo "bench $mode exit $?"
python - <<PY
import json
try:
    t = open("$OUT/bench_$mode.json").read(); d = json.loads(t[t.index('{"metric"'):].splitlines()[0])
    print("$mode", round(d["ms_per_step"], 4), "%.3e" % d["value"], d["parity"]["rel_err"], d["parity"]["ok"])
    print("   ", {k: round(v["ms_per_step"], 4) for k, v in d["phases"].items()})
except Exception as e:
    print("$mode ERR", e); print(open("$OUT/bench_$mode.err").read()[-1500:])
PY
done
