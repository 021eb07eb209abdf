#!/bin/bash
# N ranks over real NVLink: multi-GPU parity tests, then the C4 bench with the pipelined exchange on / off
TAG=${1:-r3d}; N=${2:-2}
OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi -L > $OUT/gpus.txt
timeout 900 python -m pytest tests/test_mgpu.py -x -q > $OUT/pytest_mgpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_mgpu.log
tail -4 $OUT/pytest_mgpu.log
for mode in 1 0 2; do
HYMD_B200_XPIPE=$mode timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$mode \
    bench.py --gpus $N --steps 30 --warmup 5 --no-e2e > $OUT/bench_xpipe$mode.json 2> $OUT/bench_xpipe$mode.err; echo "bench xpipe=$mode exit $?"
python - <<PY
import json
try:
    t = open("$OUT/bench_xpipe$mode.json").read(); d = json.loads(t[t.index('{"metric"'):].splitlines()[0])
    print("xpipe=$mode", round(d["ms_per_step"], 4), "%.3e" % d["value"], d["parity"]["rel_err"], d["parity"]["ok"])
    print("   ", {k: round(v["ms_per_step"], 4) for k, v in d["phases"].items()})
except Exception as e:
    print("xpipe=$mode ERR", e); print(open("$OUT/bench_xpipe$mode.err").read()[-1500:])
PY
done
