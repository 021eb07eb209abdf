#!/bin/bash
# N ranks: pipelined exchange through the copy engines vs the SM copy kernel vs off; virtual-slab check first
TAG=${1:-r3e}; N=${2:-2}
OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_virtual_slabs.py -x -q -k "pipelined or exchange_modes" > $OUT/pytest_pipe.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_pipe.log
tail -3 $OUT/pytest_pipe.log
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "tensor_memory or field_forces_match or baseline" > $OUT/pytest_c2r.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_c2r.log
tail -3 $OUT/pytest_c2r.log
run() {
name=$1; shift
env "$@" timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
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
run ce HYMD_B200_XPIPE=1
run ce_novec HYMD_B200_XPIPE=1 HYMD_B200_C2R_VEC=0 HYMD_B200_READOUT_PERSIST=0
run off HYMD_B200_XPIPE=0
