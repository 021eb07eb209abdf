#!/bin/bash
# bash tools/gpu_sweep.sh ENVVAR v1 v2 ... : bench phases for each value of an env knob
VAR=$1; shift
for v in "$@"; do
  env $VAR=$v python bench.py --steps 10 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print('$VAR=$v', 'ms/step', round(d['ms_per_step'],3), {k:round(x['ms_per_step'],3) for k,x in d['phases'].items()})"
done
