mkdir -p gpurun_out/x1
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/x1/pytest.log 2>&1; tail -15 gpurun_out/x1/pytest.log
timeout 600 python bench.py --steps 20 --no-cpu-baseline > gpurun_out/x1/bench.json 2> gpurun_out/x1/bench.err; tail -3 gpurun_out/x1/bench.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/x1/bench.json') if l.startswith('{')][-1])
print('ms/step',round(d['ms_per_step'],3),'value %.3g'%d['value'], 'e2e %.3g'%d['e2e']['value'])
for k,v in d['phases'].items(): print('   ',k, round(v['ms_per_step'],3), v.get('frac_of_peak'))
PY
