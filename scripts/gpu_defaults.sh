( time python bench.py > gpurun_out/r2_default.json 2> gpurun_out/r2_default.err ) 2>&1 | tail -3
tail -c 300 gpurun_out/r2_default.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_default.json').read().strip().split('\n')[-1])
print('default run: steps', d['steps'], 'value', round(d['value']/1e9,1), 'e2e', round(d['e2e']['value']/1e9,1), 'single', round(d['impl_detail']['single_stream']['value']/1e9,1), 'frac', round(d['roofline']['frac'],3), 'cpu', d['cpu_baseline']['value'], d['cpu_baseline']['cores'])
PY
