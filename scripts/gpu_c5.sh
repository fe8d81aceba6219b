# config 5 (HBM-resident index, 10^9 nodes) on the final tree
export PATH=/usr/local/cuda/bin:$PATH
timeout 200 python bench.py --config 5 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2m_c5_1g.json 2> gpurun_out/r2m_c5_1g.err; tail -c 300 gpurun_out/r2m_c5_1g.err
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/r2m_c5_1g.json').read().strip().split('\n')[-1]); r = d['roofline']
    print('c5 value', round(d['value']/1e9, 2), 'e2e', round(d['e2e']['value']/1e9, 2), 'K1 us', round(1e3*r['kernel_ms']['ms'], 1), r['bound'], 'frac', round(r['frac'], 3), {k: (round(v/1e9, 2) if isinstance(v, float) else v) for k, v in r.get('random_sector', {}).items()})
except Exception as ex:
    print('ERR', ex)
PY
