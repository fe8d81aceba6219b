python bench.py --config 4 > gpurun_out/r2g_c4.json 2> gpurun_out/r2g_c4.err; tail -c 200 gpurun_out/r2g_c4.err
python bench.py --config 3 > gpurun_out/r2g_c3.json 2> gpurun_out/r2g_c3.err; tail -c 200 gpurun_out/r2g_c3.err
python - <<'PY'
import json
for f in ('r2g_c4', 'r2g_c3'):
    d = json.loads(open('gpurun_out/%s.json' % f).read().strip().split('\n')[-1]); i = d['impl_detail']
    print(f, 'ms/asm', round(d['ms_per_step'], 1), 'M bases/s', round(d['value']/1e6), '| one thread', round(i['one_host_thread']['ms_per_assembly'], 1),
          {k[:20]: round(v, 1) for k, v in i['one_host_thread']['split_ms_per_assembly'].items()}, '| ref index once', round(i['reference_index_built_once']['ms_per_assembly'], 1))
PY
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "map or call or golden or config1" 2>&1 | tail -2
