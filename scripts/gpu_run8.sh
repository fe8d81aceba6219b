python -m pytest tests -m gpu -x -q -k "call or map or config1" 2>&1 | tail -4
python bench.py --config 5 --steps 20 --warmup 5 > gpurun_out/r2_c5_1g.json 2> gpurun_out/r2_c5.err; tail -c 300 gpurun_out/r2_c5.err
python bench.py --config 3 --assemblies 4 > gpurun_out/r2_c3.json 2> gpurun_out/r2_c3.err; tail -c 300 gpurun_out/r2_c3.err
python bench.py --config 4 --assemblies 4 > gpurun_out/r2_c4.json 2> gpurun_out/r2_c4.err; tail -c 300 gpurun_out/r2_c4.err
python bench.py --config 3 --assemblies 2 --k 51 > gpurun_out/r2_c3_k51.json 2> gpurun_out/r2_c3k.err; tail -c 300 gpurun_out/r2_c3k.err
python - <<'PY'
import json
for f in ['r2_c5_1g','r2_c3','r2_c4','r2_c3_k51']:
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().split('\n')[-1])
        print(f, round(d['value']/1e9,3), 'ms/step', round(d['ms_per_step'],3), json.dumps(d.get('impl_detail'))[:600])
        if 'roofline' in d: r=d['roofline']; print('  roofline', r['bound'], round(r['frac'],3), r['kernel_ms'], r.get('random_sector'), 'e2e', round(d['e2e']['value']/1e9,2), r['events_per_base'])
    except Exception as ex: print(f,'ERR',ex)
PY
