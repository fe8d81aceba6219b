set -x
( time python -m pytest tests -m gpu -x -q ) 2>&1 | tail -6
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_final_1gpu.json 2> gpurun_out/r2_final.err; tail -c 300 gpurun_out/r2_final.err
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2_final_ref.json 2>> gpurun_out/r2_final.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 60 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 4 --warmup 2 --no-cpu-baseline --streams 1 > gpurun_out/r2_launches.log 2>&1
python bench.py --config 5 --ref-len 2000000000 --steps 10 --warmup 3 > gpurun_out/r2_c5_2g.json 2> gpurun_out/r2_c5_2g.err; tail -c 600 gpurun_out/r2_c5_2g.err
python - <<'PY'
import json
for f in ['r2_final_1gpu','r2_final_ref','r2_c5_2g']:
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().split('\n')[-1])
        print(f, 'value', round(d['value']/1e9,2), 'e2e', round(d['e2e']['value']/1e9,2), 'ms/step', d['ms_per_step'])
        if 'roofline' in d:
            r=d['roofline']; print('   ', r['bound'], round(r['frac'],3), r['kernel_ms'], d['impl_detail'].get('index_device_bytes'), d['impl_detail'].get('index_build_s'), d.get('cpu_baseline',{}).get('value'))
    except Exception as ex: print(f,'ERR',ex)
PY
